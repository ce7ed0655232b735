"""End-to-end parity of the CUDA engine through the reference-facing API (get_module() -> forward_for_eval) against
(a) the committed fixtures produced by the repaired reference and (b) the CPU oracle stage by stage.
Tolerance (BASELINE.json north_star): argmax masks bit-exact, mask logits within 1e-3 max-abs."""
import glob
import os

import numpy as np
import pytest
import torch

from gpu_util import from_T, report, to_T

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))


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
    report("embedding", emb, emb_w, 2e-4)


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
    from aocb200.sequence import run_sequence
    from aocb200.synth import make_clip
    g = torch.load(path)
    frames, labels = make_clip(g["seed"], g["H"], g["W"], g["K"], g["T"])
    first = labels[0].clone()
    if g["drop"] is not None:
        first[first == g["drop"]] = 0
    hk = Hooked(model)
    np.random.seed(g["seed"])
    preds = run_sequence(hk, frames, first, g["K"], mem_every=g["mem_every"], unc_ratio=1.0, device=torch.device("cuda:0"))
    worst, ok = 0.0, True
    for t, (a, b) in enumerate(zip(hk.logits, g["logits"])):
        d = (a - b).abs().max().item()
        eq = (preds[t].cpu().to(torch.uint8) == g["preds"][t]).float().mean().item()
        print("[parity] %s frame %d: max|dlogit|=%.3e argmax-equal=%.6f" % (os.path.basename(path), t + 1, d, eq))
        worst = max(worst, d)
        ok = ok and eq == 1.0
    assert worst <= 1e-3, worst
    assert ok, "argmax masks differ from the reference"

"""Pins the CPU oracle (oracle/aoc_oracle.py) against fixtures produced by the REPAIRED REFERENCE itself
(tools/make_golden.py, generated in the build container from /root/reference)."""
import glob
import os

import numpy as np
import pytest
import torch

from aocb200.sequence import run_sequence
from aocb200.synth import make_clip
from oracle.aoc_oracle import AOCOracle, kmeans2_points

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tiny_*.pt"))
              if not p.endswith("_fp64.pt"))


class Hooked:
    def __init__(self, inner):
        self.inner, self.logits = inner, []

    def forward_for_eval(self, *a, **k):
        out = self.inner.forward_for_eval(*a, **k)
        if out[0] is not None:
            self.logits.append(self.inner.last_logits.clone())
        return out


def load_case(path):
    g = torch.load(path)
    frames, labels = make_clip(g["seed"], g["H"], g["W"], g["K"], g["T"])
    first = labels[0].clone()
    if g["drop"] is not None:
        first[first == g["drop"]] = 0
    return g, frames, first


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-3] for p in GOLD])
def test_oracle_matches_reference_fixture(path, state_dict):
    torch.set_num_threads(os.cpu_count())
    g, frames, first = load_case(path)
    assert g["weights_seed"] == 1234
    o = Hooked(AOCOracle(state_dict))
    with torch.no_grad():
        np.random.seed(g["seed"])
        preds = run_sequence(o, frames, first, g["K"], mem_every=g["mem_every"], unc_ratio=1.0)
    assert len(preds) == len(g["preds"]) == g["T"] - 1
    for t, (a, b) in enumerate(zip(o.logits, g["logits"])):
        assert (a - b).abs().max().item() < 5e-4, (t, (a - b).abs().max().item())
        assert torch.equal(preds[t].to(torch.uint8), g["preds"][t]), t


def test_kmeans_restatement_matches_scipy():
    from scipy.cluster.vq import kmeans2
    import warnings
    rs = np.random.RandomState(3)
    for n, k in ((500, 16), (40, 16), (16, 16), (300, 5)):
        x = np.abs(rs.randn(n, 100)).astype(np.float32)
        x[: n // 3] += 2.0
        np.random.seed(n + k)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            c0, l0 = kmeans2(x, k, minit="points", iter=20)
        np.random.seed(n + k)
        c1, l1 = kmeans2_points(x, k, 20)
        assert np.array_equal(l0, l1)
        assert np.abs(c0 - c1).max() < 1e-5

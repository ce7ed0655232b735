"""Pins the oracle against the reference and writes tests/golden/*.pt  (run in the build container).

    python tools/make_golden.py

1. loads the repaired reference from /root/reference (tools/ref_loader.py) and the oracle
   restatement (oracle/aoc_oracle.py) with the same synthetic state_dict;
2. runs both through the same per-sequence driver on the same synthetic clips with the same
   numpy RNG seed (k-means init stream) and asserts agreement;
3. stores the REFERENCE's outputs (not the oracle's) as small fixtures, so that on machines
   without /root/reference the oracle is still checked against the real thing.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aocb200.params import synthetic_state_dict  # noqa: E402
from aocb200.sequence import run_sequence  # noqa: E402
from aocb200.synth import make_clip  # noqa: E402
from oracle.aoc_oracle import AOCOracle  # noqa: E402
from tools.ref_loader import load_reference, make_reference_model  # noqa: E402

CASES = [
    # name, seed, H, W, K, T, mem_every, drop object id from first label (absent-object quirk) or None
    ("tiny_k2", 11, 97, 129, 2, 4, 2, None),
    ("tiny_k3_absent", 12, 97, 113, 3, 3, 2, 2),
    ("tiny_k1", 13, 65, 97, 1, 3, 5, None),
]


class _Hooked:
    """Wraps a forward_for_eval provider and records logits per frame."""

    def __init__(self, inner, get_logits):
        self.inner, self.get_logits, self.logits = inner, get_logits, []

    def forward_for_eval(self, *a, **k):
        out = self.inner.forward_for_eval(*a, **k)
        if out[0] is not None:
            self.logits.append(self.get_logits())
        return out


def main():
    torch.set_num_threads(os.cpu_count())
    sd = synthetic_state_dict(1234)
    ref_model = make_reference_model()
    ref_model.load_state_dict(sd)
    cap = {}
    ref_model.dynamic_seghead.register_forward_hook(lambda m, i, o: cap.__setitem__("logits", o[0].detach().clone()))
    oracle = AOCOracle(sd)
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, seed, H, W, K, T, mem_every, drop in CASES:
        frames, labels = make_clip(seed, H, W, K, T)
        first = labels[0].clone()
        if drop is not None:
            first[first == drop] = 0
        with torch.no_grad():
            np.random.seed(seed)
            r = _Hooked(ref_model, lambda: cap["logits"])
            ref_preds, ref_probs = run_sequence(r, frames, first, K, mem_every=mem_every, unc_ratio=1.0, keep_probs=True)
            np.random.seed(seed)
            o = _Hooked(oracle, lambda: oracle.last_logits.clone())
            ora_preds, ora_probs = run_sequence(o, frames, first, K, mem_every=mem_every, unc_ratio=1.0, keep_probs=True)
        worst = 0.0
        for t, (a, b) in enumerate(zip(r.logits, o.logits)):
            d = (a - b).abs().max().item()
            worst = max(worst, d)
            eq = (ref_preds[t] == ora_preds[t]).float().mean().item()
            print("%s frame %d: max|dlogit|=%.3e argmax-equal=%.6f logit-range=[%.2f,%.2f]" %
                  (name, t + 1, d, eq, a.min().item(), a.max().item()))
            assert eq == 1.0, "oracle argmax differs from the reference"
        assert worst < 5e-4, worst  # both CPU fp32; differences are op-composition rounding (logit range ~25)
        torch.save({
            "seed": seed, "H": H, "W": W, "K": K, "T": T, "mem_every": mem_every, "drop": drop,
            "weights_seed": 1234,
            "logits": r.logits,# reference pre-upsample logits [1,O,h,w]
            "preds": [p.to(torch.uint8) for p in ref_preds],           # reference argmax masks [H,W]
        }, os.path.join(out_dir, name + ".pt"))
        print("wrote", name)


if __name__ == "__main__":
    load_reference()
    main()

"""Writes tests/golden/full480_k5_fp64.pt: frame 1 of a 481x849, 5-object clip evaluated by the oracle in float32 and
float64 (same scheme as tools/make_fp64_truth.py) -- the yardstick for fp32 noise at BASELINE.json's full size, where
the summations (GroupNorm over 25 773 pixels, K = 18 432 convolutions, logit range +-40) are longer than in the tiny
fixtures.  Run: python tools/make_fullsize_truth.py   (several minutes of CPU; needs no /root/reference)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aocb200.params import synthetic_state_dict  # noqa: E402
from aocb200.synth import make_clip, restrict_size  # noqa: E402
from tools.make_fp64_truth import _km32, oracle64  # noqa: E402

SEED, K = 3, 5


def frame1(orc, frames, first, dt):
    H, W = frames.shape[2:]
    gt = torch.tensor([K])
    with torch.no_grad():
        _, emb, mem = orc.forward_for_eval([[None, None]], [], [], None, None, frames[0:1].to(dt), [H, W], gt)
        lab = first.view(1, 1, H, W)
        np.random.seed(SEED)
        probs, _, _ = orc.forward_for_eval(mem, [emb], [lab], emb, lab, frames[1:2].to(dt), [H, W], gt)
    return orc.last_logits.clone(), torch.argmax(probs[0], 0).to(torch.uint8)


def main():
    torch.set_num_threads(os.cpu_count())
    from oracle.aoc_oracle import AOCOracle
    H, W = restrict_size(480, 854)
    frames, labels = make_clip(SEED, H, W, K, 2)
    sd = synthetic_state_dict(1234)
    l32, p32 = frame1(AOCOracle(sd), frames, labels[0], torch.float32)
    m64 = oracle64()
    l64, p64 = frame1(m64.AOCOracle({k: v.double() for k, v in sd.items()}, kmeans_fn=_km32), frames, labels[0], torch.float64)
    noise = (l32.double() - l64).abs().max().item()
    top2 = torch.topk(l64[0], 2, dim=0)[0]
    print("480p K=5 frame 1: oracle fp32 vs fp64 max|dlogit| %.3e (logit range %.1f); argmax fp32 vs fp64 differs at %d px"
          % (noise, l64.abs().max().item(), int((p32 != p64).sum())))
    torch.save({"seed": SEED, "K": K, "H": H, "W": W, "logits_fp64": l64, "oracle32_noise": noise,
                "pred_fp32": p32, "margin_fp64": (top2[0] - top2[1]).float()},
               os.path.join(ROOT, "tests", "golden", "full480_k5_fp64.pt"))


if __name__ == "__main__":
    main()

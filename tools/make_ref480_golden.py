"""Writes tests/golden/full480_k5_ref.npz: the REPAIRED REFERENCE itself (tools/ref_loader.py, /root/reference) on
BASELINE.json's headline configuration -- cfg3: 480p (481x849 after MultiRestrictSize), 5 objects -- so that the oracle
(and, on the GPU, the engine) is pinned to the reference at FULL size and not only on the tiny clips of
tools/make_golden.py.  Two predicted frames; the second is teacher-forced: both sides are handed the REFERENCE's
frame-1 label map as previous mask and second bank mask (a label map that differs in three near-tie pixels changes the
per-object row counts, hence numpy's k-means draws, hence everything), and sees a two-frame bank and a filled decoder
memory.

    python tools/make_ref480_golden.py          (build container only; ~3 minutes of CPU)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aocb200.params import synthetic_state_dict  # noqa: E402
from aocb200.synth import make_clip, restrict_size  # noqa: E402

SEED, K, N_PRED = 3, 5, 2


def frame_seed(t):
    return SEED if t == 1 else 100 * SEED + t


def drive(model, get_logits, frames, first, fed=None):
    """-> per predicted frame: (logits [1,O,h,w], argmax [H,W] uint8).  fed: label maps handed back instead of the model's own"""
    H, W = frames.shape[2:]
    gt = torch.tensor([K])
    out = []
    with torch.no_grad():
        _, emb, mem = model.forward_for_eval([[None, None]], [], [], None, None, frames[0:1], [H, W], gt)
        lab = first.view(1, 1, H, W).long()
        refs, masks, prev_e, prev_m = [emb], [lab], emb, lab
        for t in range(1, N_PRED + 1):
            np.random.seed(frame_seed(t))
            probs, emb, mem = model.forward_for_eval(mem, refs, masks, prev_e, prev_m, frames[t:t + 1], [H, W], gt)
            pred = torch.argmax(probs[0], 0).to(torch.uint8)
            out.append((get_logits().clone(), pred))
            m = (fed[t - 1] if fed is not None else pred).view(1, 1, H, W).long()
            refs.append(emb); masks.append(m)
            prev_e, prev_m = emb, m
    return out


def main():
    torch.set_num_threads(os.cpu_count())
    from tools.ref_loader import load_reference, make_reference_model
    load_reference()
    from oracle.aoc_oracle import AOCOracle
    H, W = restrict_size(480, 854)
    frames, labels = make_clip(SEED, H, W, K, N_PRED + 1)
    sd = synthetic_state_dict(1234)
    ref = make_reference_model()
    ref.load_state_dict(sd)
    cap = {}
    ref.dynamic_seghead.register_forward_hook(lambda m, i, o: cap.__setitem__("logits", o[0].detach().clone()))
    r = drive(ref, lambda: cap["logits"], frames, labels[0])
    fed = [p for _, p in r]
    orc = AOCOracle(sd)
    o = drive(orc, lambda: orc.last_logits, frames, labels[0], fed=fed)
    for t, ((lr, pr), (lo, po)) in enumerate(zip(r, o)):
        d = (lr - lo).abs().max().item()
        print("frame %d: oracle vs reference max|dlogit| %.3e, argmax differs at %d of %d px, logit range [%.1f, %.1f]"
              % (t + 1, d, int((pr != po).sum()), pr.numel(), lr.min().item(), lr.max().item()), flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "full480_k5_ref.npz"), seed=SEED, K=K, H=H, W=W,
                        n_pred=N_PRED, weights_seed=1234,
                        logits=np.stack([l.numpy().reshape(K + 1, -1) for l, _ in r]).astype(np.float32),
                        preds=np.stack([p.numpy() for _, p in r]),
                        oracle_dist=np.array([(lr - lo).abs().max().item() for (lr, _), (lo, _) in zip(r, o)]),
                        oracle_px=np.array([int((pr != po).sum()) for (_, pr), (_, po) in zip(r, o)]))
    print("wrote full480_k5_ref.npz")


if __name__ == "__main__":
    main()

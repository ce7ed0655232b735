"""Writes tests/golden/cfg4_720p_k10.npz and tests/golden/cfg5_1080p_k5_bank3.npz: oracle parity fixtures at
BASELINE.json's configs[3] / configs[4] sizes (the CPU oracle needs minutes per frame there, so the GPU tests compare
with these committed vectors instead of running it).

    cfg4  721x1281 (features 181x321), 10 objects, frame 1 predicted from a one-frame bank
    cfg5  1073x1921 (features 269x481), 5 objects, frames 1..3 predicted with the bank growing every frame -> the
          compared frame 3 sees a three-frame bank

Scheme of tools/make_fullsize_truth.py: the oracle in float32 (= the reference's arithmetic) and in float64 (the
"exact" yardstick) on the same clip / weights / numpy RNG stream; the float64 run is fed the float32 run's predicted
label maps so that both see the same bank masks and k-means draws (the draws depend on per-object pixel counts).

k-means is a DISCRETE step: at these sizes a 1e-6 change of an embedding flips the cluster label of a few boundary rows,
the centroids move by ~1e-2 and the logits by 0.1-0.3 (first attempt of this fixture: float32 vs float64 differed by
2.9e-1 for exactly that reason) -- a property of the reference algorithm, not of an evaluation.  So the float32 run's
scipy results (code book + labels per object) are RECORDED and REPLAYED in the float64 run, and the compared frame's
proxies of EVERY predicted frame (centroid, centroid_avg per object) are stored so that the GPU test can pin the engine to the same proxies and
compare logits at fp32-rounding level; the engine's own k-means is checked bit for bit on identical inputs elsewhere
(tests/test_gpu_ops.py::test_kmeans_vs_restatement).

Stored: the label maps fed back (uint8), the compared frame's logits of both runs (float64 rounded to float32:
2^-24 relative, three orders below the fp32 noise it measures), the float32 run's distance to float64, its argmax, and
the float32 run's proxies of the compared frame.

Run: python tools/make_cfg_truth.py [cfg4|cfg5]   (tens of minutes of CPU; needs no /root/reference)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aocb200.params import synthetic_state_dict  # noqa: E402
from aocb200.synth import make_clip, restrict_size  # noqa: E402
from tools.make_fp64_truth import oracle64  # noqa: E402

CASES = {
    "cfg4": dict(file="cfg4_720p_k10.npz", src=(720, 1280), K=10, seed=11, n_pred=1),
    "cfg5": dict(file="cfg5_1080p_k5_bank3.npz", src=(1080, 1920), K=5, seed=12, n_pred=3),
}


class KMRecord:
    """scipy kmeans2 as the oracle calls it, keeping every (code book, labels) result in call order"""

    def __init__(self):
        self.calls = []

    def __call__(self, x, k):
        from oracle.aoc_oracle import _scipy_kmeans2
        c, l = _scipy_kmeans2(x, k)
        self.calls.append((np.array(c, copy=True), np.array(l, copy=True)))
        return c, l


class KMReplay:
    def __init__(self, calls):
        self.it = iter(calls)

    def __call__(self, x, k):
        c, l = next(self.it)
        assert c.shape[0] == k and l.shape[0] == x.shape[0]
        return c.astype(x.dtype), l


def record_proxies(mod, sink):
    """wrap <oracle module>.adaptive_proxies so that the (centroid, centroid_avg) lists of every frame land in `sink`"""
    inner = mod.adaptive_proxies

    def wrapped(*a, **k):
        out = inner(*a, **k)
        sink.append(out)
        return out
    mod.adaptive_proxies = wrapped
    return inner


def frame_seed(seed, t):
    return seed if t == 1 else 100 * seed + t


def run(orc, frames, first, K, seed, n_pred, dt, fed=None):
    """-> (logits of frame n_pred, its argmax [H,W] uint8, label maps predicted for frames 1..n_pred-1).
    fed: label maps to feed back instead of this run's own predictions."""
    H, W = frames.shape[2:]
    gt = torch.tensor([K])
    own = []
    with torch.no_grad():
        _, emb, mem = orc.forward_for_eval([[None, None]], [], [], None, None, frames[0:1].to(dt), [H, W], gt)
        lab = first.view(1, 1, H, W).long()
        refs, masks, prev_e, prev_m = [emb], [lab], emb, lab
        for t in range(1, n_pred + 1):
            t0 = time.time()
            np.random.seed(frame_seed(seed, t))
            probs, emb, mem = orc.forward_for_eval(mem, refs, masks, prev_e, prev_m, frames[t:t + 1].to(dt), [H, W], gt)
            pred = torch.argmax(probs[0], 0).to(torch.uint8)
            print("  frame %d (%s, bank %d): %.0f s" % (t, str(dt).split(".")[-1], len(refs), time.time() - t0), flush=True)
            if t < n_pred:
                own.append(pred)
                m = (fed[t - 1] if fed is not None else pred).view(1, 1, H, W).long()
                refs.append(emb); masks.append(m)
                prev_e, prev_m = emb, m
    return orc.last_logits.clone(), pred, own


def make(name):
    c = CASES[name]
    from oracle.aoc_oracle import AOCOracle
    H, W = restrict_size(c["src"][0], c["src"][1], 10 ** 9)
    K, seed, n_pred = c["K"], c["seed"], c["n_pred"]
    frames, labels = make_clip(seed, H, W, K, n_pred + 1)
    sd = synthetic_state_dict(1234)
    print(name, "%dx%d K=%d, %d predicted frame(s)" % (H, W, K, n_pred), flush=True)
    import oracle.aoc_oracle as m32
    rec, prox = KMRecord(), []
    inner = record_proxies(m32, prox)
    try:
        l32, p32, fed = run(AOCOracle(sd, kmeans_fn=rec), frames, labels[0], K, seed, n_pred, torch.float32)
    finally:
        m32.adaptive_proxies = inner
    m64 = oracle64()
    l64, p64, _ = run(m64.AOCOracle({k: v.double() for k, v in sd.items()}, kmeans_fn=KMReplay(rec.calls)), frames,
                      labels[0], K, seed, n_pred, torch.float64, fed=fed)
    O = K + 1
    assert len(prox) == n_pred                              # one adaptive_proxies call per predicted frame
    cen, avg = np.zeros((n_pred, O, 16, 100), np.float32), np.zeros((n_pred, O, 16, 100), np.float32)
    ncen, navg = np.zeros((n_pred, O), np.int32), np.zeros((n_pred, O), np.int32)
    for t, frame_prox in enumerate(prox):                  # every predicted frame's proxies (the decoder memory carries
        for o, pr in enumerate(frame_prox):                # earlier frames' features into the compared frame)
            if pr is not None:
                ncen[t, o], navg[t, o] = pr[0].shape[0], pr[1].shape[0]
                cen[t, o, :ncen[t, o]] = pr[0].numpy(); avg[t, o, :navg[t, o]] = pr[1].numpy()
    noise = (l32.double() - l64).abs().max().item()
    print("%s frame %d: oracle fp32 vs fp64 max|dlogit| %.3e (logit range %.1f); argmax fp32 vs fp64 differs at %d px"
          % (name, n_pred, noise, l64.abs().max().item(), int((p32 != p64).sum())), flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", c["file"]),
                        seed=seed, K=K, H=H, W=W, n_pred=n_pred,
                        fed=np.stack([m.numpy() for m in fed]) if fed else np.zeros((0, H, W), np.uint8),
                        logits_fp32=l32.numpy(), logits_fp64=l64.float().numpy(), oracle32_noise=noise,
                        pred_fp32=p32.numpy(), prox_cen=cen, prox_avg=avg, prox_ncen=ncen, prox_navg=navg)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for n in (sys.argv[1:] or ["cfg4", "cfg5"]):
        make(n)

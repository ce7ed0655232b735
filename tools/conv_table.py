"""Per-layer table of the convolution calls of one steady-state frame (480p, 5 objects): shape, flags, CUDA-event time.
Timed INSIDE the replayed CUDA graphs of the product path (external event pairs = event-record nodes around each call,
see aocb200/lib.py), so a row carries no host launch gap."""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import Stepper
    from aocb200.lib import lib
    from aocb200.model import get_module
    from aocb200.params import synthetic_state_dict
    from aocb200.synth import make_clip, restrict_size
    K = 5
    H, W = restrict_size(480, 854)
    frames, labels = make_clip(0, H, W, K, 12)
    dev = torch.device("cuda:0")
    model = get_module()(None, None)
    model.load_state_dict(synthetic_state_dict(1234))
    model = model.cuda(0).eval()
    np.random.seed(0)
    L = lib()
    L.profile = {"aoc_conv2d_nhwc_tc": []}
    st = Stepper(model, frames, labels[0], K, dev, False)
    for _ in range(3):
        st.step()
    names = [n for _, n in L.protos["aoc_conv2d_nhwc_tc"][1]]
    ix = {n: names.index(n) for n in names}
    reps = 3
    prof = []
    for _ in range(reps):
        L.profile["aoc_conv2d_nhwc_tc"] = []
        L.replayed = set()
        st.step()
        torch.cuda.synchronize()
        prof += [(e0.elapsed_time(e1), a) for e0, e1, a in L.profile["aoc_conv2d_nhwc_tc"]]
        prof += [(e0.elapsed_time(e1), a) for e0, e1, a, tag in L.profile_graph.get("aoc_conv2d_nhwc_tc", ())
                 if tag in L.replayed]
    L.profile = None
    per = len(prof) // reps
    tab = OrderedDict()
    for i, (t_ms, a) in enumerate(prof):
        g = lambda n: a[ix[n]]
        key = (g("N"), g("H"), g("W"), g("Cin"), g("Cout"), g("kh"), g("stride"), g("dil"),
               "aff" if (g("in_a") or g("in_b") or g("in_relu")) else "-", "res" if g("residual") else "-",
               "stats" if g("tile_stats") else "-")
        t = tab.setdefault(key, [0, 0.0])
        t[0] += 1
        t[1] += t_ms * 1e3
    tot = sum(v[1] for v in tab.values()) / reps
    print("%d conv calls per frame, %.1f us per frame" % (per, tot))
    print("%3s %4s %4s %5s %5s k s d  %-5s %-4s %-6s %5s %9s %9s %8s" % ("N", "H", "W", "Cin", "Cout", "aff", "res", "stats", "calls",
                                                                   "us/call", "us/frame", "TFLOP/s"))
    for key, (c, us) in sorted(tab.items(), key=lambda kv: -kv[1][1]):
        N, H_, W_, Cin, Cout, k, s, d, aff, res, stt = key
        Ho, Wo = (H_ - 1) // s + 1, (W_ - 1) // s + 1
        fl = 2.0 * N * Ho * Wo * Cout * k * k * Cin
        print("%3d %4d %4d %5d %5d %d %d %2d %-5s %-4s %-6s %5d %9.1f %9.1f %8.1f" %
              (N, H_, W_, Cin, Cout, k, s, d, aff, res, stt, c // reps, us / c, us / reps, fl / (us / c) / 1e6))


if __name__ == "__main__":
    main()

"""Runs a synthetic clip through the CUDA engine (used under ncu: the profiled range is bracketed with
cudaProfilerStart/Stop so `ncu --profile-from-start off` skips warm-up).

    python tools/run_clip.py [--frames 2] [--warm 2] [--H 480 --W 854] [--K 5] [--mem-every 5]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--H", type=int, default=480)
    ap.add_argument("--W", type=int, default=854)
    ap.add_argument("--K", type=int, default=5)
    ap.add_argument("--native", action="store_true", help="skip MultiRestrictSize's max-size clamp")
    args = ap.parse_args()
    from bench import Stepper
    from aocb200.model import get_module
    from aocb200.params import synthetic_state_dict
    from aocb200.synth import make_clip, restrict_size
    H, W = restrict_size(args.H, args.W, 10 ** 9 if args.native else 1040)
    frames, labels = make_clip(0, H, W, args.K, 1 + args.warm + args.frames)
    dev = torch.device("cuda:0")
    model = get_module()(None, None)
    model.load_state_dict(synthetic_state_dict(1234))
    model = model.cuda(0).eval()
    np.random.seed(0)
    st = Stepper(model, frames, labels[0], args.K, dev, False)
    for _ in range(args.warm):
        st.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    e0.record()
    t0 = time.perf_counter()
    for _ in range(args.frames):
        st.step()
    t_issue = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    print("host issue time %.3f ms/frame (python + launches, GPU running behind)" % (1e3 * t_issue / args.frames))
    torch.cuda.profiler.stop()
    print("frames %d  %.3f ms/frame (%dx%d, K=%d)" % (args.frames, e0.elapsed_time(e1) / args.frames, H, W, args.K))
    bk = model.engine().bank
    print("bank at the profiled frame: %d frame(s) of %d pixels, %d object-sorted rows (padded), per-object pixel counts %s"
          % (bk.n, bk.hw, bk.index["rows"], bk.index["counts"]))


if __name__ == "__main__":
    main()

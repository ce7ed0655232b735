"""Upper bound of what overlapping independent work buys on ONE GPU: two independent sequences (two engines, two streams,
one host thread that alternates their steps) against one sequence alone.  The frame is a strict chain of ~380 kernels, a
third of them latency-bound launches on the 31x54 backbone maps that leave most SMs idle; a second, independent chain can
only use what the first leaves free (the persistent convolution CTAs take a whole SM each).  Prints frames/s of one
sequence, of two interleaved sequences (aggregate), and the ratio -- the ceiling for any cross-frame pipelining of the
backbone under the previous frame's decoder."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from aocb200.model import get_module  # noqa: E402
from aocb200.params import synthetic_state_dict  # noqa: E402


def make_model(dev):
    m = get_module()(None, None)
    m.load_state_dict(synthetic_state_dict(1234))
    return m.to(dev).eval()


def run(steppers, streams, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    for s in streams:
        s.wait_event(e0)
    t0 = time.perf_counter()
    for _ in range(steps):
        for st, s in zip(steppers, streams):
            with torch.cuda.stream(s):
                st.step()
    host = time.perf_counter() - t0
    for s in streams:
        cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), host * 1e3


def main():
    dev = torch.device("cuda:0")
    steps, warm = 20, 6
    frames, first, _ = bench.make_workload(7, 1 + warm + steps + 1)
    prio = "--prio" in sys.argv
    out = {}
    for n in (1, 2, 1, 2):
        np.random.seed(5)
        models = [make_model(dev) for _ in range(n)]
        streams = [torch.cuda.Stream(device=dev, priority=(-1 if (prio and i == 0) else 0)) for i in range(n)]
        sts = []
        for m, s in zip(models, streams):
            with torch.cuda.stream(s):
                sts.append(bench.Stepper(m, frames, first, bench.K_OBJ, dev, False))
        run(sts, streams, warm)
        ms, host = run(sts, streams, steps)
        fps = 1e3 * n * steps / ms
        out.setdefault(n, []).append(fps)
        print("%d sequence(s): %.2f ms for %d x %d frames = %.1f frames/s aggregate (host %.1f ms)" % (n, ms, n, steps, fps, host),
              flush=True)
        del sts, models
        torch.cuda.empty_cache()
    print("two interleaved sequences / one sequence: %.3f" % (max(out[2]) / max(out[1])))


if __name__ == "__main__":
    main()

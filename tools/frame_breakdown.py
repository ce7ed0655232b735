"""Host / device time of the three captured segments of a steady-state frame (480p, K objects)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import Stepper
    from aocb200.engine import _Graph
    from aocb200.model import get_module
    from aocb200.params import synthetic_state_dict
    from aocb200.synth import make_clip, restrict_size
    K = 5
    H, W = restrict_size(480, 854)
    frames, labels = make_clip(0, H, W, K, 40)
    dev = torch.device("cuda:0")
    model = get_module()(None, None)
    model.load_state_dict(synthetic_state_dict(1234))
    model = model.cuda(0).eval()
    np.random.seed(0)
    st = Stepper(model, frames, labels[0], K, dev, False)
    for _ in range(7):
        st.step()
    torch.cuda.synchronize()
    # instrument graph replays
    rec = []
    orig = _Graph.replay

    def replay(self):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        orig(self)
        e1.record()
        rec.append((self.n, e0, e1, time.perf_counter() - t0))
    _Graph.replay = replay
    for i in range(3):
        torch.cuda.synchronize()
        rec.clear()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.step()
        e1.record()
        th = time.perf_counter() - t0
        torch.cuda.synchronize()
        print("frame %d (t=%d): host %.2f ms, device %.2f ms; segments (kernels, device ms, host ms): %s" % (
            i, st.t, 1e3 * th, e0.elapsed_time(e1),
            ", ".join("(%d, %.2f, %.2f)" % (n, a.elapsed_time(b), 1e3 * h) for n, a, b, h in rec)))


if __name__ == "__main__":
    main()

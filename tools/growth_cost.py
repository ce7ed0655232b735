"""Host time of the pieces of a bank-change step (index rebuild + its one synchronisation, RNG draws, re-capture of the
bank-dependent graph), bench schedule (480p, 5 objects, bank +1 frame every 5 steps)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import Stepper, make_workload
    from aocb200.engine import Engine
    from aocb200.model import get_module
    from aocb200.params import synthetic_state_dict
    n = 22
    frames, first, _ = make_workload(0, n + 1)
    dev = torch.device("cuda:0")
    model = get_module()(None, None)
    model.load_state_dict(synthetic_state_dict(1234))
    model = model.cuda(0).eval()
    rec = {}

    def wrap(name):
        inner = getattr(Engine, name)

        def f(self, *a, **k):
            t0 = time.perf_counter()
            out = inner(self, *a, **k)
            rec[name] = rec.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
            return out
        setattr(Engine, name, f)
    for nm in ("_bank_index", "_capture", "_draw_kmeans_init", "_sync_bank"):
        wrap(nm)
    sync = "sync" in sys.argv           # drain the device before every step: the index synchronisation then costs nothing
    for rep in range(2):
        np.random.seed(1000)
        st = Stepper(model, frames, first, 5, dev, False)
        for i in range(n):
            if sync:
                torch.cuda.synchronize()
            rec.clear()
            t0 = time.perf_counter()
            st.step()
            th = 1e3 * (time.perf_counter() - t0)
            if rep == 1 and (th > 3.0 or i % 5 == 0):
                print("step %2d (t=%d): host %.2f ms  %s" % (i, st.t, th, "  ".join("%s %.2f" % kv for kv in sorted(rec.items()))))
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()

"""Per-step device time over the bench schedule (480p, 5 objects, bank +1 frame every 5 steps)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import Stepper, make_workload
    from aocb200.model import get_module
    from aocb200.params import synthetic_state_dict
    n = 30
    frames, first, _ = make_workload(0, n + 1)
    dev = torch.device("cuda:0")
    model = get_module()(None, None)
    model.load_state_dict(synthetic_state_dict(1234))
    model = model.cuda(0).eval()
    smi = None
    if "smi" in sys.argv:                                   # with bench.py's nvidia-smi clock sampler polling alongside
        from bench import ClockSampler
        smi = ClockSampler(0)
        smi.start()
    for rep in range(3 if smi else 2):
        np.random.seed(1000)
        e2e = len(sys.argv) > 1 and sys.argv[1] == "e2e"      # host buffers + a stream sync per step, as in bench.py's e2e arm
        st = Stepper(model, frames, first, 5, dev, e2e)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        host = []
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(n):
            t0 = time.perf_counter()
            st.step()
            host.append(1e3 * (time.perf_counter() - t0))
            ev[i + 1].record()
        torch.cuda.synchronize()
        dv = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
        print("pass %d: total %.1f ms, %.2f ms/step" % (rep, sum(dv), sum(dv) / n))
        print("  device ms:", " ".join("%.1f" % v for v in dv))
        print("  host ms:  ", " ".join("%.1f" % v for v in host))
    if smi:
        print(smi.stop())


if __name__ == "__main__":
    main()

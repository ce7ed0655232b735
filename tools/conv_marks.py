"""Whole-kernel timeline of CTA 0 of one convolution launch (trace build: AOCB200_BUILD_TAG=trace
AOCB200_NVCC_FLAGS=-DAOC_CONV_TRACE python -m aocb200.build; run with AOCB200_LIB_TAG=trace): where do the ~16 us of a
latency-bound 31x54 backbone layer go?  Cycles (1.965 GHz) since kernel entry."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aocb200.engine import Engine, T  # noqa: E402
from aocb200.params import synthetic_state_dict  # noqa: E402

NEV = 16
CASES = (("bb.layer3 1024->256 1x1 @31x54", 1, 31, 54, 1024, 256, 1, 0, False),
         ("bb.layer3 256->256 3x3 @31x54", 1, 31, 54, 256, 256, 3, 1, False),
         ("bb.layer3 256->1024 1x1 +res @31x54", 1, 31, 54, 256, 1024, 1, 0, True),
         ("bb.layer4 512->2048 1x1 @31x54", 1, 31, 54, 512, 2048, 1, 0, False),
         ("tiny 16->64 1x1 @31x54", 1, 31, 54, 16, 64, 1, 0, False))


def main():
    dev = torch.device("cuda:0")
    eng = Engine(synthetic_state_dict(1234), dev)
    g = torch.Generator().manual_seed(0)
    for name, N, H, W, Cin, Cout, k, pad, res in CASES:
        x = T(torch.randn(N * H * W * Cin, generator=g).to(dev), N, H, W, Cin)
        w = (torch.randn(Cout, k, k, Cin, generator=g) / (Cin * k * k) ** 0.5).to(dev)
        r = T(torch.randn(N * H * W * Cout, generator=g).to(dev), N, H, W, Cout) if res else None
        eng.w.conv[name] = (w, None, (Cout, k, k, Cin))
        out = eng.conv(x, name, pad=pad, res=r, relu=True)
        for _ in range(3):
            eng.conv(x, name, pad=pad, res=r, relu=True, out=out)
        buf = torch.zeros(NEV * 256, dtype=torch.int64, device=dev)
        eng.L.conv_trace(buf.data_ptr())
        eng.conv(x, name, pad=pad, res=r, relu=True, out=out)
        torch.cuda.synchronize()
        eng.L.conv_trace(None)
        tr = buf.cpu().view(NEV, 256)
        t0 = int(tr[15, 0])
        rel = lambda v: (int(v) - t0) if int(v) else -1
        nz = lambda row: [int(v) for v in tr[row] if int(v)]
        print(name)
        print("  prologue done %d, previous grid complete %d, first activation TMA issued %d, first raw box seen %d, "
              "first MMA issued %d, last MMA issued %d" % (rel(tr[15, 1]), rel(tr[15, 2]), rel(min(nz(0))), rel(min(nz(1))),
                                                           rel(min(nz(7))), rel(max(nz(7)))))
        print("  accumulators drained %d, epilogue / partial tile written %d, arrived %d, all K slices arrived %d, "
              "finished %d" % (rel(tr[15, 4]), rel(max(nz(14))) if nz(14) else -1, rel(tr[15, 5]), rel(tr[15, 6]),
                               rel(tr[15, 7])))


if __name__ == "__main__":
    main()

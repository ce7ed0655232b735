"""What paces the tensor-core convolution?  Times the largest layers with parts of the pipeline switched off
(tooling build only: `AOCB200_BUILD_TAG=trace AOCB200_NVCC_FLAGS=-DAOC_CONV_TRACE python -m aocb200.build`, then
`AOCB200_LIB_TAG=trace python tools/conv_attrib.py`).  Results of the ablated runs are garbage; only the time matters."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aocb200.engine import Engine, T  # noqa: E402
from aocb200.params import synthetic_state_dict  # noqa: E402
from tools.bench_conv import SHAPES  # noqa: E402

CASES = [(0, "full kernel"), (1, "no weight copies"), (2, "no activation TMA"), (3, "no TMA at all"),
         (8, "no transform arithmetic"), (4, "no correction MMAs"), (16, "no main MMAs"), (20, "no MMAs"),
         (11, "no TMA, no transform arithmetic"), (31, "barriers only")]


def main():
    dev = torch.device("cuda:0")
    eng = Engine(synthetic_state_dict(1234), dev)
    g = torch.Generator().manual_seed(0)
    for name, N, H, W, Cin, Cout, k, stride, pad, dil, aff in SHAPES:
        if not any(s in name for s in ("dec.conv1", "dec.aspp", "dec.half 512->128", "bb.dec")):
            continue
        x = T(torch.randn(N * H * W * Cin, generator=g).to(dev), N, H, W, Cin)
        w = (torch.randn(Cout, k, k, Cin, generator=g) / (Cin * k * k) ** 0.5).to(dev)
        eng.w.conv[name] = (w, None, (Cout, k, k, Cin))
        a = (torch.rand(N * Cin, generator=g) + 0.5).to(dev) if aff else None
        b = (torch.randn(N * Cin, generator=g) * 0.1).to(dev) if aff else None
        out = eng.conv(x, name, stride=stride, pad=pad, dil=dil, in_scale=a, in_shift=b, in_relu=aff)
        print(name)
        for bits, what in CASES:
            assert eng.L.set_option(b"conv_dbg", bits) == 0
            reps = 10
            for _ in range(2):
                eng.conv(x, name, stride=stride, pad=pad, dil=dil, in_scale=a, in_shift=b, in_relu=aff, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                eng.conv(x, name, stride=stride, pad=pad, dil=dil, in_scale=a, in_shift=b, in_relu=aff, out=out)
            e1.record()
            torch.cuda.synchronize()
            print("   %-36s %8.1f us" % (what, 1e3 * e0.elapsed_time(e1) / reps))
        eng.L.set_option(b"conv_dbg", 0)


if __name__ == "__main__":
    main()

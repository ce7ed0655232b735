"""Runs one layer shape of tools/bench_conv.py a few times (for `ncu -k regex:conv2 -s 3 -c 1 python tools/one_conv.py <substring>`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aocb200.engine import Engine, T  # noqa: E402
from aocb200.params import synthetic_state_dict  # noqa: E402
from tools.bench_conv import SHAPES  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    eng = Engine(synthetic_state_dict(1234), dev)
    g = torch.Generator().manual_seed(0)
    for name, N, H, W, Cin, Cout, k, stride, pad, dil, aff in SHAPES:
        if sys.argv[1] not in name:
            continue
        x = T(torch.randn(N * H * W * Cin, generator=g).to(dev), N, H, W, Cin)
        w = (torch.randn(Cout, k, k, Cin, generator=g) / (Cin * k * k) ** 0.5).to(dev)
        eng.w.conv[name] = (w, None, (Cout, k, k, Cin))
        a = (torch.rand(N * Cin, generator=g) + 0.5).to(dev) if aff else None
        b = (torch.randn(N * Cin, generator=g) * 0.1).to(dev) if aff else None
        out = None
        for _ in range(5):
            out = eng.conv(x, name, stride=stride, pad=pad, dil=dil, in_scale=a, in_shift=b, in_relu=aff, out=out)
        torch.cuda.synchronize()
        print("ran", name)


if __name__ == "__main__":
    main()

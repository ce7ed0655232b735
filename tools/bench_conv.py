"""CUDA-event timing of the tensor-core convolution on the layer shapes of the 480p / 5-object frame."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aocb200.engine import Engine, T  # noqa: E402
from aocb200.params import synthetic_state_dict  # noqa: E402

SHAPES = [
    ("tiny 16->64 1x1 @31x54 (1 stage, 14 CTAs)", 1, 31, 54, 16, 64, 1, 1, 0, 1, False),
    ("tiny 64->64 1x1 @31x54 (4 stages)", 1, 31, 54, 64, 64, 1, 1, 0, 1, False),
    ("tiny 256->64 1x1 @31x54 (16 stages)", 1, 31, 54, 256, 64, 1, 1, 0, 1, False),
    ("tiny 1024->64 1x1 @31x54 (64 stages)", 1, 31, 54, 1024, 64, 1, 1, 0, 1, False),
    ("tiny 1024->128 1x1 @121x213 (64 stages, 202 CTAs)", 1, 121, 213, 1024, 128, 1, 1, 0, 1, False),

    # name, N, H, W, Cin, Cout, k, stride, pad, dil, affine
    ("dec.conv1 320->128 3x3 @121x213x6", 6, 121, 213, 320, 128, 3, 1, 1, 1, False),
    ("dec.conv2 128->128 3x3 @121x213x6", 6, 121, 213, 128, 128, 3, 1, 1, 1, False),
    ("dec.l1.conv1 164->64 1x1 @full x6", 6, 121, 213, 164, 64, 1, 1, 0, 1, True),
    ("dec.l1.conv2 64->64 3x3 @full x6", 6, 121, 213, 64, 64, 3, 1, 1, 1, True),
    ("dec.l1.conv3 64->256 1x1 @full x6", 6, 121, 213, 64, 256, 1, 1, 0, 1, True),
    ("dec.half 512->128 1x1 @61x107x6", 6, 61, 107, 512, 128, 1, 1, 0, 1, True),
    ("dec.half 128->128 3x3 d2 @61x107x6", 6, 61, 107, 128, 128, 3, 1, 2, 2, True),
    ("dec.half 128->512 1x1 @61x107x6", 6, 61, 107, 128, 512, 1, 1, 0, 1, True),
    ("dec.aspp 512->128 3x3 d12 @half x6", 6, 61, 107, 512, 128, 3, 1, 12, 12, True),
    ("bb.layer1 64->64 3x3 @121x213", 1, 121, 213, 64, 64, 3, 1, 1, 1, False),
    ("bb.layer1 64->256 1x1 @121x213", 1, 121, 213, 64, 256, 1, 1, 0, 1, False),
    ("bb.layer1 256->64 1x1 @121x213", 1, 121, 213, 256, 64, 1, 1, 0, 1, False),
    ("bb.layer2 128->128 3x3 @61x107", 1, 61, 107, 128, 128, 3, 1, 1, 1, False),
    ("bb.layer2 128->512 1x1 @61x107", 1, 61, 107, 128, 512, 1, 1, 0, 1, False),
    ("bb.layer2 512->128 1x1 @61x107", 1, 61, 107, 512, 128, 1, 1, 0, 1, False),
    ("bb.layer4 512->512 3x3 d2 @31x54", 1, 31, 54, 512, 512, 3, 1, 2, 2, False),
    ("bb.layer4 2048->512 1x1 @31x54", 1, 31, 54, 2048, 512, 1, 1, 0, 1, False),
    ("bb.layer4 512->2048 1x1 @31x54", 1, 31, 54, 512, 2048, 1, 1, 0, 1, False),
    ("bb.layer3 256->256 3x3 @31x54", 1, 31, 54, 256, 256, 3, 1, 1, 1, False),
    ("bb.layer3 256->1024 1x1 @31x54", 1, 31, 54, 256, 1024, 1, 1, 0, 1, False),
    ("bb.layer3 1024->256 1x1 @31x54", 1, 31, 54, 1024, 256, 1, 1, 0, 1, False),
    ("bb.aspp 2048->256 3x3 d6 @31x54", 1, 31, 54, 2048, 256, 3, 1, 6, 6, False),
    ("bb.dec 304->256 3x3 @121x213", 1, 121, 213, 304, 256, 3, 1, 1, 1, False),
    ("bb.stem 4->64 7x7 s2 @481x849", 1, 481, 849, 4, 64, 7, 2, 3, 1, False),
    ("bb.stem space-to-depth 16->64 4x4 @242x426", 1, 242, 426, 16, 64, 4, 1, 1, 1, False),
]


def main():
    dev = torch.device("cuda:0")
    eng = Engine(synthetic_state_dict(1234), dev)
    g = torch.Generator().manual_seed(0)
    tot = 0.0
    for name, N, H, W, Cin, Cout, k, stride, pad, dil, aff in SHAPES:
        x = T(torch.randn(N * H * W * Cin, generator=g).to(dev), N, H, W, Cin)
        w = (torch.randn(Cout, k, k, Cin, generator=g) / (Cin * k * k) ** 0.5).to(dev)
        eng.w.conv[name] = (w, None, (Cout, k, k, Cin))
        a = (torch.rand(N * Cin, generator=g) + 0.5).to(dev) if aff else None
        b = (torch.randn(N * Cin, generator=g) * 0.1).to(dev) if aff else None
        out = eng.conv(x, name, stride=stride, pad=pad, dil=dil, in_scale=a, in_shift=b, in_relu=aff)
        for _ in range(3):
            eng.conv(x, name, stride=stride, pad=pad, dil=dil, in_scale=a, in_shift=b, in_relu=aff, out=out)
        # replay from a CUDA graph: the time of back-to-back launches without the host (ctypes + tensor-map encode) in the loop
        reps = 20
        cs = torch.cuda.Stream()
        cs.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cs):
            gr.capture_begin()
            for _ in range(reps):
                eng.conv(x, name, stride=stride, pad=pad, dil=dil, in_scale=a, in_shift=b, in_relu=aff, out=out)
            gr.capture_end()
        torch.cuda.current_stream().wait_stream(cs)
        gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        fl = 2.0 * N * out.H * out.W * Cout * k * k * Cin
        by = 4.0 * (N * H * W * Cin + N * out.H * out.W * Cout)
        if "space-to-depth" not in name:       # the total stays the 29 shapes of the round-1 table
            tot += us
        print("%-40s %9.1f us  %7.1f TFLOP/s (fp32-equivalent; 3 split-fp16 MMAs issued per product)  %7.1f GB/s in+out" %
              (name, us, fl / us / 1e6, by / us / 1e3))
    print("total %.1f us" % tot)


if __name__ == "__main__":
    main()

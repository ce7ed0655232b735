"""Per-stage pipeline timeline of the tensor-core convolution (CTA 0): where do the cycles of a K-loop stage go?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aocb200.engine import Engine, T  # noqa: E402
from aocb200.params import synthetic_state_dict  # noqa: E402

EV = ["tma_issue(raw)", "raw_seen", "alu_done", "op_empty_seen", "published", "mma:b_full", "mma:op_full", "mma:issued",
      "w:slot_free", "corr:ready", "corr:issued", "drain:full", "drain:done", "drain:corr", "epi:done"]
NEV = 16


def main():
    dev = torch.device("cuda:0")
    eng = Engine(synthetic_state_dict(1234), dev)
    g = torch.Generator().manual_seed(0)
    cases = (("bb.layer3 256->256 3x3 @31x54", 1, 31, 54, 256, 256, 3, 1),
             ("dec.conv1 320->128 3x3 @121x213x6", 6, 121, 213, 320, 128, 3, 1),
             ("dec.l1.conv3 64->256 1x1 @121x213x6", 6, 121, 213, 64, 256, 1, 0),
             ("dec.half 128->512 1x1 @61x107x6", 6, 61, 107, 128, 512, 1, 0))
    sel = [a for a in sys.argv[1:] if not a.isdigit()]
    for name, N, H, W, Cin, Cout, k, pad in cases:
        if sel and not any(s in name for s in sel):
            continue
        x = T(torch.randn(N * H * W * Cin, generator=g).to(dev), N, H, W, Cin)
        w = (torch.randn(Cout, k, k, Cin, generator=g) / (Cin * k * k) ** 0.5).to(dev)
        eng.w.conv[name] = (w, None, (Cout, k, k, Cin))
        out = eng.conv(x, name, pad=pad)
        eng.L.set_option(b"conv_dbg", int(os.environ.get("CONV_DBG", "0")))      # ablation bits (tools/conv_attrib.py)
        buf = torch.zeros(NEV * 256, dtype=torch.int64, device=dev)
        eng.L.conv_trace(buf.data_ptr())
        eng.conv(x, name, pad=pad, out=out)
        torch.cuda.synchronize()
        eng.L.conv_trace(None)
        eng.L.set_option(b"conv_dbg", 0)
        tr = buf.cpu().view(NEV, 256)
        t0 = int(tr[0, 0])
        print(name)
        print("stage " + " ".join("%15s" % e for e in EV))
        short = k * k * ((Cin + 15) // 16) < 32          # short K loops: the running stage index covers several tiles
        for s in (list(range(0, 64)) if short else list(range(0, 12)) + list(range(96, 120))):
            print("%5d " % s + " ".join("%15d" % (int(tr[e, s]) - t0 if int(tr[e, s]) else -1) for e in range(len(EV))))
        d = (tr[7, 120] - tr[7, 40]).item() / 80.0
        print("steady state: %.0f cycles per stage (MMA issue to MMA issue, stages 40..120)" % d)


if __name__ == "__main__":
    main()

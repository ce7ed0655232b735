import torch

from aocb200.engine import T


def dev():
    return torch.device("cuda:0")


def to_T(x_nchw, eng, ld=None, off=0):
    """CPU/GPU NCHW tensor -> engine activation (optionally inside a wider buffer to exercise ld/off)."""
    N, C, H, W = x_nchw.shape
    ld = C if ld is None else ld
    buf = torch.full((N * H * W * ld,), 7.25, dtype=torch.float32, device=eng.dev)
    buf.view(N, H, W, ld)[..., off:off + C] = x_nchw.to(eng.dev).permute(0, 2, 3, 1)
    return T(buf, N, H, W, C, ld, off)


def from_T(t):
    v = t.buf.view(t.N, t.H, t.W, t.ld)[..., t.off:t.off + t.C]
    return v.permute(0, 3, 1, 2).contiguous().cpu()


def maxdiff(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item()


def report(name, got, want, tol):
    d = maxdiff(got, want)
    scale = want.abs().max().item()
    print("[parity] %-40s max|d|=%.3e  (ref max %.3e, tol %.1e)" % (name, d, scale, tol))
    assert d <= tol, "%s: max|d|=%.3e > %.1e" % (name, d, tol)

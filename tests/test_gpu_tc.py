"""tcgen05 (3xTF32) kernels against fp64 / the fp32 SIMT kernels of the same library."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import from_T, report, to_T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(state_dict):
    from aocb200.engine import Engine
    return Engine(state_dict, torch.device("cuda:0"))


def test_gemm_tf32x3(eng):
    g = torch.Generator().manual_seed(0)
    M, N, K = 256, 768, 100
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    want = (A.double() @ B.double().t())
    ws = torch.empty(eng.L.tc_image_bytes(M, 104, 128) + eng.L.tc_image_bytes(N, 104, 256), dtype=torch.uint8, device="cuda")
    res = {}
    for variant in (0, 1):
        C = torch.full((M, N), float("nan"), device="cuda")
        eng.L.gemm_tf32x3_test(A.cuda().data_ptr(), B.cuda().data_ptr(), C.data_ptr(), M, N, K, variant, ws.data_ptr(),
                               ws.numel(), eng.stream)
        torch.cuda.synchronize()
        err = (C.cpu().double() - want).abs().max().item()
        res[variant] = err
        print("[parity] tcgen05 3xTF32 gemm variant %d: max|d|=%.3e (|C| max %.2f; fp32 matmul err %.3e)" %
              (variant, err, want.abs().max().item(), (A @ B.t()).double().sub(want).abs().max().item()))
    # the TMEM accumulator truncates: ~0.5 ulp of the running sum per MMA, 39 MMAs for K = 104 -> ~1.2e-6 relative
    # (the matching kernel removes most of it by centring its operands, the convolution by chunked accumulation)
    assert res[0] < 2e-6 * want.abs().max().item(), res
    assert res[1] > 1.0, "the LBO/SBO-swapped descriptor must NOT give the right answer"


def _conv_tc_case(eng, N, H, W, Cin, Cout, k, stride, pad, dil, relu=False, res=False, scale=False, shift=False,
                  in_relu=False, bias=True, ld_in=None, off_in=0, seed=0, chunk=0, check=True):
    """tcgen05 convolution against the float64 evaluation of the same op; the fp32 torch result is the yardstick:
    the kernel must be at least as close to exact arithmetic as 3x the CPU fp32 error (+ 2 ulp of the output range)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) if bias else None
    sc = torch.rand(N, Cin, generator=g) + 0.5 if scale else None
    sh = torch.randn(N, Cin, generator=g) * 0.3 if shift else None

    def ref(dt):
        xin = x.to(dt)
        if scale:
            xin = xin * sc.to(dt)[:, :, None, None]
        if shift:
            xin = xin + sh.to(dt)[:, :, None, None]
        if in_relu:
            xin = F.relu(xin)
        out = F.conv2d(xin, w.to(dt), None if b is None else b.to(dt), stride, pad, dil)
        if res:
            out = out + r.to(dt)
        return F.relu(out) if relu else out

    r = torch.randn(N, Cout, (H + 2 * pad - dil * (k - 1) - 1) // stride + 1,
                    (W + 2 * pad - dil * (k - 1) - 1) // stride + 1, generator=g) if res else None
    want64, want32 = ref(torch.float64), ref(torch.float32)
    name = "tc.%d" % seed
    eng.w.conv[name] = (w.permute(0, 2, 3, 1).contiguous().cuda(), None if b is None else b.cuda(), (Cout, k, k, Cin))
    xt = to_T(x, eng, ld_in, off_in)
    rt = to_T(r, eng) if res else None
    out = to_T(torch.zeros_like(want32), eng, ld=want32.shape[1] + 8, off=4)
    old = eng.conv_chunk
    eng.conv_chunk = chunk
    eng.conv(xt, name, stride=stride, pad=pad, dil=dil, relu=relu, res=rt,
             in_scale=None if sc is None else sc.cuda().contiguous(),
             in_shift=None if sh is None else sh.cuda().contiguous(), in_relu=in_relu, out=out)
    eng.conv_chunk = old
    torch.cuda.synchronize()
    got = from_T(out).double()
    e_tc = (got - want64).abs().max().item()
    e_32 = (want32.double() - want64).abs().max().item()
    rng = want64.abs().max().item()
    print("[parity] conv-tc %dx%d s%d d%d %d->%d M=%d%s%s%s: |tc-fp64|=%.3e  |cpu fp32-fp64|=%.3e  (range %.2f)" %
          (k, k, stride, dil, Cin, Cout, N * want32.shape[2] * want32.shape[3], " scale" if scale else "",
           " shift" if shift else "", " in_relu" if in_relu else "", e_tc, e_32, rng))
    if check:
        assert e_tc <= 3.0 * e_32 + 2.4e-7 * rng, (e_tc, e_32)
    return e_tc, e_32


def test_stem_space_to_depth(eng, state_dict):
    """resnet.py:108-110 -- the 7x7 / stride-2 / pad-3 stem -- evaluated as a 4x4 / stride-1 convolution over the
    space-to-depth frame (aoc_image_to_s2d16_f32 + rearranged weights) against the float64 evaluation of the ORIGINAL
    convolution with the folded weights, odd and even image sizes; the 7x7 form of the same kernel is the yardstick."""
    eng.tc_conv = True
    name = "feature_extracter.backbone.conv1"
    w7, b7, _ = eng.w.conv[name]
    w = w7[..., :3].permute(0, 3, 1, 2).double().cpu()
    for H, W, seed in ((65, 97, 1), (64, 98, 2), (129, 80, 3)):
        img = torch.randn(1, 3, H, W, generator=torch.Generator().manual_seed(seed))
        want = F.relu(F.conv2d(img.double(), w, b7.double().cpu(), 2, 3))
        x4 = eng.new(1, H, W, 4)
        eng.L.image_to_nhwc4_f32(img.cuda().data_ptr(), x4.ptr, H, W, eng.stream)
        y7 = eng.conv(x4, name, stride=2, pad=3, relu=True)
        x16 = eng.new(1, (H + 1) // 2 + 1, (W + 1) // 2 + 1, 16)
        eng.L.image_to_s2d16_f32(img.cuda().data_ptr(), x16.ptr, H, W, eng.stream)
        y4 = eng.conv(x16, name + ".s2d", stride=1, pad=1, relu=True)
        torch.cuda.synchronize()
        assert (y4.H, y4.W) == (y7.H, y7.W) == tuple(want.shape[2:])
        e4 = (from_T(y4).double() - want).abs().max().item()
        e7 = (from_T(y7).double() - want).abs().max().item()
        print("[parity] stem %dx%d: |s2d-fp64|=%.3e  |7x7-fp64|=%.3e (range %.2f)" % (H, W, e4, e7, want.abs().max().item()))
        assert e4 <= 2.0 * e7 + 2.4e-7 * want.abs().max().item(), (e4, e7)


def test_conv_tc_vs_fp64(eng):
    eng.tc_conv = True
    c = _conv_tc_case
    c(eng, 1, 33, 41, 4, 64, 7, 2, 3, 1, relu=True, seed=101)                       # stem (Cin padded 3 -> 4)
    c(eng, 1, 17, 23, 64, 256, 1, 1, 0, 1, relu=True, res=True, seed=102)           # 1x1 + residual
    c(eng, 1, 17, 23, 128, 128, 3, 2, 1, 1, relu=True, seed=103)                    # stride 2
    c(eng, 1, 9, 13, 512, 512, 3, 1, 4, 4, relu=True, seed=104)                     # dilated
    c(eng, 3, 19, 21, 164, 64, 1, 1, 0, 1, scale=True, bias=False, seed=105)        # gate, Cin % 16 != 0
    c(eng, 2, 19, 21, 256, 100, 1, 1, 0, 1, seed=106)                               # Cout = 100
    c(eng, 2, 13, 17, 24, 64, 1, 1, 0, 1, seed=107)                                 # prehead
    c(eng, 6, 61, 107, 320, 128, 3, 1, 1, 1, bias=False, seed=108)                  # decoder conv1 shape
    c(eng, 2, 1, 1, 512, 128, 1, 1, 0, 1, relu=True, bias=False, seed=109)          # 1x1 spatial (pooled branch)
    c(eng, 2, 12, 14, 48, 64, 3, 1, 6, 6, scale=True, bias=False, ld_in=80, off_in=16, seed=110)   # channel slice
    c(eng, 1, 31, 54, 2048, 256, 3, 1, 6, 6, relu=True, seed=111)                   # ASPP: K = 18432
    c(eng, 6, 61, 107, 256, 512, 1, 1, 0, 1, bias=False, seed=112)
    c(eng, 1, 31, 54, 1024, 2048, 1, 2, 0, 1, seed=113)                             # 1x1 stride-2 downsample
    c(eng, 1, 121, 213, 304, 256, 3, 1, 1, 1, relu=True, seed=114)                  # DeepLab decoder
    # fused GroupNorm-apply (+ReLU) in front of the convolution: zero padding must stay zero after the shift
    c(eng, 3, 19, 21, 64, 64, 3, 1, 2, 2, scale=True, shift=True, in_relu=True, bias=False, seed=115)
    c(eng, 2, 25, 33, 128, 512, 1, 1, 0, 1, scale=True, shift=True, in_relu=True, bias=False, seed=116)
    c(eng, 2, 25, 33, 128, 128, 3, 2, 1, 1, scale=True, shift=True, in_relu=True, bias=False, seed=117)
    c(eng, 2, 25, 33, 100, 64, 3, 1, 1, 1, in_relu=True, seed=118)


def test_conv_tc_splitk(eng):
    """layers with few output tiles (the 31x54 backbone maps) are split along K over the SMs; partial sums are added in
    a fixed order by a second kernel that also applies bias / residual / ReLU.  Both schedules must meet the bar."""
    eng.tc_conv = True
    c = _conv_tc_case
    for on in (1, 0):
        assert eng.L.set_option(b"conv_splitk", on) == 0
        c(eng, 1, 31, 54, 256, 256, 3, 1, 1, 1, relu=True, seed=121)                    # layer3 conv2
        c(eng, 1, 31, 54, 1024, 256, 1, 1, 0, 1, relu=True, seed=122)                   # layer3 conv1
        c(eng, 1, 31, 54, 512, 512, 3, 1, 2, 2, relu=True, res=True, seed=123)          # layer4, residual
        c(eng, 1, 31, 54, 2048, 256, 3, 1, 18, 18, relu=True, seed=124)                 # ASPP d18
        c(eng, 1, 5, 7, 304, 100, 3, 1, 1, 1, seed=125)                                 # one tile, odd 16-channel tail
        c(eng, 2, 1, 1, 2048, 256, 1, 1, 0, 1, relu=True, seed=126)                     # image-pool branch
    eng.L.set_option(b"conv_splitk", 1)


def test_conv_tc_tf32_mode(eng):
    """the 3xTF32 operand mode stays in the library for weights / activations beyond the fp16 range (the engine selects
    it when a folded weight exceeds 6e4): same bar as the default split-fp16 mode, including values fp16 cannot hold"""
    eng.tc_conv = True
    c = _conv_tc_case
    try:
        eng.conv_mode = 0                                       # AOC_CONV_TF32X3: a per-call argument, recorded per weight image
        c(eng, 1, 31, 54, 256, 256, 3, 1, 1, 1, relu=True, seed=131)
        c(eng, 2, 25, 33, 128, 512, 1, 1, 0, 1, scale=True, shift=True, in_relu=True, bias=False, seed=132)
        c(eng, 1, 121, 213, 64, 64, 3, 1, 1, 1, relu=True, res=True, seed=133)
        c(eng, 1, 31, 54, 2048, 256, 3, 1, 12, 12, relu=True, seed=134)
    finally:
        eng.conv_mode = 1


def test_conv_overflow_flag(eng):
    """The split-fp16 operand mode clamps at 65504.  A convolution whose input reaches 1e5 must raise the sticky device
    word (and give a wrong result); the same layer in 3xTF32 mode must not raise it and must be right."""
    eng.tc_conv = True
    g = torch.Generator().manual_seed(7)
    N, H, W, Cin, Cout = 1, 20, 30, 64, 64
    x = torch.randn(N, Cin, H, W, generator=g)
    x[0, 5, 7, 11] = 1.0e5
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    want = F.conv2d(x.double(), w.double(), None, 1, 1, 1)
    eng.w.conv["tc.ovf"] = (w.permute(0, 2, 3, 1).contiguous().cuda(), None, (Cout, 3, 3, Cin))
    res = {}
    try:
        for mode in (1, 0):
            eng.conv_mode = mode
            eng._ovf.zero_()
            out = eng.conv(to_T(x, eng), "tc.ovf", pad=1)
            torch.cuda.synchronize()
            res[mode] = (int(eng._ovf.item()), (from_T(out).double() - want).abs().max().item())
    finally:
        eng.conv_mode = 1
        eng._ovf.zero_()
    print("[parity] conv with a 1e5 activation: split-fp16 flag %d err %.3e | 3xTF32 flag %d err %.3e" % (res[1] + res[0]))
    assert res[1][0] == 1 and res[1][1] > 1.0, res      # clamped at 65504: off by ~3.4e4 * |w|
    assert res[0][0] == 0 and res[0][1] < 1e-2, res     # 1e5 * 1e-7 relative
    # below the guard threshold nothing is raised
    x[0, 5, 7, 11] = 3.0e4
    out = eng.conv(to_T(x, eng), "tc.ovf", pad=1)
    torch.cuda.synchronize()
    assert int(eng._ovf.item()) == 0


def test_conv_tc_chunking(eng):
    """the truncating TMEM accumulation: a single chain over K = 18432 is visibly biased, short chains are not"""
    eng.tc_conv = True
    e_long, _ = _conv_tc_case(eng, 1, 31, 54, 2048, 256, 3, 1, 6, 6, relu=True, seed=111, chunk=1 << 20, check=False)
    e_short, _ = _conv_tc_case(eng, 1, 31, 54, 2048, 256, 3, 1, 6, 6, relu=True, seed=111, chunk=8)
    print("[parity] conv-tc chunking: single chain %.3e vs 8-stage chains %.3e" % (e_long, e_short))
    assert e_short < 0.2 * e_long


@pytest.mark.parametrize("f16", [1, 0], ids=["split-fp16", "3xTF32"])
def test_global_match_tc_vs_simt(eng, f16):
    from test_gpu_ops import _rand_scene
    eng.L.set_option(b"match_f16", f16)
    for seed, h, w, K, F_, absent in ((1, 25, 33, 3, 2, None), (2, 61, 107, 5, 2, 4), (3, 33, 37, 1, 3, None)):
        embs, masks, _ = _rand_scene(seed, h, w, K, F_, absent)
        outs = {}
        for tc in (False, True):
            eng.tc_match = tc
            eng.bank.reset()
            eng.keep_debug = True
            np.random.seed(seed)
            eng.match_features([e.cuda() for e in embs[:F_]], [m.cuda() for m in masks[:F_]], embs[F_].cuda(),
                               masks[F_].cuda(), to_T(embs[F_ + 1], eng), K)
            torch.cuda.synchronize()
            outs[tc] = eng.debug["g"].clone().cpu()
        eng.keep_debug = False
        eng.tc_match = True
        report("global match tcgen05 vs simt (seed %d)" % seed, outs[True], outs[False], 5e-6)
    eng.L.set_option(b"match_f16", 1)


def test_conv_tc_halo_variant(eng):
    """The halo variant of the convolution (3x3, stride 1, pad = dilation = 1, at least one 16 x 8 tile per SM: the halo
    patch of a channel box is transformed once and the nine taps read it through shifted shared-memory descriptors) against
    float64, on shapes that exercise: an odd number of 16-channel stages (half a channel box), ragged borders in both
    directions, Cout = 64 / 96 / 256, the fused input affine with shift (zero padding must stay zero) + ReLU, residual, bias,
    output ReLU and the statistics rows; and against the per-tap kernel on the same inputs."""
    eng.tc_conv = True
    c = _conv_tc_case
    L = eng.L
    assert L.set_option(b"conv_halo", 1) == 0
    cases = [dict(N=3, H=70, W=90, Cin=48, Cout=96, seed=201, relu=True),                     # ncc = 3 (odd), Cout % 128 != 0
             dict(N=6, H=61, W=107, Cin=64, Cout=64, seed=202, scale=True, shift=True, in_relu=True, bias=False),
             dict(N=1, H=121, W=213, Cin=128, Cout=256, seed=203, res=True, relu=True),
             dict(N=4, H=50, W=75, Cin=320, Cout=128, seed=204, bias=False)]
    cases.append(dict(N=6, H=61, W=107, Cin=128, Cout=128, seed=205, d=2, scale=True, shift=True, in_relu=True, bias=False))   # dilation 2
    for kw in cases:
        kw = dict(kw)
        N, H, W, Cin, Cout = (kw.pop(k) for k in ("N", "H", "W", "Cin", "Cout"))
        d = kw.pop("d", 1)
        assert L.conv_tiles_per_image(N, H, W, Cin, Cout, 3, 3, 1, d, d, 1) == 4 * ((W + 7) // 8) * ((H + 15) // 16), "halo not selected"
        e_h, e32 = c(eng, N, H, W, Cin, Cout, 3, 1, d, d, **kw)
        L.set_option(b"conv_halo", 0)
        try:
            e_t, _ = c(eng, N, H, W, Cin, Cout, 3, 1, d, d, **kw)
        finally:
            L.set_option(b"conv_halo", 1)
        print("[parity]    halo %.3e   per-tap %.3e   cpu fp32 %.3e" % (e_h, e_t, e32))


def test_conv_tc_tail_split(eng):
    """Tail splitting (the tiles of a partial last wave are cut into K slices; slice 0 collects the parked accumulators and
    runs the normal epilogue): results against float64 and against the unsplit schedule, and the epilogue statistics rows
    (per-channel sum / sum of squares) against the unsplit schedule -- on the two layer types it serves in the 480p frame."""
    eng.tc_conv = True
    L = eng.L
    c = _conv_tc_case
    L.set_option(b"conv_tail_min_stages", 0)               # the launcher only takes it for K loops of >= 192 stages: test it on all
    try:
        _tail_split_cases(eng, L, c)
    finally:
        L.set_option(b"conv_tail_min_stages", 192)


def _tail_split_cases(eng, L, c):
    for kw in (dict(N=6, H=61, W=107, Cin=512, Cout=128, k=1, pad=0, dil=1, seed=301, scale=True, shift=True, in_relu=True, bias=False),
               dict(N=6, H=61, W=107, Cin=128, Cout=128, k=3, pad=12, dil=12, seed=302, scale=True, bias=False),
               dict(N=1, H=121, W=213, Cin=256, Cout=64, k=1, pad=0, dil=1, seed=303, relu=True, res=True)):
        kw = dict(kw)
        N, H, W, Cin, Cout, k, pad, dil = (kw.pop(x) for x in ("N", "H", "W", "Cin", "Cout", "k", "pad", "dil"))
        L.set_option(b"conv_tail", 1)
        e_t, e32 = c(eng, N, H, W, Cin, Cout, k, 1, pad, dil, **kw)
        L.set_option(b"conv_tail", 0)
        try:
            e_u, _ = c(eng, N, H, W, Cin, Cout, k, 1, pad, dil, **kw)
        finally:
            L.set_option(b"conv_tail", 1)
        print("[parity]    tail-split %.3e   unsplit %.3e   cpu fp32 %.3e" % (e_t, e_u, e32))
    # statistics rows through the engine's own path
    g = torch.Generator().manual_seed(7)
    x = torch.randn(6, 128, 61, 107, generator=g)
    w = torch.randn(128, 128, 3, 3, generator=g) / (128 * 9) ** 0.5
    eng.w.conv["tc.tail"] = (w.permute(0, 2, 3, 1).contiguous().cuda(), None, (128, 3, 3, 128))
    xt = to_T(x, eng)
    res = []
    for on in (1, 0):
        L.set_option(b"conv_tail", on)
        y, st = eng.conv(xt, "tc.tail", pad=6, dil=6, stats=True)
        res.append((from_T(y).clone(), eng.dense_stats(st).clone().cpu()))
    L.set_option(b"conv_tail", 1)
    dy = (res[0][0] - res[1][0]).abs().max().item()
    ds = ((res[0][1] - res[1][1]).abs() / (res[1][1].abs() + 1.0)).max().item()
    want = torch.stack([from_T(res[1][0] if False else y).double().sum((2, 3)), (from_T(y).double() ** 2).sum((2, 3))], 1).reshape(-1)
    dw = ((res[0][1] - want).abs() / (want.abs() + 1.0)).max().item()
    print("[parity] tail-split vs unsplit: outputs %.3e, statistics (relative) %.3e; statistics vs recomputed %.3e" % (dy, ds, dw))
    assert dy < 5e-6 and ds < 1e-5 and dw < 1e-5

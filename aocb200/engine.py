"""Host side of the B200 engine: weight repacking and the per-frame kernel schedule.

Everything numerical runs in libaocb200.so (hand-written sm_100a CUDA, C ABI in include/aocb200.h); torch is used
for device memory, the current stream and a few index/dtype conversions on label maps.  The schedule mirrors
AOCNet.forward_for_eval / before_seghead_process (networks/aoc/aocnet.py:84-372) and CalibrationDecoding.forward
(networks/aoc/decoding_module.py:96-225) of the reference, with SURVEY.md Appendix A's repair set.
"""
import os

import numpy as np
import torch

from .lib import lib
from .params import EMB, HEAD

MAXO = 16
PROXY_SLOTS = 36        # proxy slots per object at cluster_num <= 16 (2 * kmax + 4; see Engine.kmax)
KMEANS_MAX_K = 64
META_INTS = 2 * MAXO + 3
BANK_ALIGN = 256


class T:
    """NHWC fp32 activation: a channel slice [off, off+C) of rows of width ld inside `buf`."""
    __slots__ = ("buf", "N", "H", "W", "C", "ld", "off")

    def __init__(self, buf, N, H, W, C, ld=None, off=0):
        self.buf, self.N, self.H, self.W, self.C = buf, N, H, W, C
        self.ld = C if ld is None else ld
        self.off = off

    @property
    def ptr(self):
        return self.buf.data_ptr() + 4 * self.off

    @property
    def HW(self):
        return self.H * self.W

    def slice(self, off, C):
        return T(self.buf, self.N, self.H, self.W, C, self.ld, self.off + off)

    def nchw(self):
        """torch view [N,C,H,W] (channels-last strides) of a full-width activation."""
        assert self.off == 0 and self.ld == self.C
        return self.buf.view(self.N, self.H, self.W, self.C).permute(0, 3, 1, 2)


def _p(t):
    return None if t is None else t.data_ptr()


class TileStats:
    """per-tile partial statistics written by a convolution epilogue ([N * tpi][2][C] floats); reduced on demand"""
    __slots__ = ("ts", "tpi", "N", "C", "_dense")

    def __init__(self, ts, tpi, N, C):
        self.ts, self.tpi, self.N, self.C, self._dense = ts, tpi, N, C, None


class Weights:
    """Device-resident, repacked parameters (done once per state_dict).  The folding arithmetic (FrozenBatchNorm into the
    convolutions, the constant branches of the conditioning blocks) runs on the HOST in fp32 and each result is uploaded
    once: no torch kernel is launched on the device, the first kernels of a fresh engine are the library's own."""

    def __init__(self, sd, device):
        f = lambda k: sd[k].detach().to(device="cpu", dtype=torch.float32)
        self.dev = device
        self.conv = {}      # name -> (w [Cout,kh,kw,Cin], bias|None, (Cout,kh,kw,Cin))
        self.vec = {}       # misc vectors / matrices
        has = lambda k: k in sd

        def add_conv(name, bn=None, pad_cin=None):
            w = f(name + ".weight")
            b = f(name + ".bias") if has(name + ".bias") else None
            if bn is not None:  # FrozenBatchNorm2d (normalization.py:17-23) folded into the conv
                scale = f(bn + ".weight") * (f(bn + ".running_var") + 1e-5).rsqrt()
                shift = f(bn + ".bias") - f(bn + ".running_mean") * scale
                w = w * scale.view(-1, 1, 1, 1)
                b = shift if b is None else b * scale + shift
            if pad_cin is not None and w.shape[1] < pad_cin:
                w = torch.cat([w, w.new_zeros(w.shape[0], pad_cin - w.shape[1], w.shape[2], w.shape[3])], 1)
            w = w.permute(0, 2, 3, 1).contiguous()
            self.conv[name] = (w, None if b is None else b.contiguous(), tuple(w.shape))

        bb = "feature_extracter.backbone"
        add_conv(bb + ".conv1", bb + ".bn1", pad_cin=4)
        # the stem as a 4x4 / stride-1 convolution over the space-to-depth frame (aoc_image_to_s2d16_f32): tap a' of the 4
        # and source-row parity py cover source tap r = 2 a' + py - 1 of the 7 (r = -1: no such tap, zero weights)
        w7 = self.conv[bb + ".conv1"][0]                                  # [64][7][7][4], BatchNorm folded
        w4 = w7.new_zeros(w7.shape[0], 4, 4, 16)
        for a_ in range(4):
            for py in (0, 1):
                r = 2 * a_ + py - 1
                for b_ in range(4):
                    for px in (0, 1):
                        s_ = 2 * b_ + px - 1
                        if 0 <= r <= 6 and 0 <= s_ <= 6:
                            w4[:, a_, b_, (py * 2 + px) * 4:(py * 2 + px) * 4 + 4] = w7[:, r, s_, :]
        self.conv[bb + ".conv1.s2d"] = (w4.contiguous(), self.conv[bb + ".conv1"][1], tuple(w4.shape))
        for lname, blocks in (("layer1", 3), ("layer2", 4), ("layer3", 23), ("layer4", 3)):
            for i in range(blocks):
                p = "%s.%s.%d" % (bb, lname, i)
                for j in (1, 2, 3):
                    add_conv("%s.conv%d" % (p, j), "%s.bn%d" % (p, j))
                if has(p + ".downsample.0.weight"):
                    add_conv(p + ".downsample.0", p + ".downsample.1")
        a = "feature_extracter.aspp"
        for i in (1, 2, 3, 4):
            add_conv("%s.aspp%d.atrous_conv" % (a, i), "%s.aspp%d.bn" % (a, i))
        add_conv(a + ".global_avg_pool.1", a + ".global_avg_pool.2")
        add_conv(a + ".conv1", a + ".bn1")
        d = "feature_extracter.decoder"
        add_conv(d + ".conv1", d + ".bn1")
        add_conv(d + ".last_conv.0", d + ".last_conv.1")
        add_conv(d + ".last_conv.4", d + ".last_conv.5")
        add_conv("embedding_conv")
        add_conv("dynamic_prehead.conv")
        self.vec["seperate_conv.weight"] = f("seperate_conv.weight").reshape(256, 9).contiguous()
        self.vec["seperate_conv.bias"] = f("seperate_conv.bias").contiguous()
        h = "dynamic_seghead"
        for k in sd.keys():
            if not k.startswith(h + "."):
                continue
            if k.endswith(".weight") and sd[k].dim() == 4 and "phi_layer" not in k:
                add_conv(k[:-len(".weight")])
        # everything else as flat fp32 tensors
        for k in sd.keys():
            if k.startswith("feature_extracter."):
                continue
            base = k.rsplit(".", 1)[0]
            if base in self.conv:
                continue
            self.vec[k] = f(k).reshape(-1).contiguous() if sd[k].dim() != 2 else f(k).contiguous()
        self.vec["dis_bias"] = torch.cat([f("bg_bias").reshape(1), f("fg_bias").reshape(1).expand(MAXO - 1)]).contiguous()
        # conditioning_block: CL_2 / CL_3 are input independent under repair R7 (they return mlp_layer.bias), so
        # mlp_layer([c1, c2, c3]) = W[:, :C] c1 + (W[:, C:] [c2; c3] + b)      (conditioning_layer.py:63-86)
        for blk, C in (("CLB2", 256), ("CLB3", 256), ("CLB4", 512), ("CLB5", 512)):
            p = "%s.%s" % (h, blk)
            Wm, bm = f(p + ".mlp_layer.weight"), f(p + ".mlp_layer.bias")
            c23 = torch.cat([f(p + ".CL_2.mlp_layer.bias"), f(p + ".CL_3.mlp_layer.bias")])
            self.vec[p + ".fold.weight"] = Wm[:, :C].contiguous()
            self.vec[p + ".fold.bias"] = (bm + Wm[:, C:] @ c23).contiguous()
        # largest |weight| a convolution will see (range check of the split-fp16 operand mode), taken on the host
        self.conv_wmax = max(float(w.abs().max()) for w, _, _ in self.conv.values())
        up = lambda t: None if t is None else t.contiguous().to(device)
        self.conv = {k: (up(w), up(b), shp) for k, (w, b, shp) in self.conv.items()}
        self.vec = {k: up(v) for k, v in self.vec.items()}


class Bank:
    def __init__(self):
        self.version = 0          # bumped whenever frames are appended or the bank is reset; NEVER reused: captured
        self.reset()              # graphs of the bank-dependent segment are keyed on it (a new sequence's one-frame
                                  # bank must not match the previous sequence's one-frame bank)

    def reset(self):
        self.refs, self.masks = [], []
        self.emb_all = self.ids_all = None
        self.hw = None
        self.n = 0
        self.version += 1
        self.index = None         # object-sorted index of the current version (see Engine._bank_index)


class _Graph:
    """a captured segment; replay() keeps the library's launch counter honest (bench.py `gpu_launches`)"""
    __slots__ = ("g", "n", "L")

    def __init__(self, g, n, L):
        self.g, self.n, self.L = g, n, L

    def replay(self):
        self.g.replay()
        self.L.launches += self.n
        if self.L.profile is not None:
            self.L.replayed.add(id(self))       # bench.py's in-graph kernel timing reads the events of replayed graphs


class Engine:
    def __init__(self, state_dict, device):
        self.dev = torch.device(device)
        self.L = lib()
        self.L.check_device(self.dev.index if self.dev.index is not None else torch.cuda.current_device())
        self.w = Weights(state_dict, self.dev)
        self.bank = Bank()
        self.kmeans_iters = 20
        self.cluster_num = 16           # matching.py:507; up to KMEANS_MAX_K (the kernels run at width 16 or 64)
        self.debug = {}
        self.keep_debug = False
        self.force_proxies = None       # test hook, see _apply_forced_proxies
        self.shard = None               # bank sharding over several GPUs (aocb200/shard.py::setup_bank_sharding)
        # tensor-core (tcgen05, 3xTF32) kernels vs the fp32 SIMT kernels; both are CUDA, same results to ~1e-6 relative
        tc = os.environ.get("AOCB200_TC", "1") != "0"
        self.tc_conv = tc and os.environ.get("AOCB200_TC_CONV", "1") != "0"
        self.tc_match = tc and os.environ.get("AOCB200_TC_MATCH", "1") != "0"
        self._wpacked = {}
        for kv in filter(None, os.environ.get("AOCB200_OPTS", "").split(",")):     # e.g. "conv_pdl=0,conv_splitk=0"
            k, v = kv.split("=")
            self.L.set_option(k.strip().encode(), int(v))
        # Convolution operand mode of THIS engine (a per-call argument of the library, recorded with every packed weight
        # image).  Split-fp16 operands saturate at the fp16 range: weights (FrozenBatchNorm folded in) that do not fit
        # select 3xTF32 from the start (same kernels and tests, ~12 % slower); activations are watched by the kernels
        # themselves (sticky device word `_ovf`, see _overflow_check).
        self.conv_mode = 1 if self.w.conv_wmax < 6.0e4 else 0       # AOC_CONV_SPLIT_F16 / AOC_CONV_TF32X3
        if os.environ.get("AOCB200_CONV_MODE", "") in ("tf32", "0"):
            self.conv_mode = 0
        self._ovf = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._ovf_ring = [(torch.zeros(1, dtype=torch.int32).pin_memory(), torch.cuda.Event(), [None]) for _ in range(4)]
        self._ovf_turn = 0
        self.overflow_policy = os.environ.get("AOCB200_OVERFLOW", "deferred")    # deferred | sync | off
        self.overflow_frames = []       # frame counters whose fp16 operands overflowed (results invalid / re-run)
        self.frame_no = 0
        self.conv_chunk = int(os.environ.get("AOCB200_CONV_CHUNK", "0"))   # 0 = library default
        self._meta_host = torch.empty(META_INTS, dtype=torch.int32).pin_memory()
        self.use_graphs = os.environ.get("AOCB200_GRAPHS", "1") != "0"
        # segment F (bank-dependent: 15 kernels) as a graph that is re-captured when the bank grows, or as plain launches
        # (default: plain launches -- 15 kernels, ~40 us of host time per frame, nothing to re-capture at a bank change; the
        # re-capture cost 0.3 ms per change and, sporadically, a 50-120 ms stall of the instantiation)
        self.segf_graph = os.environ.get("AOCB200_SEGF_GRAPH", "0") != "0"
        self.stem_s2d = os.environ.get("AOCB200_STEM_S2D", "1") != "0"      # space-to-depth stem (0: the 7x7 / stride-2 form)
        self._segA, self._static = {}, {}
        self._cap_stream = self._pool = None
        self._gt_cache = (None, 0, 0)
        self._ws_keep = []
        self._ws = {}
        # eval-loop label bookkeeping on the device: bit o of the word = label o was seen in a
        # ground-truth frame (eval_manager_mm.py:252-270); entropy threshold of the confident mask (:339-349)
        self._exist = torch.full((1,), -1, dtype=torch.int32, device=self.dev)
        self._exist_bits = -1
        self.unc_ratio = 1.0
        self.want_probs = True          # DeviceSequence switches the [1,O,H,W] probability output off (nobody reads it)
        self.last_logits = self.last_label = self.last_conf_label = None

    # ------------------------------------------------------------------ plumbing
    @property
    def kmax(self):
        """compile-time width of the k-means / proxy kernels for the current cluster_num"""
        assert 1 <= self.cluster_num <= KMEANS_MAX_K, "cluster_num must be in 1..%d" % KMEANS_MAX_K
        return 16 if self.cluster_num <= 16 else KMEANS_MAX_K

    @property
    def proxy_slots(self):
        return 2 * self.kmax + 4

    @property
    def stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def empty(self, n, dtype=torch.float32):
        return torch.empty(int(n), dtype=dtype, device=self.dev)

    def zeros(self, n, dtype=torch.float32):
        """zero-filled 32-bit buffer (library fill kernel: no torch kernel inside the captured frame)"""
        t = torch.empty(int(n), dtype=dtype, device=self.dev)
        assert t.element_size() == 4
        self.L.fill_u32(t.data_ptr(), 0, int(n), self.stream)
        return t

    def vec_op(self, a, b, op):
        out = self.empty(a.numel())
        self.L.vec_op_f32(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), op, self.stream)
        return out

    def new(self, N, H, W, C):
        return T(self.empty(N * H * W * C), N, H, W, C)

    def grow(self, name, n, dtype):
        """persistent typed buffer of at least n elements, grown geometrically (bank-sized arrays: object-sorted rows, row
        maps); contents are not preserved"""
        b = self._ws.get(name)
        if b is None or b.numel() < n or b.dtype != dtype:
            if b is not None:
                self._ws_keep.append(b)
                n = max(int(n), 2 * b.numel())
            b = torch.empty(int(n), dtype=dtype, device=self.dev)
            self._ws[name] = b
        return b

    def ws(self, name, nbytes, zero_head=0):
        """persistent scratch (never handed to the caller); zero_head: bytes at the front that the library expects zeroed
        when it first sees the buffer (split-K arrival counters) -- cleared once, at allocation"""
        b = self._ws.get(name)
        if b is None or b.numel() < nbytes:
            if b is not None:
                self._ws_keep.append(b)      # a captured graph may still point at the superseded buffer
                nbytes = max(int(nbytes), 2 * b.numel())     # geometric growth: a growing bank must not pay a cudaMalloc
            b = torch.empty(int(nbytes), dtype=torch.uint8, device=self.dev)   # (tens of ms, device-synchronising) per step
            if zero_head:
                self.L.fill_u32(b.data_ptr(), 0, zero_head // 4, self.stream)
            self._ws[name] = b
        return b

    # ------------------------------------------------------------------ layer helpers
    def conv(self, x, name, stride=1, pad=0, dil=1, relu=False, res=None, in_scale=None, out=None, in_shift=None,
             in_relu=False, stats=False):
        """y = conv(relu?(x * in_scale[n,c] + in_shift[n,c])) + bias (+ res) (ReLU).  stats=True also returns the
        per-(sample, channel) sum / sum of squares of y ([N][2][C] doubles), taken in the convolution epilogue."""
        w, b, (Cout, kh, kw, Cin) = self.w.conv[name]
        assert Cin == x.C, (name, Cin, x.C)
        Ho = (x.H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        Wo = (x.W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        if out is None:
            out = self.new(x.N, Ho, Wo, Cout)
        assert out.C == Cout and out.H == Ho and out.W == Wo and out.N == x.N
        if self.tc_conv:
            mode = self.conv_mode
            wp = self._wpacked.get(name)
            if wp is None or wp[0] is not w or wp[2] != mode:
                buf = torch.empty(self.L.conv_packed_weight_bytes(Cout, Cin, kh, kw, mode), dtype=torch.uint8,
                                  device=self.dev)
                self.L.conv_pack_weights(w.data_ptr(), Cout, Cin, kh, kw, mode, buf.data_ptr(), self.stream)
                # the convolution's weight-copy warp does not wait for the preceding kernel of the stream (programmatic
                # dependent launch: weights are constants, their first tiles are fetched under the previous layer's tail),
                # so a packed image must be COMPLETE before its first use -- once per layer, on the warm-up pass
                assert not torch.cuda.is_current_stream_capturing(), "weights are packed on the warm-up pass, not in a capture"
                torch.cuda.synchronize(self.dev)
                wp = (w, buf, mode)
                self._wpacked[name] = wp
            ts = None
            if stats:
                tpi = self.L.conv_tiles_per_image(x.N, x.H, x.W, Cin, Cout, kh, kw, stride, pad, dil, mode)
                ts = self.empty(x.N * tpi * 2 * Cout)
            # scratch for the library's K splits: whole-layer split-K of the small maps (few pixel tiles) and tail splitting of
            # a partial last wave (one parked accumulator tile per SM); its 4 KB header of arrival counters stays zero
            wsn = self.L.conv_workspace_bytes(x.N, x.H, x.W, Cout, kh, kw, stride, pad, dil)
            wsp = self.ws("conv_splitk", wsn, zero_head=4096).data_ptr()
            self.L.conv2d_nhwc_tc(x.ptr, wp[1].data_ptr(), _p(b), None if res is None else res.ptr, _p(in_scale),
                                  _p(in_shift), 1 if in_relu else 0, out.ptr, _p(ts), x.N, x.H, x.W, Cin, x.ld, Cout,
                                  out.ld, 0 if res is None else res.ld, kh, kw, stride, pad, dil, 1 if relu else 0,
                                  self.conv_chunk, mode, self._ovf.data_ptr(), wsp, wsn, self.stream)
            return (out, TileStats(ts, tpi, x.N, Cout)) if stats else out
        assert in_shift is None and not in_relu, "the fp32 SIMT convolution only fuses an input scale"
        self.L.conv2d_nhwc_f32(x.ptr, w.data_ptr(), _p(b), None if res is None else res.ptr, _p(in_scale), out.ptr,
                               x.N, x.H, x.W, Cin, x.ld, Cout, out.ld, 0 if res is None else res.ld, kh, kw, stride,
                               pad, dil, 1 if relu else 0, self.stream)
        return (out, self.stats(out)) if stats else out

    def dense_stats(self, st):
        """[N][2][C] double statistics from whatever a producer handed over"""
        if not isinstance(st, TileStats):
            return st
        if st._dense is None:
            st._dense = self.empty(st.N * 2 * st.C, torch.float64)
            self.L.tile_stats_reduce_f32(st.ts.data_ptr(), st.N, st.tpi, st.C, st._dense.data_ptr(), self.stream)
        return st._dense

    def stats(self, x, phi=None, thr=None):
        L = self.L
        n = L.channel_stats_workspace_bytes(x.N, x.HW, x.C)
        ws = self.ws("stats", n)
        st = self.empty(x.N * 2 * x.C, torch.float64)
        L.channel_stats_f32(x.ptr, x.N, x.HW, x.C, x.ld, _p(phi), _p(thr), st.data_ptr(), ws.data_ptr(), n, self.stream)
        return st

    def affine(self, x, a, b=None, res=None, res_scale=None, relu=False, out=None, want_stats=False):
        """y = x*a[n,c] + b[n,c] (+ res*res_scale[n,c]) (ReLU); want_stats -> (y, [N][2][C] statistics of y, same pass)"""
        if out is None:
            out = self.new(x.N, x.H, x.W, x.C)
        if want_stats:
            n = self.L.channel_stats_workspace_bytes(x.N, x.HW, x.C)
            ws = self.ws("stats", n)
            st = self.empty(x.N * 2 * x.C, torch.float64)
            self.L.affine_stats_nc_f32(x.ptr, a.data_ptr(), _p(b), None if res is None else res.ptr, _p(res_scale),
                                       out.ptr, x.N, x.HW, x.C, x.ld, out.ld, 0 if res is None else res.ld,
                                       1 if relu else 0, st.data_ptr(), ws.data_ptr(), n, self.stream)
            return out, st
        self.L.affine_nc_f32(x.ptr, a.data_ptr(), _p(b), None if res is None else res.ptr, _p(res_scale), out.ptr,
                             x.N, x.HW, x.C, x.ld, out.ld, 0 if res is None else res.ld, 1 if relu else 0, self.stream)
        return out

    def gn_ab(self, x, name, groups, st=None):
        """nn.GroupNorm(groups, C) of x as per-(sample, channel) coefficients: y = x * a + b"""
        a, b = self.empty(x.N * x.C), self.empty(x.N * x.C)
        if isinstance(st, TileStats) and x.C // groups <= 32:
            self.L.gn_coeffs_tiles_f32(st.ts.data_ptr(), st.tpi, self.w.vec[name + ".weight"].data_ptr(),
                                       self.w.vec[name + ".bias"].data_ptr(), x.N, x.C, groups, x.HW, 1e-5, a.data_ptr(),
                                       b.data_ptr(), self.stream)
            return a, b
        st = self.stats(x) if st is None else self.dense_stats(st)
        self.L.gn_coeffs_f32(st.data_ptr(), self.w.vec[name + ".weight"].data_ptr(),
                             self.w.vec[name + ".bias"].data_ptr(), x.N, x.C, groups, x.HW, 1e-5, a.data_ptr(),
                             b.data_ptr(), self.stream)
        return a, b

    def gn(self, x, name, groups, relu=False, res=None, out=None, st=None, res_scale=None):
        a, b = self.gn_ab(x, name, groups, st)
        return self.affine(x, a, b, res=res, res_scale=res_scale, relu=relu, out=out)

    def gct_gate(self, x, name, st=None, pre=None):
        st = self.stats(x) if st is None else self.dense_stats(st)
        a = self.empty(x.N * x.C)
        v = self.w.vec
        self.L.gct_coeffs_f32(st.data_ptr(), v[name + ".alpha"].data_ptr(), v[name + ".gamma"].data_ptr(),
                              v[name + ".beta"].data_ptr(), _p(pre), x.N, x.C, 1e-5, a.data_ptr(), self.stream)
        return a

    def gap(self, x, st=None):
        st = self.stats(x) if st is None else self.dense_stats(st)
        out = self.empty(x.N * x.C)
        self.L.gap_from_stats_f32(st.data_ptr(), x.N, x.C, x.HW, out.data_ptr(), self.stream)
        return out

    def linear(self, x, wname, N, act=0, ldx=None, weight=None, bias=None, out=None, ldy=None):
        W = self.w.vec[wname + ".weight"] if weight is None else weight
        b = self.w.vec[wname + ".bias"] if bias is None else bias
        M, K = W.shape
        if out is None:
            out = self.empty(N * M)
        self.L.linear_f32(x.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), N, M, K, K if ldx is None else ldx,
                          M if ldy is None else ldy, act, self.stream)
        return out

    def resize_bilinear(self, x, Ho, Wo, out=None):
        if out is None:
            out = self.new(x.N, Ho, Wo, x.C)
        self.L.resize_bilinear_nhwc_f32(x.ptr, None, None, 0, out.ptr, x.N, x.H, x.W, Ho, Wo, x.C, x.ld, out.ld,
                                        self.stream)
        return out

    def copy_channels(self, x, out):
        self.L.copy_channels_f32(x.ptr, out.ptr, x.N * x.HW, x.C, x.ld, out.ld, self.stream)
        return out

    # ------------------------------------------------------------------ backbone (networks/deeplab/*)
    def _res_block(self, x, p, stride, dil):
        y = self.conv(x, p + ".conv1", relu=True)
        y = self.conv(y, p + ".conv2", stride=stride, pad=dil, dil=dil, relu=True)
        if (p + ".downsample.0") in self.w.conv:
            x = self.conv(x, p + ".downsample.0", stride=stride)
        return self.conv(y, p + ".conv3", relu=True, res=x)

    def extract_feature(self, img):
        """aocnet.py:109-112 + deeplab.py:27-38.  img [1,3,H,W] fp32 -> (emb T[1,h,w,100], low T[1,h,w,256])"""
        L = self.L
        assert img.dim() == 4 and img.shape[0] == 1 and img.shape[1] == 3
        img = img.to(device=self.dev, dtype=torch.float32).contiguous()
        H, W = int(img.shape[2]), int(img.shape[3])
        bb = "feature_extracter.backbone"
        if self.stem_s2d:
            # resnet.py:108-110 (7x7 / stride 2 / pad 3) as a 4x4 / stride-1 / pad-1 convolution over the space-to-depth
            # frame: 16 stages of 16 real channels instead of 49 stages of 4 real channels padded to 16
            x16 = self.new(1, (H + 1) // 2 + 1, (W + 1) // 2 + 1, 16)
            L.image_to_s2d16_f32(img.data_ptr(), x16.ptr, H, W, self.stream)
            x = self.conv(x16, bb + ".conv1.s2d", stride=1, pad=1, relu=True)
        else:
            x4 = self.new(1, H, W, 4)
            L.image_to_nhwc4_f32(img.data_ptr(), x4.ptr, H, W, self.stream)
            x = self.conv(x4, bb + ".conv1", stride=2, pad=3, relu=True)
        Hp, Wp = (x.H + 2 - 3) // 2 + 1, (x.W + 2 - 3) // 2 + 1
        xp = self.new(1, Hp, Wp, 64)
        L.maxpool3x3s2_nhwc_f32(x.ptr, xp.ptr, 1, x.H, x.W, 64, self.stream)
        x = xp
        low = None
        for lname, blocks, stride, dils in (("layer1", 3, 1, [1, 1, 1]), ("layer2", 4, 2, [1] * 4),
                                            ("layer3", 23, 2, [1] * 23), ("layer4", 3, 1, [2, 4, 8])):
            for b in range(blocks):
                x = self._res_block(x, "%s.%s.%d" % (bb, lname, b), stride if b == 0 else 1, dils[b])
            if lname == "layer1":
                low = x
        # ASPP (deeplab/aspp.py:62-74): 4 atrous branches + image pooling, written into one concat buffer
        a = "feature_extracter.aspp"
        cat = self.new(1, x.H, x.W, 1280)
        self.conv(x, a + ".aspp1.atrous_conv", relu=True, out=cat.slice(0, 256))
        for i, dl in ((2, 6), (3, 12), (4, 18)):
            self.conv(x, "%s.aspp%d.atrous_conv" % (a, i), pad=dl, dil=dl, relu=True, out=cat.slice(256 * (i - 1), 256))
        g = T(self.gap(x), 1, 1, 1, 2048)
        g = self.conv(g, a + ".global_avg_pool.1", relu=True)
        self.resize_bilinear(g, x.H, x.W, out=cat.slice(1024, 256))
        y = self.conv(cat, a + ".conv1", relu=True)
        # decoder (deeplab/decoder.py:32-41)
        d = "feature_extracter.decoder"
        cat2 = self.new(1, low.H, low.W, 304)
        self.resize_bilinear(y, low.H, low.W, out=cat2.slice(0, 256))
        self.conv(low, d + ".conv1", relu=True, out=cat2.slice(256, 48))
        y = self.conv(cat2, d + ".last_conv.0", pad=1, relu=True)
        y = self.conv(y, d + ".last_conv.4", pad=1, relu=True)
        # semantic embedding (aocnet.py:19-25)
        v = self.w.vec
        z = self.new(1, y.H, y.W, 256)
        L.dwconv3x3_nhwc_f32(y.ptr, v["seperate_conv.weight"].data_ptr(), v["seperate_conv.bias"].data_ptr(), z.ptr, 1,
                             y.H, y.W, 256, self.stream)
        if self.tc_conv:
            a1, b1 = self.gn_ab(z, "bn1", 32)
            z, s2 = self.conv(z, "embedding_conv", in_scale=a1, in_shift=b1, in_relu=True, stats=True)
            emb = self.gn(z, "bn2", 25, relu=True, st=s2)
        else:
            z = self.gn(z, "bn1", 32, relu=True)
            z = self.conv(z, "embedding_conv")
            emb = self.gn(z, "bn2", 25, relu=True)
        return emb, low

    # ------------------------------------------------------------------ matching (aocnet.py:128-358)
    def _as_nhwc_emb(self, e, h, w):
        """[1,100,h,w] tensor (ours: channels-last view; foreign: NCHW) -> flat NHWC tensor [h*w*100]"""
        assert e.shape[0] == 1 and e.shape[1] == EMB and e.shape[2] == h and e.shape[3] == w, tuple(e.shape)
        if e.device == self.dev and e.dtype == torch.float32 and e.stride()[1] == 1 and e.stride()[3] == EMB and \
                e.stride()[2] == w * EMB:
            return e.permute(0, 2, 3, 1).reshape(-1)
        src = e.to(device=self.dev, dtype=torch.float32).contiguous()
        out = self.empty(h * w * EMB)
        self.L.nchw_to_nhwc_f32(src.data_ptr(), out.data_ptr(), 1, EMB, h * w, EMB, self.stream)
        return out

    def _label_ids(self, mask, h, w, out=None):
        """[1,1,H,W] integer label map -> uint8 ids [h*w] (nearest resize, aocnet.py:128-135)"""
        m = mask.to(device=self.dev)
        if m.dtype not in (torch.uint8, torch.int64):
            m = m.to(torch.int64)                     # (rare: the reference hands uint8 ground truth or int64 argmax maps)
        m = m.contiguous()
        Hm, Wm = int(m.shape[-2]), int(m.shape[-1])
        if out is None:
            out = self.empty(h * w, torch.uint8)
        fn = self.L.resize_nearest_u8 if m.dtype == torch.uint8 else self.L.resize_nearest_i64
        fn(m.data_ptr(), out.data_ptr(), Hm, Wm, h, w, self.stream)
        return out

    def _sync_bank(self, ref_embeddings, ref_masks, h, w):
        bk = self.bank
        hw = h * w
        ok = bk.hw == hw and len(ref_embeddings) >= bk.n and len(ref_embeddings) == len(ref_masks)
        if ok:
            for i in range(bk.n):
                if bk.refs[i] is not ref_embeddings[i] or bk.masks[i] is not ref_masks[i]:
                    ok = False
                    break
        if not ok:
            bk.reset()
            bk.hw = hw
        F = len(ref_embeddings)
        cap = 0 if bk.emb_all is None else bk.ids_all.numel() // hw
        if F > cap:
            ncap = max(F, 2 * cap, 8)
            emb_all = self.empty(ncap * hw * EMB)
            ids_all = self.empty(ncap * hw, torch.uint8)
            if bk.n:
                emb_all[:bk.n * hw * EMB].copy_(bk.emb_all[:bk.n * hw * EMB])
                ids_all[:bk.n * hw].copy_(bk.ids_all[:bk.n * hw])
            bk.emb_all, bk.ids_all = emb_all, ids_all
        for i in range(bk.n, F):
            e = self._as_nhwc_emb(ref_embeddings[i], h, w)
            self.L.copy_channels_f32(e.data_ptr(), bk.emb_all.data_ptr() + 4 * i * hw * EMB, hw, EMB, EMB, EMB,
                                     self.stream)
            self._label_ids(ref_masks[i], h, w, out=bk.ids_all[i * hw:(i + 1) * hw])
            bk.refs.append(ref_embeddings[i])
            bk.masks.append(ref_masks[i])
            bk.version += 1
        bk.n = F
        return F

    def _bank_index(self, ref_embeddings, ref_masks, h, w, O):
        """Object-sorted view of the bank (matching.py:2486-2495, :533-545).  Rebuilt -- with the one host
        synchronisation of the path, for the per-object pixel counts the k-means RNG draws need -- only when the
        bank changed (every MEM_EVERY-th frame in the reference eval loop)."""
        L, st = self.L, self.stream
        hw = h * w
        F = self._sync_bank(ref_embeddings, ref_masks, h, w)
        bk = self.bank
        ix = bk.index
        if ix is not None and ix["version"] == bk.version and ix["O"] == O:
            return ix
        total = F * hw
        meta = self.empty(META_INTS, torch.int32)
        cap_rows = total + O * BANK_ALIGN
        row_src = self.grow("bank.row_src", cap_rows, torch.int32)
        nat2sorted = self.grow("bank.nat2sorted", max(total, 1), torch.int32)
        nws = L.bank_workspace_bytes(total, O)
        L.bank_index_build(bk.ids_all.data_ptr(), total, O, BANK_ALIGN, meta.data_ptr(), row_src.data_ptr(), cap_rows,
                           nat2sorted.data_ptr(), self.ws("bank", nws).data_ptr(), nws, st)
        self._meta_host.copy_(meta, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()   # the host needs the per-object counts (RNG draws)
        mh = self._meta_host.numpy().copy()
        counts = [int(mh[o]) for o in range(O)]
        rows = int(mh[2 * MAXO + 1])
        S = self.grow("bank.S", max(rows, 1) * EMB, torch.float32)       # rebuilt in place: the stream was just drained, and a
        r2 = self.grow("bank.r2", max(rows, 1), torch.float32)          # graph of the previous bank version is never replayed again
        L.bank_gather_f32(bk.emb_all.data_ptr(), row_src.data_ptr(), rows, S.data_ptr(), r2.data_ptr(), st)
        ix = dict(version=bk.version, O=O, total=total, meta=meta, mh=mh, counts=counts, rows=rows, S=S, r2=r2,
                  nat2sorted=nat2sorted, maxrows=max(counts) if counts else 0, hw=hw)
        bk.index = ix
        return ix

    def _draw_kmeans_init(self, ix, O):
        """k chain + init rows from numpy's GLOBAL RNG, exactly the stream scipy's kmeans2(minit='points') consumes
        (matching.py:556,562): one np.random.choice per object with pixels, in id order."""
        kk = np.zeros(MAXO, dtype=np.int32)
        init = np.zeros((MAXO, self.kmax), dtype=np.int32)
        k = self.cluster_num
        counts = ix["counts"]
        for o in range(O):
            k = min(k, counts[o])                     # matching.py:556 -- carries over to later objects
            if k == 0:
                continue
            kk[o] = k
            init[o, :k] = np.random.choice(counts[o], size=int(k), replace=False)   # == scipy _kpoints
        return kk, init

    def _match_front(self, q, ix, O, kk_d, init_d, head):
        """bank-dependent part: pixel-level global matching, adaptive proxies (k-means), bank attention heads.
        -> (g [hw*O], P, pvalid, cent, labels)"""
        L, st = self.L, self.stream
        hw, rows, total = q.HW, ix["rows"], ix["total"]
        bk = self.bank
        bias = self.w.vec["dis_bias"]
        S, r2, meta = ix["S"], ix["r2"], ix["meta"]
        # --- pixel-level global matching (matching.py:2384)
        g = self.empty(hw * O)
        if self.shard is not None:
            # one sequence over several GPUs: this rank's share of the bank rows, partial minima exchanged by the kernel
            sh = self.shard
            assert self.tc_match and rows > 0 and hw <= sh["cap_hw"]
            nws2 = L.global_match_tc_workspace_bytes(hw, rows)
            L.global_match_tc_sharded(q.ptr, hw, S.data_ptr(), r2.data_ptr(), meta.data_ptr(), rows, bias.data_ptr(), O,
                                      sh["rank"], sh["world"], sh["areas"], sh["cap_hw"], sh["state"].data_ptr(),
                                      self.ws("gm_tc", nws2).data_ptr(), nws2, g.data_ptr(), st)
        elif self.tc_match and rows > 0:
            nws2 = L.global_match_tc_workspace_bytes(hw, rows)
            L.global_match_tc(q.ptr, hw, S.data_ptr(), r2.data_ptr(), meta.data_ptr(), rows, bias.data_ptr(), O,
                              self.ws("gm_tc", nws2).data_ptr(), nws2, g.data_ptr(), st)
        else:
            mins = self.empty(hw * O)
            L.global_match_simt_f32(q.ptr, hw, S.data_ptr(), r2.data_ptr(), meta.data_ptr(), bias.data_ptr(), O,
                                    mins.data_ptr(), g.data_ptr(), st)
        # --- adaptive object proxies (matching.py:533-595)
        kmax, slots = self.kmax, self.proxy_slots
        cent = self.empty(MAXO * kmax * EMB)
        labels = self.empty(max(rows, 1), torch.int32)
        P = self.zeros(MAXO * slots * EMB)
        pvalid = self.zeros(MAXO * slots, torch.int32)
        kws = L.kmeans_workspace_bytes(ix["maxrows"], O, kmax)
        L.kmeans_proxies_f32(S.data_ptr(), meta.data_ptr(), ix["nat2sorted"].data_ptr(), kk_d.data_ptr(),
                             init_d.data_ptr(), O, ix["maxrows"], self.kmeans_iters, kmax, cent.data_ptr(),
                             labels.data_ptr(), P.data_ptr(), pvalid.data_ptr(), self.ws("kmeans", kws).data_ptr(), kws, st)
        if self.force_proxies is not None:
            self._apply_forced_proxies(P, pvalid, O)
        # --- bank attention heads / k=1 proxies (attention.py:155-189)
        hws = L.head_pool_workspace_bytes(max(total, hw))
        hbuf = self.ws("headpool", hws)
        L.head_pool_f32(bk.emb_all.data_ptr(), bk.ids_all.data_ptr(), total, O, 1e-5, head.data_ptr(), HEAD, 0, EMB,
                        P.data_ptr() + 4 * 2 * kmax * EMB, slots * EMB, hbuf.data_ptr(), hws, st)
        return g, P, pvalid, cent, labels

    def _apply_forced_proxies(self, P, pvalid, O):
        """TEST HOOK (tests/test_gpu_fullsize.py): replace the k-means proxies of this frame by given ones -- dict with
        prox_cen / prox_avg [O,16,100] and prox_ncen / prox_navg [O] as stored by tools/make_cfg_truth.py.  k-means is a
        discrete step (a 1e-6 change of an embedding can flip a boundary row's cluster and move the logits by 0.1), so
        full-size logit parity is measured with both sides on the SAME proxies; the k-means kernels themselves are
        compared bit for bit on identical inputs in tests/test_gpu_ops.py.  Plain launches only."""
        assert not torch.cuda.is_current_stream_capturing(), "force_proxies needs AOCB200_GRAPHS=0 / use_graphs = False"
        f = self.force_proxies
        assert self.kmax == 16
        Pv, vv = P.view(MAXO, PROXY_SLOTS, EMB), pvalid.view(MAXO, PROXY_SLOTS)
        ar = torch.arange(16, device=self.dev).view(1, 16)
        for key, cnt, lo in (("prox_cen", "prox_ncen", 0), ("prox_avg", "prox_navg", 16)):
            Pv[:O, lo:lo + 16] = torch.as_tensor(f[key], dtype=torch.float32, device=self.dev)
            n = torch.as_tensor(f[cnt], dtype=torch.int32, device=self.dev).view(O, 1)
            vv[:O, lo:lo + 16] = (ar < n).to(torch.int32)

    def _match_back(self, q, g, P, pvalid, head, prev_e, prev_ids, O):
        """bank-independent part: previous-frame heads, cluster/proxy matching, local matching, pre-head.
        -> x T[O,h,w,164]"""
        L, st = self.L, self.stream
        h, w, hw = q.H, q.W, q.HW
        bias = self.w.vec["dis_bias"]
        hws = L.head_pool_workspace_bytes(hw)
        hbuf = self.ws("headpool", hws)
        prev_pos = self.empty(O * EMB)
        L.head_pool_f32(prev_e.data_ptr(), prev_ids.data_ptr(), hw, O, 1e-5, head.data_ptr(), HEAD, 2 * EMB, 3 * EMB,
                        prev_pos.data_ptr(), EMB, hbuf.data_ptr(), hws, st)
        # --- cluster-level + proxy-level matching (matching.py:602-637, :149-197)
        gc = self.empty(hw * O * 2)
        gp = self.empty(hw * O)
        L.proxy_match_f32(q.ptr, hw, P.data_ptr(), pvalid.data_ptr(), bias.data_ptr(), O, self.kmax, gc.data_ptr(),
                          gp.data_ptr(), st)
        # --- local matching on the half-resolution grid (matching.py:2710-2851)
        hh, ww = h // 2 + 1, w // 2 + 1
        ldl = (6 * O + 3) // 4 * 4
        xq = self.resize_bilinear(q, hh, ww)
        yp = self.resize_bilinear(T(prev_e, 1, h, w, EMB), hh, ww)
        ids_lr = self.empty(hh * ww, torch.uint8)
        L.resize_nearest_u8(prev_ids.data_ptr(), ids_lr.data_ptr(), h, w, hh, ww, st)
        x2, y2 = self.empty(hh * ww), self.empty(hh * ww)
        L.row_sqnorm_f32(xq.ptr, hh * ww, x2.data_ptr(), st)
        L.row_sqnorm_f32(yp.ptr, hh * ww, y2.data_ptr(), st)
        loc_lr = T(self.zeros(hh * ww * ldl), 1, hh, ww, ldl)
        L.local_match_f32(xq.ptr, yp.ptr, x2.data_ptr(), y2.data_ptr(), ids_lr.data_ptr(), hh, ww, O, bias.data_ptr(),
                          loc_lr.ptr, ldl, st)
        loc = self.resize_bilinear(loc_lr, h, w)
        # local proxy matching: previous-frame pixels replaced by their object's mean proxy (aocnet.py:325-337)
        yq = self.new(1, hh, ww, EMB)
        L.resize_bilinear_nhwc_f32(None, prev_ids.data_ptr(), prev_pos.data_ptr(), O, yq.ptr, 1, h, w, hh, ww, EMB, EMB,
                                   EMB, st)
        L.row_sqnorm_f32(yq.ptr, hh * ww, y2.data_ptr(), st)
        locp_lr = T(self.zeros(hh * ww * ldl), 1, hh, ww, ldl)
        L.local_match_f32(xq.ptr, yq.ptr, x2.data_ptr(), y2.data_ptr(), ids_lr.data_ptr(), hh, ww, O, bias.data_ptr(),
                          locp_lr.ptr, ldl, st)
        locp = self.resize_bilinear(locp_lr, h, w)
        # --- fg->bg, concat, pre-head (aocnet.py:349-362, decoding_module.py:228-240)
        pre = self.new(O, h, w, 24)
        L.prehead_assemble_f32(g.data_ptr(), gc.data_ptr(), gp.data_ptr(), loc.ptr, locp.ptr, ldl, prev_ids.data_ptr(),
                               hw, O, pre.ptr, st)
        x = self.new(O, h, w, EMB + 64)
        t, s_t = self.conv(pre, "dynamic_prehead.conv", stats=True)
        self.gn(t, "dynamic_prehead.bn", 16, relu=True, out=x.slice(EMB, 64), st=s_t)
        L.broadcast_rows_f32(q.ptr, x.ptr, O, hw, EMB, q.ld, x.ld, st)
        if self.keep_debug:
            self.debug.update(g=g, gc=gc, gp=gp, loc=loc, locp=locp, pre=pre, head=head, P=P, pvalid=pvalid,
                              prev_ids=prev_ids, ldl=ldl)
        return x

    def match_features(self, ref_embeddings, ref_masks, prev_embedding, prev_mask, emb, K):
        """-> (x T[O,h,w,164] decoder input, head [O*400], prev_ids) ; aocnet.py:128-362 (eager schedule)"""
        O = K + 1
        assert 1 <= O <= MAXO, "at most %d objects" % (MAXO - 1)
        h, w = emb.H, emb.W
        ix = self._bank_index(ref_embeddings, ref_masks, h, w, O)
        kk, init = self._draw_kmeans_init(ix, O)
        kk_d = torch.from_numpy(kk).to(self.dev)
        init_d = torch.from_numpy(init).to(self.dev)
        head = self.empty(O * HEAD)
        g, P, pvalid, cent, labels = self._match_front(emb, ix, O, kk_d, init_d, head)
        prev_e = self._as_nhwc_emb(prev_embedding, h, w)
        prev_ids = self._label_ids(prev_mask, h, w)
        x = self._match_back(emb, g, P, pvalid, head, prev_e, prev_ids, O)
        if self.keep_debug:
            rows_ = max(ix["rows"], 1)                    # (the bank arrays are capacity buffers: views of the live part)
            self.debug.update(labels=labels, cent=cent, meta=ix["mh"].copy(), S=ix["S"][:rows_ * EMB], r2=ix["r2"][:rows_],
                              nat2sorted=ix["nat2sorted"][:max(ix["total"], 1)], kk=kk, init=init)
        return x, head, prev_ids

    # ------------------------------------------------------------------ calibration decoder (decoding_module.py)
    def ia_gate(self, x, head_t, ldh, name):
        a = self.linear(head_t, name + ".IA", x.N, act=1, ldx=ldh)
        return self.affine(x, a)

    def gn_bottleneck(self, x, p, stride=1, dil=1, pre=None, x_stats=None, want_stats=False):
        """layers/gct.py:68-91 applied to x*pre[n,c] (pre = the IA gate / conditioning-block scale in front of the
        block, folded in instead of materialised); x_stats = statistics of x when the producer already has them;
        want_stats -> (out, statistics of out)"""
        if not self.tc_conv:
            if pre is not None:
                x = self.affine(x, pre)
            gate = self.gct_gate(x, p + ".GCT1")
            y = self.conv(x, p + ".conv1", in_scale=gate)
            y = self.gn(y, p + ".bn1", 32, relu=True)
            y = self.conv(y, p + ".conv2", stride=stride, pad=dil, dil=dil)
            y = self.gn(y, p + ".bn2", 32, relu=True)
            y = self.conv(y, p + ".conv3")
            if (p + ".downsample.0") in self.w.conv:
                r = self.conv(x, p + ".downsample.0", stride=stride)
                r = self.gn(r, p + ".downsample.1", 32)
            else:
                r = x
            out = self.gn(y, p + ".bn3", 32, relu=True, res=r)
            return (out, self.stats(out)) if want_stats else out
        # tensor-core path: GroupNorm statistics come out of each convolution's epilogue and the normalisation (+ReLU)
        # is applied inside the NEXT convolution's operand path -- the normalised tensors bn1/bn2 never exist in HBM
        gate = self.gct_gate(x, p + ".GCT1", st=x_stats, pre=pre)         # = pre * GCT gate of (x*pre)
        y1, s1 = self.conv(x, p + ".conv1", in_scale=gate, stats=True)
        a1, b1 = self.gn_ab(y1, p + ".bn1", 32, s1)
        y2, s2 = self.conv(y1, p + ".conv2", stride=stride, pad=dil, dil=dil, in_scale=a1, in_shift=b1, in_relu=True,
                           stats=True)
        a2, b2 = self.gn_ab(y2, p + ".bn2", 32, s2)
        y3, s3 = self.conv(y2, p + ".conv3", in_scale=a2, in_shift=b2, in_relu=True, stats=True)
        a3, b3 = self.gn_ab(y3, p + ".bn3", 32, s3)
        if (p + ".downsample.0") in self.w.conv:
            r, sr = self.conv(x, p + ".downsample.0", stride=stride, in_scale=pre, stats=True)
            ar, br = self.gn_ab(r, p + ".downsample.1", 32, sr)
            # relu(GN3(y3) + GN_ds(r)) = relu(y3*a3 + (b3 + br) + r*ar)
            return self.affine(y3, a3, self.vec_op(b3, br, 0), res=r, res_scale=ar, relu=True, want_stats=want_stats)
        return self.affine(y3, a3, b3, res=x, res_scale=pre, relu=True, want_stats=want_stats)

    def cond_scale(self, x, p, beta=0.3):
        """conditioning_block / conditioning_layer (conditioning_layer.py:24-86) with CL_2/CL_3 folded:
        -> the FiLM scale a[n,c] = 1 + tanh(.) of x -> a*x (the caller folds it into the next block)"""
        L, v = self.L, self.w.vec
        O, hw, C = x.N, x.HW, x.C
        phi = self.empty(O * hw)
        L.cond_phi_f32(x.ptr, v[p + ".CL_1.phi_layer.weight"].data_ptr(), v[p + ".CL_1.phi_layer.bias"].data_ptr(),
                       phi.data_ptr(), O, hw, C, x.ld, self.stream)
        thr = self.empty(O)
        rank = max(1, int(beta * x.W * x.H))
        L.kth_largest_f32(phi.data_ptr(), O, hw, rank, thr.data_ptr(), self.stream)
        st = self.stats(x, phi, thr)
        gapm = self.gap(x, st)                                   # masked sum / (h*w)
        c1 = self.linear(gapm, p + ".CL_1.mlp_layer", O)
        return self.linear(c1, p + ".fold", O, act=1)

    def cond_block(self, x, p, beta=0.3):
        return self.affine(x, self.cond_scale(x, p, beta))

    def delta_head(self, x, head, st=None):
        """cat([head, sum_objects(GAP(x)) - GAP(x)]) -> [O, 400 + C]"""
        O, C = x.N, x.C
        px = self.gap(x, st)
        out = self.empty(O * (HEAD + C))
        self.L.copy_channels_f32(head.data_ptr(), out.data_ptr(), O, HEAD, HEAD, HEAD + C, self.stream)
        self.L.delta_sum_f32(px.data_ptr(), out.data_ptr() + 4 * HEAD, O, C, HEAD + C, self.stream)
        return out

    def decoder_aspp(self, x, p, pre=None, x_stats=None):
        """layers/aspp.py:57-70 applied to x*pre[n,c]"""
        O = x.N
        st = self.stats(x) if x_stats is None else x_stats
        cat = self.new(O, x.H, x.W, 640)
        for i, d in ((1, 0), (2, 6), (3, 12), (4, 18)):
            q = "%s.aspp%d" % (p, i)
            gate = self.gct_gate(x, q + ".GCT", st=st, pre=pre)
            y, s_y = self.conv(x, q + ".atrous_conv", pad=d, dil=max(d, 1), in_scale=gate, stats=True)
            self.gn(y, q + ".bn", 32, relu=True, out=cat.slice(128 * (i - 1), 128), st=s_y)
        gp = self.gap(x, st)
        g = T(gp if pre is None else self.vec_op(gp, pre, 1), O, 1, 1, x.C)
        g = self.conv(g, p + ".global_avg_pool.1", relu=True)
        self.resize_bilinear(g, x.H, x.W, out=cat.slice(512, 128))
        gate = self.gct_gate(cat, p + ".GCT")
        y, s_y = self.conv(cat, p + ".conv1", in_scale=gate, stats=True)
        return self.gn(y, p + ".bn1", 32, relu=True, st=s_y)

    def modulator(self, x, mem, head, p, tag):
        """decoding_module.py:192-210"""
        cat = self.new(x.N, x.H, x.W, x.C + mem.C)
        self.copy_channels(x, cat.slice(0, x.C))
        self.copy_channels(mem, cat.slice(x.C, mem.C))
        x = cat
        st = None
        for i in (1, 2, 3):
            a = self.linear(head, "%s.%s_Reweight_Layer_%d.IA" % (p, tag, i), x.N, act=1, ldx=HEAD)     # IA gate, folded
            name = "%s.%s_Bottleneck_%d" % (p, tag, i)
            if i < 3:
                x, st = self.gn_bottleneck(x, name, pre=a, x_stats=st, want_stats=True)
            else:
                x = self.gn_bottleneck(x, name, pre=a, x_stats=st)
        return x

    def _mem_T(self, m, like):
        """caller-held memory tensor ([O,256,h2,w2] view of ours, or foreign NCHW) -> T or None"""
        if m is None or tuple(m.shape) != (like.N, like.C, like.H, like.W):
            return None
        if m.device == self.dev and m.dtype == torch.float32 and m.stride()[1] == 1 and m.stride()[3] == like.C:
            return T(m.permute(0, 2, 3, 1).reshape(-1), like.N, like.H, like.W, like.C)
        src = m.to(device=self.dev, dtype=torch.float32).contiguous()
        out = self.new(like.N, like.H, like.W, like.C)
        self.L.nchw_to_nhwc_f32(src.data_ptr(), out.ptr, like.N, like.C, like.HW, like.C, self.stream)
        return out

    def calibration_decoding(self, x, head, memory, low):
        """decoding_module.py:96-149 -> (logits tensor [O*h*w] as [O][h][w], [mem0, mem1] as T)"""
        L, v = self.L, self.w.vec
        p = "dynamic_seghead"
        O, h, w = x.N, x.H, x.W
        # every per-(object, channel) gate of the reference (IA_gate, conditioning_block) is a scale a[n,c] in front
        # of a block: it is folded into the block (GCT statistics, conv operand path, residual) instead of written out
        a = self.linear(head, p + ".IA1.IA", O, act=1, ldx=HEAD)
        x, st = self.gn_bottleneck(x, p + ".layer1", pre=a, want_stats=True)
        x, st = self.gn_bottleneck(x, p + ".layer2", 1, 2, pre=self.cond_scale(x, p + ".CLB2"), x_stats=st, want_stats=True)
        x, st = self.gn_bottleneck(x, p + ".layer3", 2, 1, pre=self.cond_scale(x, p + ".CLB3"), x_stats=st, want_stats=True)
        x, st = self.gn_bottleneck(x, p + ".layer4", 1, 2, pre=self.cond_scale(x, p + ".CLB4"), x_stats=st, want_stats=True)
        x, st = self.gn_bottleneck(x, p + ".layer5", 1, 4, pre=self.cond_scale(x, p + ".CLB5"), x_stats=st, want_stats=True)
        dh = self.delta_head(x, head, st)
        a = self.linear(dh, p + ".IA9.IA", O, act=1, ldx=HEAD + x.C)
        x = self.decoder_aspp(x, p + ".ASPP", pre=a, x_stats=st)
        cur1 = x
        m0 = self._mem_T(memory[0], cur1) or cur1
        x = self.modulator(x, m0, head, p, "M1")
        cur2 = x
        m1 = self._mem_T(memory[1], cur2) or cur2
        x = self.modulator(x, m1, head, p, "M2")
        # decoder_final (decoding_module.py:162-190, repair R9)
        cat = self.new(O, h, w, 512)                       # [low (256, broadcast over objects) | x upsampled (256)]
        L.broadcast_rows_f32(low.ptr, cat.ptr, O, h * w, 256, low.ld, 512, self.stream)
        xu = cat.slice(256, 256)
        L.resize_bicubic_nhwc_f32(x.ptr, xu.ptr, O, x.H, x.W, h, w, 256, x.ld, 512, self.stream)
        gate = self.gct_gate(cat, p + ".GCT_sc")
        sc, s_sc = self.conv(cat, p + ".conv_sc", in_scale=gate, stats=True)
        cat2 = self.new(O, h, w, 320)                      # [x (256) | sc (64)]
        self.copy_channels(xu, cat2.slice(0, 256))
        self.gn(sc, p + ".bn_sc", 16, relu=True, out=cat2.slice(256, 64), st=s_sc)
        x = cat2
        a = self.linear(self.delta_head(x, head), p + ".IA10.IA", O, act=1, ldx=HEAD + x.C)
        y, s_y = self.conv(x, p + ".conv1", pad=1, in_scale=a, stats=True)
        a1, b1 = self.gn_ab(y, p + ".bn1", 32, s_y)
        x, st = self.affine(y, a1, b1, relu=True, want_stats=True)
        a = self.linear(self.delta_head(x, head, st), p + ".IA11.IA", O, act=1, ldx=HEAD + x.C)
        y, s_y = self.conv(x, p + ".conv2", pad=1, in_scale=a, stats=True)
        x = self.gn(y, p + ".bn2", 32, relu=True, st=s_y)
        wfg = self.linear(head, p + ".IA_final_fg", O)
        wbg = self.linear(head, p + ".IA_final_bg", O)
        fg, bg, logits = self.empty(O * h * w), self.empty(O * h * w), self.empty(O * h * w)
        L.dyn_logits_f32(x.ptr, wfg.data_ptr(), wbg.data_ptr(), fg.data_ptr(), bg.data_ptr(), logits.data_ptr(), O,
                         h * w, x.C, x.ld, self.stream)
        return logits, [cur1, m1]

    # ------------------------------------------------------------------ per-frame entry (aocnet.py:84-107)
    def _num_objects(self, gt_ids):
        if isinstance(gt_ids, int):
            return gt_ids
        # A device tensor costs one D2H sync: read once per tensor OBJECT (held strongly, so its identity cannot be
        # recycled by the allocator) and version; a caller that builds a fresh tensor per frame pays the read per frame,
        # DeviceSequence passes a Python int.
        t, ver, k = self._gt_cache
        if t is not gt_ids or ver != getattr(gt_ids, "_version", 0):
            k = int(gt_ids[0]) if hasattr(gt_ids, "__getitem__") else int(gt_ids)
            self._gt_cache = (gt_ids, getattr(gt_ids, "_version", 0), k)
        return k

    def set_seen_labels(self, labels=None):
        """Labels that appeared in a ground-truth frame of this sequence so far (`label_all_list`,
        eval_manager_mm.py:262-265); None = every slot.  Shapes only `last_label` / `last_conf_label` (the argmax and
        the confident mask over the seen slots); the probabilities forward_for_eval returns are always the reference's
        plain softmax, so a sequence driver that sets this (DeviceSequence) cannot change what another caller gets."""
        self._set_exist_bits(self.seen_bits(labels))

    @staticmethod
    def seen_bits(labels):
        bits = -1
        if labels is not None:
            bits = 0
            for v in labels:
                if 0 <= int(v) < MAXO:
                    bits |= 1 << int(v)
            bits -= (1 << 32) if bits >= (1 << 31) else 0
        return bits

    def _set_exist_bits(self, bits):
        if bits != self._exist_bits:
            self._exist.copy_(torch.tensor([bits], dtype=torch.int32))    # pageable source: staged before the call returns
            self._exist_bits = bits

    def _upsample_softmax(self, logits, O, h, w, H, W):
        """aocnet.py:100-107 -> probs (the plain softmax the reference returns; None when `want_probs` is off), plus the
        eval loop's derived maps: label = argmax over the seen slots, conf = label or 125 (see set_seen_labels)"""
        probs = torch.empty((1, O, H, W), dtype=torch.float32, device=self.dev) if self.want_probs else None
        label = torch.empty((H, W), dtype=torch.uint8, device=self.dev)
        conf = torch.empty((H, W), dtype=torch.uint8, device=self.dev)
        self.L.upsample_softmax_label_f32(logits.data_ptr(), _p(probs), label.data_ptr(), conf.data_ptr(), None,
                                          self._exist.data_ptr(), float(self.unc_ratio), O, h, w, H, W, self.stream)
        return probs, label, conf

    # ------------------------------------------------------------------ fp16 operand range guard
    # The split-fp16 convolution clamps |x| at 65504.  Every convolution launch ORs "an operand reached 6e4" into the
    # device word `_ovf`; after each frame the word is copied to pinned host memory behind an event.
    #   sync      wait for the frame, and if the word is set: switch this engine to 3xTF32 operands, restore numpy's RNG
    #             state (the k-means draws of the frame), run the frame again, warn.  Costs one host wait per frame.
    #   deferred  (default) never waits: the word of frame t is looked at when a later call finds its event complete; a
    #             set word raises AocError naming the frame (its outputs, already handed out, are invalid) after
    #             switching the engine to 3xTF32, so the caller can repeat the sequence.
    #   off       no check.
    def _overflow_trip(self, frame, rerun):
        import warnings
        from .lib import AocError
        torch.cuda.synchronize(self.dev)
        self.overflow_frames.append(frame)
        self.conv_mode = 0
        self.L.fill_u32(self._ovf.data_ptr(), 0, 1, self.stream)
        self.drop_graphs()                                  # graphs captured with split-fp16 kernels / weight images
        for slot in self._ovf_ring:
            slot[2][0] = None
        msg = ("aocb200: a convolution input of frame %d reached the fp16 range (|x| >= 6e4); the split-fp16 operand "
               "mode would have clamped it.  This engine now uses 3xTF32 operands (fp32 range, ~12 %% slower)." % frame)
        if rerun:
            warnings.warn(msg + "  The frame was run again.", RuntimeWarning)
        else:
            raise AocError(msg + "  Outputs returned for that frame (and later ones) are invalid: repeat the sequence, "
                           "or set AOCB200_OVERFLOW=sync for transparent re-runs.")

    def _overflow_poll(self, block=False):
        for host, ev, tag in self._ovf_ring:
            if tag[0] is not None and (block or ev.query()):
                if block:
                    ev.synchronize()
                frame, tag[0] = tag[0], None
                if int(host[0]) != 0:
                    self._overflow_trip(frame, rerun=False)

    def forward_for_eval(self, memory_prev_list, ref_embeddings, ref_masks, prev_embedding, prev_mask, current_frame,
                         pred_size, gt_ids):
        pol = self.overflow_policy if (self.tc_conv and self.conv_mode == 1) else "off"
        if pol == "off":
            self.frame_no += 1
            return self._forward(memory_prev_list, ref_embeddings, ref_masks, prev_embedding, prev_mask, current_frame,
                                 pred_size, gt_ids)
        self._overflow_poll()
        rng = np.random.get_state() if pol == "sync" else None
        out = self._forward(memory_prev_list, ref_embeddings, ref_masks, prev_embedding, prev_mask, current_frame,
                            pred_size, gt_ids)
        frame = self.frame_no
        self.frame_no += 1
        host, ev, tag = self._ovf_ring[self._ovf_turn % len(self._ovf_ring)]
        self._ovf_turn += 1
        if tag[0] is not None:                               # slot still carries an unchecked frame: check it first
            ev.synchronize()
            old, tag[0] = tag[0], None
            if int(host[0]) != 0:
                self._overflow_trip(old, rerun=False)
        host.copy_(self._ovf, non_blocking=True)
        ev.record(torch.cuda.current_stream(self.dev))
        tag[0] = frame
        if pol == "sync":
            ev.synchronize()
            tag[0] = None
            if int(host[0]) != 0:
                self._overflow_trip(frame, rerun=True)
                np.random.set_state(rng)
                return self._forward(memory_prev_list, ref_embeddings, ref_masks, prev_embedding, prev_mask,
                                     current_frame, pred_size, gt_ids)
        return out

    def _forward(self, memory_prev_list, ref_embeddings, ref_masks, prev_embedding, prev_mask, current_frame,
                 pred_size, gt_ids):
        if self.use_graphs and not self.keep_debug:
            return self._forward_graphed(memory_prev_list, ref_embeddings, ref_masks, prev_embedding, prev_mask,
                                         current_frame, pred_size, gt_ids)
        emb, low = self.extract_feature(current_frame)
        emb_out = emb.nchw()
        if prev_embedding is None:
            return None, emb_out, memory_prev_list
        K = self._num_objects(gt_ids)
        O = K + 1
        x, head, _ = self.match_features(ref_embeddings, ref_masks, prev_embedding, prev_mask, emb, K)
        logits, mem = self.calibration_decoding(x, head, list(memory_prev_list[0]), low)
        H, W = int(pred_size[0]), int(pred_size[1])
        probs, label, conf = self._upsample_softmax(logits, O, emb.H, emb.W, H, W)
        self.last_logits = logits.view(1, O, emb.H, emb.W)
        self.last_label, self.last_conf_label = label, conf
        return probs, emb_out, [[mem[0].nchw(), mem[1].nchw()]]

    # ------------------------------------------------------------------ CUDA-graph schedule
    # The frame is three captured segments, replayed with a handful of host calls instead of ~570 launches:
    #   A  backbone + embedding            static image buffer -> emb, low            (per input size)
    #   F  bank-dependent matching         emb, sorted bank, RNG init rows -> g, P    (re-captured when the bank grows)
    #   C  local matching + decoder + softmax                                           (per size / object count)
    # Inputs owned by the caller (previous embedding / mask, decoder memory) are copied into static buffers before the
    # replay; everything handed back is a fresh copy, so the caller's bank never aliases a recycled buffer.
    def drop_graphs(self):
        """forget every captured segment and its static buffers (the next frame captures afresh)"""
        self._segA.clear(); self._static.clear()
        self._cap_stream = self._pool = None                # (their memory pool dies with its last graph: take a new one)

    def _capture(self, fn):
        L = self.L
        n0 = L.launches
        g = torch.cuda.CUDAGraph()
        # torch.cuda.graph() would gc.collect() + empty_cache() + device-synchronise on every capture (tens of ms; the
        # bank-dependent segment is re-captured whenever the bank grows): capture by hand on a side stream instead
        if self._cap_stream is None:
            self._cap_stream = torch.cuda.Stream(device=self.dev)
            self._pool = torch.cuda.graph_pool_handle()
        cur = torch.cuda.current_stream(self.dev)
        self._cap_stream.wait_stream(cur)
        gr = _Graph(g, 0, L)
        with torch.cuda.stream(self._cap_stream):
            g.capture_begin(pool=self._pool)
            L.capture_tag = id(gr)
            try:
                out = fn()
            finally:
                L.capture_tag = None
                g.capture_end()
        cur.wait_stream(self._cap_stream)
        gr.n = L.launches - n0
        L.launches = n0                     # captured, not executed: every replay() accounts for its kernels
        return gr, out

    def _forward_graphed(self, memory_prev_list, ref_embeddings, ref_masks, prev_embedding, prev_mask, current_frame,
                         pred_size, gt_ids):
        assert current_frame.dim() == 4 and current_frame.shape[0] == 1 and current_frame.shape[1] == 3
        H, W = int(current_frame.shape[2]), int(current_frame.shape[3])
        segA = self._segA.get((H, W))
        if segA is None:
            img = torch.empty((1, 3, H, W), dtype=torch.float32, device=self.dev)
            img.copy_(current_frame, non_blocking=True)
            self.extract_feature(img)                                   # warm-up: packs weights, sizes workspaces
            g, (emb, low) = self._capture(lambda: self.extract_feature(img))
            segA = dict(img=img, graph=g, emb=emb, low=low)
            self._segA[(H, W)] = segA
        else:
            segA["img"].copy_(current_frame, non_blocking=True)
        # The bank index (and its one host synchronisation, when the bank changed) does not depend on the current frame:
        # it is brought up to date BEFORE the backbone is launched, so that the host's share of a bank change (RNG draws,
        # re-capture of segment F) runs under the backbone instead of leaving the GPU idle behind a drained stream.
        ix_early = None
        if prev_embedding is not None:
            ix_early = self._bank_index(ref_embeddings, ref_masks, segA["emb"].H, segA["emb"].W,
                                        self._num_objects(gt_ids) + 1)
        segA["graph"].replay()
        emb, low = segA["emb"], segA["low"]
        h, w, hw = emb.H, emb.W, emb.HW
        emb_out = emb.nchw().clone(memory_format=torch.channels_last)
        if prev_embedding is None:
            return None, emb_out, memory_prev_list
        K = self._num_objects(gt_ids)
        O = K + 1
        assert 1 <= O <= MAXO, "at most %d objects" % (MAXO - 1)
        Hp, Wp = int(pred_size[0]), int(pred_size[1])
        memory = list(memory_prev_list[0])

        # ---- static inputs of this (size, object count)
        keyS = (h, w, O, self.kmax)
        st = self._static.get(keyS)
        if st is None:
            st = dict(prev_e=self.empty(hw * EMB), prev_ids=self.empty(hw, torch.uint8), head=self.empty(O * HEAD),
                      kk=self.empty(MAXO, torch.int32), init=self.empty(MAXO * self.kmax, torch.int32),
                      g=self.empty(hw * O), P=self.empty(MAXO * self.proxy_slots * EMB),
                      pvalid=self.empty(MAXO * self.proxy_slots, torch.int32),
                      mem=[self.new(O, (h + 1) // 2, (w + 1) // 2, 256), self.new(O, (h + 1) // 2, (w + 1) // 2, 256)],
                      host=[(torch.empty(MAXO * (1 + self.kmax), dtype=torch.int32).pin_memory(), torch.cuda.Event())
                            for _ in range(4)], turn=0, segF=None, segC={})
            self._static[keyS] = st
        # ---- bank index (host sync only when the bank changed) and this frame's RNG draws
        ix = ix_early
        kk, init = self._draw_kmeans_init(ix, O)
        hbuf, hev = st["host"][st["turn"] % 4]
        st["turn"] += 1
        hev.synchronize()                                   # the copy that last used this pinned buffer has finished
        hn = hbuf.numpy()
        hn[:MAXO] = kk
        hn[MAXO:] = init.reshape(-1)
        st["kk"].copy_(hbuf[:MAXO], non_blocking=True)
        st["init"].copy_(hbuf[MAXO:], non_blocking=True)
        hev.record()
        st["prev_e"].copy_(self._as_nhwc_emb(prev_embedding, h, w), non_blocking=True)
        self._label_ids(prev_mask, h, w, out=st["prev_ids"])
        has = []
        for i in (0, 1):
            m = self._mem_T(memory[i], st["mem"][i])
            has.append(m is not None)
            if m is not None:
                st["mem"][i].buf.copy_(m.buf, non_blocking=True)

        def front():
            g_, P_, pv_, _, _ = self._match_front(emb, ix, O, st["kk"], st["init"], st["head"])
            st["g"].copy_(g_); st["P"].copy_(P_); st["pvalid"].copy_(pv_)
            return None

        def back():
            x = self._match_back(emb, st["g"], st["P"], st["pvalid"], st["head"], st["prev_e"], st["prev_ids"], O)
            logits, mem = self.calibration_decoding(x, st["head"], [st["mem"][0].nchw() if has[0] else None,
                                                                    st["mem"][1].nchw() if has[1] else None], low)
            probs, label, conf = self._upsample_softmax(logits, O, h, w, Hp, Wp)
            return logits, mem, probs, label, conf

        # ---- segment F: bank-dependent (global matching, k-means proxies, bank heads)
        if ix["rows"] == 0 or not self.segf_graph:
            front()                                          # degenerate empty bank / AOCB200_SEGF_GRAPH=0: plain launches
        else:
            segF = st["segF"]
            if segF is None or segF["version"] != ix["version"] or segF["emb"] is not emb:
                if segF is None:
                    front()                                  # first use: warm-up (workspaces)
                else:                                        # workspaces must not be (re)allocated inside a capture
                    self.ws("gm_tc", self.L.global_match_tc_workspace_bytes(hw, ix["rows"]))
                    self.ws("kmeans", self.L.kmeans_workspace_bytes(ix["maxrows"], O, self.kmax))
                    self.ws("headpool", self.L.head_pool_workspace_bytes(max(ix["total"], hw)))
                gF, _ = self._capture(front)
                segF = dict(graph=gF, version=ix["version"], emb=emb)
                st["segF"] = segF
            segF["graph"].replay()
        # ---- segment C: everything after the bank
        keyC = (Hp, Wp, has[0], has[1], id(emb), float(self.unc_ratio), bool(self.want_probs))
        segC = st["segC"].get(keyC)
        if segC is None:
            back()                                           # warm-up
            gC, out = self._capture(back)
            segC = dict(graph=gC, out=out)
            st["segC"][keyC] = segC
        segC["graph"].replay()
        logits, mem, probs, label, conf = segC["out"]
        self.last_logits = logits.view(1, O, h, w)
        self.last_label, self.last_conf_label = label, conf      # static buffers of the graph: clone to keep
        cl = torch.channels_last
        return (None if probs is None else probs.clone()), emb_out, \
            [[mem[0].nchw().clone(memory_format=cl), mem[1].nchw().clone(memory_format=cl)]]

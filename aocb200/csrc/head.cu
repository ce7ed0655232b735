// Output head: per-object dynamic 1x1 conv logits (decoding_module.py:151-160), background-logit augmentation
// (decoding_module.py:213-225) and the final bilinear upsample (align_corners=True) + softmax over object slots
// (aocnet.py:100-107), fused with argmax.
#include "common.cuh"

namespace aoc {

// x: [O][HW][C] (NHWC), wfg/wbg: [O][C+1] (last entry = bias).  fg/bg: [O][HW].  One warp per (o, pixel).
__global__ void __launch_bounds__(256) dyn_logit_kernel(const float* __restrict__ x, const float* __restrict__ wfg,
                                                         const float* __restrict__ wbg, float* __restrict__ fg,
                                                         float* __restrict__ bg, int O, int HW, int C, int ldx) {
    int lane = threadIdx.x & 31;
    long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    int C4 = C >> 2;
    for (long long i = warp; i < (long long)O * HW; i += nw) {
        int o = (int)(i / HW);
        const float* row = x + (size_t)i * ldx;
        float a = 0.f, b = 0.f;
        for (int k = lane; k < C4; k += 32) {
            float4 v = ldg4(row + k * 4);
            const float* wf = wfg + (size_t)o * (C + 1) + k * 4;
            const float* wb = wbg + (size_t)o * (C + 1) + k * 4;
            a = fmaf(v.x, __ldg(wf + 0), a); a = fmaf(v.y, __ldg(wf + 1), a);
            a = fmaf(v.z, __ldg(wf + 2), a); a = fmaf(v.w, __ldg(wf + 3), a);
            b = fmaf(v.x, __ldg(wb + 0), b); b = fmaf(v.y, __ldg(wb + 1), b);
            b = fmaf(v.z, __ldg(wb + 2), b); b = fmaf(v.w, __ldg(wb + 3), b);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) {
            fg[i] = a + __ldg(wfg + (size_t)o * (C + 1) + C);
            bg[i] = b + __ldg(wbg + (size_t)o * (C + 1) + C);
        }
    }
}

// logits[o][p] = fg[o][p] (+ min_{o'>=1} bg[o'][p] for o == 0 when O > 1)
__global__ void augment_bg_kernel(const float* __restrict__ fg, const float* __restrict__ bg, float* __restrict__ logits,
                                  int O, int HW) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    float add = 0.f;
    if (O > 1) {
        float m = INFINITY;
        for (int o = 1; o < O; ++o) m = fminf(m, bg[(size_t)o * HW + p]);
        add = m;
    }
    logits[p] = fg[p] + add;
    for (int o = 1; o < O; ++o) logits[(size_t)o * HW + p] = fg[(size_t)o * HW + p];
}

// logits [O][h][w] -> probs [O][H][W] (softmax over O after bilinear align_corners upsample), label[H][W] = argmax
__global__ void upsample_softmax_kernel(const float* __restrict__ logits, float* __restrict__ probs,
                                        uint8_t* __restrict__ label, int O, int h, int w, int H, int W, float sh,
                                        float sw) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    int yo = i / W, xo = i - yo * W;
    float ry = sh * (float)yo, rx = sw * (float)xo;
    int y0 = min((int)ry, h - 1), x0 = min((int)rx, w - 1);
    int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    float ly1 = fminf(fmaxf(ry - (float)y0, 0.f), 1.f), lx1 = fminf(fmaxf(rx - (float)x0, 0.f), 1.f);
    float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    float v[AOC_MAX_OBJECTS];
    float mx = -INFINITY;
    int am = 0;
    for (int o = 0; o < O; ++o) {
        const float* L = logits + (size_t)o * h * w;
        float t = ly0 * (lx0 * __ldg(L + y0 * w + x0) + lx1 * __ldg(L + y0 * w + x1)) +
                  ly1 * (lx0 * __ldg(L + y1 * w + x0) + lx1 * __ldg(L + y1 * w + x1));
        v[o] = t;
        if (t > mx) { mx = t; am = o; }
    }
    float s = 0.f;
    for (int o = 0; o < O; ++o) { v[o] = expf(v[o] - mx); s += v[o]; }
    float inv = 1.0f / s;
    for (int o = 0; o < O; ++o) probs[(size_t)o * H * W + i] = v[o] * inv;
    if (label) label[i] = (uint8_t)am;
}

// The eval loop's per-frame label bookkeeping fused behind the upsample + softmax (eval_manager_mm.py:252-270 label
// filter, :318-320 argmax, :339-349 uncertainty filter; shannon_entropy.py:10-13):
//   probs[o] = softmax_o(upsampled logits), unfiltered: what AOCNet.forward_for_eval returns (optional output)
//   p_o      = probs[o] for labels seen in a ground-truth frame so far, 0 otherwise
//   label    = argmax_o p_o (lowest index wins ties, as torch.argmax)
//   ent      = -sum_{o seen} p_o * log(p_o + 1e-6)
//   conf     = ent > unc_ratio ? 125 : label      (125 = "uncertain": matches no object slot in the memory bank)
// exist: device word, bit o set <=> label o has been seen (null = all); a device word so that a captured graph follows
// the caller's set without re-capture.
__global__ void upsample_softmax_label_kernel(const float* __restrict__ logits, float* __restrict__ probs,
                                              uint8_t* __restrict__ label, uint8_t* __restrict__ conf,
                                              float* __restrict__ entropy, const int* __restrict__ exist,
                                              float unc_ratio, int O, int h, int w, int H, int W, float sh, float sw) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const unsigned ex = exist ? (unsigned)__ldg(exist) : 0xffffffffu;
    int yo = i / W, xo = i - yo * W;
    float ry = sh * (float)yo, rx = sw * (float)xo;
    int y0 = min((int)ry, h - 1), x0 = min((int)rx, w - 1);
    int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    float ly1 = fminf(fmaxf(ry - (float)y0, 0.f), 1.f), lx1 = fminf(fmaxf(rx - (float)x0, 0.f), 1.f);
    float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    float v[AOC_MAX_OBJECTS];
    float mx = -INFINITY;
    for (int o = 0; o < O; ++o) {
        const float* L = logits + (size_t)o * h * w;
        float t = ly0 * (lx0 * __ldg(L + y0 * w + x0) + lx1 * __ldg(L + y0 * w + x1)) +
                  ly1 * (lx0 * __ldg(L + y1 * w + x0) + lx1 * __ldg(L + y1 * w + x1));
        v[o] = t;
        mx = fmaxf(mx, t);
    }
    float s = 0.f;
    for (int o = 0; o < O; ++o) { v[o] = expf(v[o] - mx); s += v[o]; }
    float inv = 1.0f / s;
    float best = -1.f, ent = 0.f;
    int am = 0;
    for (int o = 0; o < O; ++o) {
        const bool seen = (ex >> o) & 1u;
        const float pa = v[o] * inv;
        if (probs) probs[(size_t)o * H * W + i] = pa;
        const float pr = seen ? pa : 0.f;
        if (pr > best) { best = pr; am = o; }
        if (seen) ent += pr * logf(pr + 1e-6f);
    }
    ent = -ent;
    label[i] = (uint8_t)am;
    if (conf) conf[i] = ent > unc_ratio ? (uint8_t)125 : (uint8_t)am;
    if (entropy) entropy[i] = ent;
}

}  // namespace aoc

using namespace aoc;

extern "C" int aoc_dyn_logits_f32(const float* x, const float* wfg, const float* wbg, float* fg, float* bg,
                                  float* logits, int O, int HW, int C, int ldx, cudaStream_t stream) {
    AOC_CHECK_ARG(x && wfg && wbg && fg && bg && logits, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= AOC_MAX_OBJECTS && C % 4 == 0 && ldx % 4 == 0, "bad dims");
    long long warps = (long long)O * HW;
    int blocks = (int)((warps + 7) / 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    dyn_logit_kernel<<<blocks, 256, 0, stream>>>(x, wfg, wbg, fg, bg, O, HW, C, ldx);
    augment_bg_kernel<<<cdiv(HW, 256), 256, 0, stream>>>(fg, bg, logits, O, HW);
    return launch_status("aoc_dyn_logits_f32");
}

extern "C" int aoc_upsample_softmax_f32(const float* logits, float* probs, uint8_t* label, int O, int h, int w,
                                        int H, int W, cudaStream_t stream) {
    AOC_CHECK_ARG(logits && probs, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= AOC_MAX_OBJECTS && h > 0 && w > 0 && H > 0 && W > 0, "bad dims");
    float sh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
    float sw = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    upsample_softmax_kernel<<<cdiv((long long)H * W, 256), 256, 0, stream>>>(logits, probs, label, O, h, w, H, W, sh,
                                                                           sw);
    return launch_status("aoc_upsample_softmax_f32");
}

extern "C" int aoc_upsample_softmax_label_f32(const float* logits, float* probs, uint8_t* label, uint8_t* conf_label,
                                              float* entropy, const int* exist_bits, float unc_ratio, int O, int h,
                                              int w, int H, int W, cudaStream_t stream) {
    AOC_CHECK_ARG(logits && label, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= AOC_MAX_OBJECTS && h > 0 && w > 0 && H > 0 && W > 0, "bad dims");
    float sh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
    float sw = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    upsample_softmax_label_kernel<<<cdiv((long long)H * W, 256), 256, 0, stream>>>(
        logits, probs, label, conf_label, entropy, exist_bits, unc_ratio, O, h, w, H, W, sh, sw);
    return launch_status("aoc_upsample_softmax_label_f32");
}

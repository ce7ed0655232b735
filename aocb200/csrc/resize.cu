// Layout conversion and resampling kernels (NHWC fp32).
// Replaces F.interpolate call sites: bilinear align_corners=True (deeplab/aspp.py:72, deeplab/decoder.py:35,
// layers/aspp.py:66, matching.py:2729-2732,2849, aocnet.py:103), bicubic align_corners=True
// (decoding_module.py:163), nearest for label maps (aocnet.py:128-135, matching.py:2805).
#include "common.cuh"

namespace aoc {

// image [3,H,W] (NCHW, N=1) -> [H*W,4] with a zero 4th channel (so the stem conv takes the float4 path)
__global__ void nchw3_to_nhwc4_kernel(const float* __restrict__ x, float* __restrict__ y, int HW) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    float4 v = make_float4(__ldg(x + p), __ldg(x + HW + p), __ldg(x + 2 * (size_t)HW + p), 0.f);
    *reinterpret_cast<float4*>(y + (size_t)p * 4) = v;
}

// image [3,H,W] -> space-to-depth image [H2 + 1][W2 + 1][16], H2 = ceil(H / 2): pixel (i, j) holds the 2 x 2 block of source
// pixels (2 (i - 1) + py, 2 (j - 1) + px) as channels (py * 2 + px) * 4 + c (c = 3 and everything outside the image: zero;
// row 0 and column 0 are zero).  A 7 x 7 / stride-2 / pad-3 convolution over the image is a 4 x 4 / stride-1 / pad-1
// convolution over this tensor (tap a' of the 4 covers source taps r = 2 a' + py - 1): the stem runs 16 stages of 16 real
// channels instead of 49 stages of 4 real + 12 zero channels.
__global__ void image_to_s2d16_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int H2p, int W2p) {
    AOC_PDL_TRIGGER();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;           // (output pixel, quarter = source pixel of the block)
    if (idx >= H2p * W2p * 4) return;
    const int q = idx & 3, pix = idx >> 2;
    const int i = pix / W2p, j = pix - i * W2p;
    const int u = 2 * (i - 1) + (q >> 1), v = 2 * (j - 1) + (q & 1);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i > 0 && j > 0 && u < H && v < W) {
        const size_t HW = (size_t)H * W, at = (size_t)u * W + v;
        o = make_float4(__ldg(x + at), __ldg(x + HW + at), __ldg(x + 2 * HW + at), 0.f);
    }
    *reinterpret_cast<float4*>(y + (size_t)idx * 4) = o;
}

// generic [N,C,HW] -> [N,HW,C] (smem-tiled transpose)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW, int ldy) {
    __shared__ float t[32][33];
    int n = blockIdx.z;
    int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) t[i][threadIdx.x] = x[((size_t)n * C + c) * HW + p];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int p = p0 + i, c = c0 + threadIdx.x;
        if (c < C && p < HW) y[((size_t)n * HW + p) * ldy + c] = t[threadIdx.x][i];
    }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW, int ldx) {
    __shared__ float t[32][33];
    int n = blockIdx.z;
    int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int p = p0 + i, c = c0 + threadIdx.x;
        if (c < C && p < HW) t[i][threadIdx.x] = x[((size_t)n * HW + p) * ldx + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) y[((size_t)n * C + c) * HW + p] = t[threadIdx.x][i];
    }
}

__device__ __forceinline__ void lin_coords(int dst, float scale, int in, int& i0, int& i1, float& l0, float& l1) {
    float real = scale * (float)dst;          // align_corners=True: src = dst*(in-1)/(out-1)
    i0 = min((int)real, in - 1);
    i1 = i0 + ((i0 < in - 1) ? 1 : 0);
    l1 = fminf(fmaxf(real - (float)i0, 0.f), 1.f);
    l0 = 1.f - l1;
}

// Bilinear, align_corners=True.  If ids != nullptr the source is a lookup: src[n,pix,:] = table[ids[n,pix]] (zero
// row when ids >= n_table) -- used for `seq_prev_frame_embedding_inst` (aocnet.py:325) without materialising it.
__global__ void resize_bilinear_kernel(const float* __restrict__ x, const uint8_t* __restrict__ ids,
                                       const float* __restrict__ table, int n_table, float* __restrict__ y, int N,
                                       int Hi, int Wi, int Ho, int Wo, int C, int ldx, int ldy, float sh, float sw) {
    AOC_PDL_TRIGGER();
    int C4 = C >> 2;
    long long total = (long long)N * Ho * Wo * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C4) * 4;
        long long pix = i / C4;
        int xo = (int)(pix % Wo);
        int yo = (int)((pix / Wo) % Ho);
        int n = (int)(pix / ((long long)Wo * Ho));
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        lin_coords(yo, sh, Hi, y0, y1, ly0, ly1);
        lin_coords(xo, sw, Wi, x0, x1, lx0, lx1);
        float4 v00, v01, v10, v11;
        size_t b = (size_t)n * Hi * Wi;
        if (ids) {
            auto fetch = [&](int yy, int xx) {
                int id = ids[b + (size_t)yy * Wi + xx];
                return id < n_table ? ldg4(table + (size_t)id * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            v00 = fetch(y0, x0); v01 = fetch(y0, x1); v10 = fetch(y1, x0); v11 = fetch(y1, x1);
        } else {
            v00 = ldg4(x + (b + (size_t)y0 * Wi + x0) * ldx + c);
            v01 = ldg4(x + (b + (size_t)y0 * Wi + x1) * ldx + c);
            v10 = ldg4(x + (b + (size_t)y1 * Wi + x0) * ldx + c);
            v11 = ldg4(x + (b + (size_t)y1 * Wi + x1) * ldx + c);
        }
        float4 o;
        o.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
        o.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
        o.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
        o.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
        *reinterpret_cast<float4*>(y + (size_t)pix * ldy + c) = o;
    }
}

__device__ __forceinline__ void cubic_coeffs(float t, float w[4]) {
    const float A = -0.75f;
    float x = t + 1.f;
    w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
    x = t;
    w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    x = 1.f - t;
    w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    x = 2.f - t;
    w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

// Bicubic (A=-0.75), align_corners=True, border indices clamped.
__global__ void resize_bicubic_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int Hi, int Wi, int Ho,
                                      int Wo, int C, int ldx, int ldy, float sh, float sw) {
    int C4 = C >> 2;
    long long total = (long long)N * Ho * Wo * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C4) * 4;
        long long pix = i / C4;
        int xo = (int)(pix % Wo);
        int yo = (int)((pix / Wo) % Ho);
        int n = (int)(pix / ((long long)Wo * Ho));
        float ry = sh * (float)yo, rx = sw * (float)xo;
        int iy = (int)floorf(ry), ix = (int)floorf(rx);
        float wy[4], wx[4];
        cubic_coeffs(ry - (float)iy, wy);
        cubic_coeffs(rx - (float)ix, wx);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        size_t b = (size_t)n * Hi * Wi;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int yy = min(max(iy - 1 + j, 0), Hi - 1);
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int xx = min(max(ix - 1 + k, 0), Wi - 1);
                float4 v = ldg4(x + (b + (size_t)yy * Wi + xx) * ldx + c);
                r.x = fmaf(wx[k], v.x, r.x); r.y = fmaf(wx[k], v.y, r.y);
                r.z = fmaf(wx[k], v.z, r.z); r.w = fmaf(wx[k], v.w, r.w);
            }
            o.x = fmaf(wy[j], r.x, o.x); o.y = fmaf(wy[j], r.y, o.y);
            o.z = fmaf(wy[j], r.z, o.z); o.w = fmaf(wy[j], r.w, o.w);
        }
        *reinterpret_cast<float4*>(y + (size_t)pix * ldy + c) = o;
    }
}

// The decoder's only bicubic resize is an exact 2x upsampling with align_corners (61 x 107 -> 121 x 213: Ho = 2 Hi - 1), where
// the four outputs (2i + a, 2j + b), a, b in {0, 1}, read the SAME 4 x 4 input patch (rows i-1..i+2, columns j-1..j+2).  One
// thread computes the 2 x 2 block from one patch: 16 loads per four outputs instead of 16 per output (the general kernel
// was bound by the L1 / LSU, 183 us for 158 MB of output); per output the arithmetic -- the same coefficients, the same fma
// order, minus the terms whose weight is exactly zero -- is that of resize_bicubic_kernel, so the results are bit-identical.
__global__ void __launch_bounds__(256) resize_bicubic2x_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int Hi,
                                                               int Wi, int Ho, int Wo, int C, int ldx, int ldy, float sh, float sw) {
    AOC_PDL_TRIGGER();
    const int C4 = C >> 2;
    const int Hb = (Ho + 1) >> 1, Wb = (Wo + 1) >> 1;                  // 2 x 2 output blocks
    const long long total = (long long)N * Hb * Wb * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        long long blk = i / C4;
        const int bj = (int)(blk % Wb);
        const int bi = (int)((blk / Wb) % Hb);
        const int n = (int)(blk / ((long long)Wb * Hb));
        const size_t b = (size_t)n * Hi * Wi;
        float4 v[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int yy = min(max(bi - 1 + j, 0), Hi - 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xx = min(max(bj - 1 + k, 0), Wi - 1);
                v[j][k] = ldg4(x + (b + (size_t)yy * Wi + xx) * ldx + c);
            }
        }
        // With sh = sw = 1/2 exactly (the launcher's condition) the fractional position of an output is 0 (even index) or
        // 1/2 (odd index), and cubic_coeffs() evaluates -- exactly, every intermediate is a small dyadic number -- to
        // {0, 1, 0, 0} and {-3/32, 19/32, 19/32, -3/32}.  A zero weight leaves the fma chain of the general kernel unchanged
        // (fma(0, v, r) = r for finite v) and a unit weight copies, so: (even, even) is the centre sample itself,
        // (even, odd) / (odd, even) are ONE four-tap chain along the row / column through the centre, and only (odd, odd)
        // needs the full 4 x 4 patch -- 112 instead of 320 fmas per thread, in the same order: bit-identical for finite inputs.
        const float W0 = -0.09375f, W1 = 0.59375f;
        const float wq[4] = {W0, W1, W1, W0};
        const size_t orow = ((size_t)n * Ho + 2 * bi) * Wo + 2 * bj;
        const bool has_x = 2 * bj + 1 < Wo, has_y = 2 * bi + 1 < Ho;
        *reinterpret_cast<float4*>(y + orow * ldy + c) = v[1][1];
        if (has_x) {
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                r.x = fmaf(wq[k], v[1][k].x, r.x); r.y = fmaf(wq[k], v[1][k].y, r.y);
                r.z = fmaf(wq[k], v[1][k].z, r.z); r.w = fmaf(wq[k], v[1][k].w, r.w);
            }
            *reinterpret_cast<float4*>(y + (orow + 1) * ldy + c) = r;
        }
        if (has_y) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o.x = fmaf(wq[j], v[j][1].x, o.x); o.y = fmaf(wq[j], v[j][1].y, o.y);
                o.z = fmaf(wq[j], v[j][1].z, o.z); o.w = fmaf(wq[j], v[j][1].w, o.w);
            }
            *reinterpret_cast<float4*>(y + (orow + Wo) * ldy + c) = o;
        }
        if (has_x && has_y) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    r.x = fmaf(wq[k], v[j][k].x, r.x); r.y = fmaf(wq[k], v[j][k].y, r.y);
                    r.z = fmaf(wq[k], v[j][k].z, r.z); r.w = fmaf(wq[k], v[j][k].w, r.w);
                }
                o.x = fmaf(wq[j], r.x, o.x); o.y = fmaf(wq[j], r.y, o.y);
                o.z = fmaf(wq[j], r.z, o.z); o.w = fmaf(wq[j], r.w, o.w);
            }
            *reinterpret_cast<float4*>(y + (orow + Wo + 1) * ldy + c) = o;
        }
    }
}

// Input edge of the eval loop (dataloaders/custom_transforms.py:387-463 MultiRestrictSize + :465-487 MultiToTensor): uint8
// HWC frame -> cv2.resize(..., INTER_CUBIC) of the float image (A = -0.75, half-pixel centres src = (dst + 0.5) * scale - 0.5,
// source indices clamped, horizontal pass then vertical pass) -> optional mirror (tmp[:, ::-1]) -> /255, -mean, /std ->
// CHW float32.  One thread per output pixel, the three channels of a tap from one 3-byte read; with Ho == H and Wo == W the
// transform does not resize (the reference hands the sample on unchanged) and this is the normalisation alone.
__global__ void __launch_bounds__(256) prepare_frame_kernel(const uint8_t* __restrict__ img, int H, int W, int Ho, int Wo,
                                                            int flip, float m0, float m1, float m2, float s0, float s1,
                                                            float s2, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ho * Wo) return;
    const int yo = i / Wo, xo = i - yo * Wo;
    const int xs = flip ? Wo - 1 - xo : xo;                       // the mirror is applied AFTER the resize
    float r[3];
    if (Ho == H && Wo == W) {
        const uint8_t* px = img + ((size_t)yo * W + xs) * 3;
        r[0] = (float)px[0]; r[1] = (float)px[1]; r[2] = (float)px[2];
    } else {
        const float sy = (float)((double)H / (double)Ho), sx = (float)((double)W / (double)Wo);
        float fy = ((float)yo + 0.5f) * sy - 0.5f, fx = ((float)xs + 0.5f) * sx - 0.5f;
        const int iy = (int)floorf(fy), ix = (int)floorf(fx);
        float wy[4], wx[4];
        cubic_coeffs(fy - (float)iy, wy);
        cubic_coeffs(fx - (float)ix, wx);
        r[0] = r[1] = r[2] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int yy = min(max(iy - 1 + j, 0), H - 1);
            float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xx = min(max(ix - 1 + k, 0), W - 1);
                const uint8_t* px = img + ((size_t)yy * W + xx) * 3;
                h0 += (float)px[0] * wx[k]; h1 += (float)px[1] * wx[k]; h2 += (float)px[2] * wx[k];
            }
            r[0] += h0 * wy[j]; r[1] += h1 * wy[j]; r[2] += h2 * wy[j];
        }
    }
    const size_t plane = (size_t)Ho * Wo;
    out[i] = (r[0] / 255.f - m0) / s0;                            // the transform's own order: /255, -mean, /std (IEEE divisions)
    out[plane + i] = (r[1] / 255.f - m1) / s1;
    out[2 * plane + i] = (r[2] / 255.f - m2) / s2;
}

// Nearest resize of a uint8 label map, PyTorch 'nearest' rule: src = min(floor(dst * (in/out)), in-1) in float.
__global__ void resize_nearest_u8_kernel(const uint8_t* __restrict__ x, uint8_t* __restrict__ y, int Hi, int Wi,
                                         int Ho, int Wo, float sh, float sw) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ho * Wo) return;
    int yo = i / Wo, xo = i - yo * Wo;
    int yi = (Ho == Hi) ? yo : min((int)floorf((float)yo * sh), Hi - 1);
    int xi = (Wo == Wi) ? xo : min((int)floorf((float)xo * sw), Wi - 1);
    y[i] = x[(size_t)yi * Wi + xi];
}

// the same for an int64 label map (what torch.argmax hands the reference loop), values clamped to 0..255
__global__ void resize_nearest_i64_kernel(const long long* __restrict__ x, uint8_t* __restrict__ y, int Hi, int Wi,
                                          int Ho, int Wo, float sh, float sw) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ho * Wo) return;
    int yo = i / Wo, xo = i - yo * Wo;
    int yi = (Ho == Hi) ? yo : min((int)floorf((float)yo * sh), Hi - 1);
    int xi = (Wo == Wi) ? xo : min((int)floorf((float)xo * sw), Wi - 1);
    const long long v = x[(size_t)yi * Wi + xi];
    y[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, uint32_t v, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}

// out = a (op) b over n floats; op 0: a + b, 1: a * b   (per-(sample, channel) coefficient vectors: a few thousand values)
__global__ void vec_op_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n,
                              int op) {
    AOC_PDL_TRIGGER();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = op == 0 ? a[i] + b[i] : a[i] * b[i];
}

static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

}  // namespace aoc

using namespace aoc;

extern "C" int aoc_image_to_nhwc4_f32(const float* x, float* y, int H, int W, cudaStream_t stream) {
    AOC_CHECK_ARG(x && y && H > 0 && W > 0, "bad args");
    nchw3_to_nhwc4_kernel<<<cdiv((long long)H * W, 256), 256, 0, stream>>>(x, y, H * W);
    return launch_status("aoc_image_to_nhwc4_f32");
}

extern "C" int aoc_image_to_s2d16_f32(const float* x, float* y, int H, int W, cudaStream_t stream) {
    AOC_CHECK_ARG(x && y && H > 0 && W > 0, "bad args");
    const int H2p = (H + 1) / 2 + 1, W2p = (W + 1) / 2 + 1;
    image_to_s2d16_kernel<<<cdiv((long long)H2p * W2p * 4, 256), 256, 0, stream>>>(x, y, H, W, H2p, W2p);
    return launch_status("aoc_image_to_s2d16_f32");
}

extern "C" int aoc_nchw_to_nhwc_f32(const float* x, float* y, int N, int C, int HW, int ldy, cudaStream_t stream) {
    AOC_CHECK_ARG(x && y && N > 0 && C > 0 && HW > 0, "bad args");
    dim3 g(cdiv(HW, 32), cdiv(C, 32), N), b(32, 8);
    nchw_to_nhwc_kernel<<<g, b, 0, stream>>>(x, y, C, HW, ldy);
    return launch_status("aoc_nchw_to_nhwc_f32");
}

extern "C" int aoc_nhwc_to_nchw_f32(const float* x, float* y, int N, int C, int HW, int ldx, cudaStream_t stream) {
    AOC_CHECK_ARG(x && y && N > 0 && C > 0 && HW > 0, "bad args");
    dim3 g(cdiv(HW, 32), cdiv(C, 32), N), b(32, 8);
    nhwc_to_nchw_kernel<<<g, b, 0, stream>>>(x, y, C, HW, ldx);
    return launch_status("aoc_nhwc_to_nchw_f32");
}

extern "C" int aoc_resize_bilinear_nhwc_f32(const float* x, const uint8_t* ids, const float* table, int n_table,
                                            float* y, int N, int Hi, int Wi, int Ho, int Wo, int C, int ldx, int ldy,
                                            cudaStream_t stream) {
    AOC_CHECK_ARG((x || (ids && table)) && y, "null pointer");
    AOC_CHECK_ARG(C % 4 == 0 && ldy % 4 == 0 && (ids || ldx % 4 == 0), "C/ld must be multiples of 4");
    long long total = (long long)N * Ho * Wo * (C / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    resize_bilinear_kernel<<<blocks, 256, 0, stream>>>(x, ids, table, n_table, y, N, Hi, Wi, Ho, Wo, C, ldx, ldy,
                                                       ac_scale(Hi, Ho), ac_scale(Wi, Wo));
    return launch_status("aoc_resize_bilinear_nhwc_f32");
}

extern "C" int aoc_resize_bicubic_nhwc_f32(const float* x, float* y, int N, int Hi, int Wi, int Ho, int Wo, int C,
                                           int ldx, int ldy, cudaStream_t stream) {
    AOC_CHECK_ARG(x && y, "null pointer");
    AOC_CHECK_ARG(C % 4 == 0 && ldy % 4 == 0 && ldx % 4 == 0, "C/ld must be multiples of 4");
    long long total = (long long)N * Ho * Wo * (C / 4);
    if (Ho == 2 * Hi - 1 && Wo == 2 * Wi - 1 && Hi > 1 && Wi > 1) {        // exact 2x: one thread per 2 x 2 output block
        total = (long long)N * ((Ho + 1) / 2) * ((Wo + 1) / 2) * (C / 4);
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 32) blocks = 148 * 32;
        resize_bicubic2x_kernel<<<blocks, 256, 0, stream>>>(x, y, N, Hi, Wi, Ho, Wo, C, ldx, ldy, ac_scale(Hi, Ho), ac_scale(Wi, Wo));
        return launch_status("aoc_resize_bicubic_nhwc_f32");
    }
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    resize_bicubic_kernel<<<blocks, 256, 0, stream>>>(x, y, N, Hi, Wi, Ho, Wo, C, ldx, ldy, ac_scale(Hi, Ho),
                                                      ac_scale(Wi, Wo));
    return launch_status("aoc_resize_bicubic_nhwc_f32");
}

extern "C" int aoc_prepare_frame_u8(const uint8_t* img_hwc, int H, int W, int Ho, int Wo, int flip, const float* mean3,
                                    const float* std3, float* out_chw, cudaStream_t stream) {
    AOC_CHECK_ARG(img_hwc && out_chw && mean3 && std3, "null pointer");
    AOC_CHECK_ARG(H > 0 && W > 0 && Ho > 0 && Wo > 0, "bad dims");
    AOC_CHECK_ARG(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "zero std");
    prepare_frame_kernel<<<cdiv((long long)Ho * Wo, 256), 256, 0, stream>>>(img_hwc, H, W, Ho, Wo, flip, mean3[0], mean3[1],
                                                                          mean3[2], std3[0], std3[1], std3[2], out_chw);
    return launch_status("aoc_prepare_frame_u8");
}

extern "C" int aoc_resize_nearest_u8(const uint8_t* x, uint8_t* y, int Hi, int Wi, int Ho, int Wo,
                                     cudaStream_t stream) {
    AOC_CHECK_ARG(x && y && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bad args");
    resize_nearest_u8_kernel<<<cdiv((long long)Ho * Wo, 256), 256, 0, stream>>>(x, y, Hi, Wi, Ho, Wo,
                                                                             (float)Hi / (float)Ho,
                                                                             (float)Wi / (float)Wo);
    return launch_status("aoc_resize_nearest_u8");
}

extern "C" int aoc_resize_nearest_i64(const long long* x, uint8_t* y, int Hi, int Wi, int Ho, int Wo,
                                      cudaStream_t stream) {
    AOC_CHECK_ARG(x && y && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bad args");
    resize_nearest_i64_kernel<<<cdiv((long long)Ho * Wo, 256), 256, 0, stream>>>(x, y, Hi, Wi, Ho, Wo,
                                                                              (float)Hi / (float)Ho,
                                                                              (float)Wi / (float)Wo);
    return launch_status("aoc_resize_nearest_i64");
}

extern "C" int aoc_fill_u32(void* p, unsigned int value, long long n_words, cudaStream_t stream) {
    AOC_CHECK_ARG(p && n_words > 0 && (((uintptr_t)p) & 3) == 0, "bad args");
    int blocks = (int)((n_words + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    fill_u32_kernel<<<blocks, 256, 0, stream>>>((uint32_t*)p, value, n_words);
    return launch_status("aoc_fill_u32");
}

extern "C" int aoc_vec_op_f32(const float* a, const float* b, float* out, int n, int op, cudaStream_t stream) {
    AOC_CHECK_ARG(a && b && out && n > 0 && (op == 0 || op == 1), "bad args");
    vec_op_kernel<<<cdiv(n, 256), 256, 0, stream>>>(a, b, out, n, op);
    return launch_status("aoc_vec_op_f32");
}

namespace aoc {
__global__ void copy_channels_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int C, int ldx,
                                     int ldy) {
    AOC_PDL_TRIGGER();
    int C4 = C >> 2;
    long long total = rows * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C4) * 4;
        long long r = i / C4;
        *reinterpret_cast<float4*>(y + (size_t)r * ldy + c) = ldg4(x + (size_t)r * ldx + c);
    }
}
}  // namespace aoc

// y[r, 0:C] = x[r, 0:C] with independent row strides (channel-slice copy into / out of concat buffers)
extern "C" int aoc_copy_channels_f32(const float* x, float* y, long long rows, int C, int ldx, int ldy,
                                     cudaStream_t stream) {
    AOC_CHECK_ARG(x && y && rows > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "bad args");
    long long total = rows * (C / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    aoc::copy_channels_kernel<<<blocks, 256, 0, stream>>>(x, y, rows, C, ldx, ldy);
    return aoc::launch_status("aoc_copy_channels_f32");
}

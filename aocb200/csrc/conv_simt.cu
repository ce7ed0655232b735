// fp32 SIMT implicit-GEMM convolution (NHWC activations, [Cout][kh][kw][Cin] weights), depthwise 3x3,
// 3x3/s2 max-pool.  This is the exact-fp32 fallback/verification path for the tcgen05 kernels in
// umma_conv.cu and the path used for shapes the tensor-core kernel does not take (Cin % 4 != 0).
//
// Replaces the reference's nn.Conv2d + FrozenBatchNorm2d (+ReLU, +residual) call sites:
//   networks/deeplab/backbone/resnet.py:23-42,108-123, networks/deeplab/aspp.py:62-74,
//   networks/deeplab/decoder.py:32-41, networks/layers/gct.py:68-91, networks/layers/aspp.py:57-70,
//   networks/aoc/decoding_module.py:162-190,228-240, networks/aoc/aocnet.py:19-25.
#include "common.cuh"

namespace aoc {

struct ConvP {
    const float* x; const float* w; const float* bias; const float* res; const float* in_scale; float* y;
    int N, H, W, Cin, ldx, Ho, Wo, Cout, ldy, ldres, kh, kw, stride, pad, dil, relu;
    int M, K;
    int vec_in, vec_out;
};

constexpr int CONV_BK = 16;

template <int BM, int BN>
__global__ void __launch_bounds__(256) conv_igemm_kernel(ConvP p) {
    constexpr int BK = CONV_BK;
    constexpr int GM = BM / 64, GN = BN / 64;       // float4 groups per thread along M / N
    constexpr int AG = BM * BK / 4 / 256;           // float4 granules of A per thread per k-tile
    constexpr int BG = BN * BK / 4 / 256;
    constexpr int LDA = BM + 4, LDB = BN + 4;
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][LDB];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int HoWo = p.Ho * p.Wo;

    // ---- per-thread A granule bookkeeping (rows fixed across k-tiles) ----
    int a_row[AG], a_kg[AG], a_hi0[AG], a_wi0[AG], a_n[AG];
    bool a_ok[AG];
#pragma unroll
    for (int g = 0; g < AG; ++g) {
        int gi = tid + g * 256;
        a_kg[g] = gi & 3;
        a_row[g] = gi >> 2;
        int m = m0 + a_row[g];
        a_ok[g] = m < p.M;
        int mm = a_ok[g] ? m : 0;
        int n = mm / HoWo, r = mm - n * HoWo;
        int ho = r / p.Wo, wo = r - ho * p.Wo;
        a_n[g] = n;
        a_hi0[g] = ho * p.stride - p.pad;
        a_wi0[g] = wo * p.stride - p.pad;
    }
    int b_row[BG], b_kg[BG];
#pragma unroll
    for (int g = 0; g < BG; ++g) {
        int gi = tid + g * 256;
        b_kg[g] = gi & 3;
        b_row[g] = gi >> 2;
    }

    float acc[GM * 4][GN * 4];
#pragma unroll
    for (int i = 0; i < GM * 4; ++i)
#pragma unroll
        for (int j = 0; j < GN * 4; ++j) acc[i][j] = 0.f;

    float4 ra[AG], rb[BG];
    const int nkt = (p.K + BK - 1) / BK;

    auto load_tile = [&](int kt) {
#pragma unroll
        for (int g = 0; g < AG; ++g) {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            int k0 = kt * BK + a_kg[g] * 4;
            if (a_ok[g] && k0 < p.K) {
                if (p.vec_in) {
                    int rs = k0 / p.Cin, ci = k0 - rs * p.Cin;
                    int r = rs / p.kw, s = rs - r * p.kw;
                    int hi = a_hi0[g] + r * p.dil, wi = a_wi0[g] + s * p.dil;
                    if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) {
                        const float* src = p.x + ((size_t)(a_n[g] * p.H + hi) * p.W + wi) * p.ldx + ci;
                        float4 t = ldg4(src);
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                        if (p.in_scale) {
                            float4 sc = ldg4(p.in_scale + (size_t)a_n[g] * p.Cin + ci);
                            v[0] *= sc.x; v[1] *= sc.y; v[2] *= sc.z; v[3] *= sc.w;
                        }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        int k = k0 + e;
                        if (k < p.K) {
                            int rs = k / p.Cin, ci = k - rs * p.Cin;
                            int r = rs / p.kw, s = rs - r * p.kw;
                            int hi = a_hi0[g] + r * p.dil, wi = a_wi0[g] + s * p.dil;
                            if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) {
                                float t = __ldg(p.x + ((size_t)(a_n[g] * p.H + hi) * p.W + wi) * p.ldx + ci);
                                if (p.in_scale) t *= __ldg(p.in_scale + (size_t)a_n[g] * p.Cin + ci);
                                v[e] = t;
                            }
                        }
                    }
                }
            }
            ra[g] = make_float4(v[0], v[1], v[2], v[3]);
        }
#pragma unroll
        for (int g = 0; g < BG; ++g) {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            int k0 = kt * BK + b_kg[g] * 4;
            int co = n0 + b_row[g];
            if (co < p.Cout && k0 < p.K) {
                const float* src = p.w + (size_t)co * p.K + k0;
                if (p.vec_in) {
                    float4 t = ldg4(src);
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (k0 + e < p.K) v[e] = __ldg(src + e);
                }
            }
            rb[g] = make_float4(v[0], v[1], v[2], v[3]);
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int g = 0; g < AG; ++g) {
            int k = a_kg[g] * 4, m = a_row[g];
            As[buf][k + 0][m] = ra[g].x; As[buf][k + 1][m] = ra[g].y;
            As[buf][k + 2][m] = ra[g].z; As[buf][k + 3][m] = ra[g].w;
        }
#pragma unroll
        for (int g = 0; g < BG; ++g) {
            int k = b_kg[g] * 4, n = b_row[g];
            Bs[buf][k + 0][n] = rb[g].x; Bs[buf][k + 1][n] = rb[g].y;
            Bs[buf][k + 2][n] = rb[g].z; Bs[buf][k + 3][n] = rb[g].w;
        }
    };

    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        int buf = kt & 1;
        if (kt + 1 < nkt) load_tile(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[GM * 4], b[GN * 4];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                float4 t = *reinterpret_cast<const float4*>(&As[buf][k][g * 64 + ty * 4]);
                a[g * 4 + 0] = t.x; a[g * 4 + 1] = t.y; a[g * 4 + 2] = t.z; a[g * 4 + 3] = t.w;
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                float4 t = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
                b[g * 4 + 0] = t.x; b[g * 4 + 1] = t.y; b[g * 4 + 2] = t.z; b[g * 4 + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < GM * 4; ++i)
#pragma unroll
                for (int j = 0; j < GN * 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nkt) {
            store_tile(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue: bias, residual, ReLU ----
#pragma unroll
    for (int gi = 0; gi < GM; ++gi)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + gi * 64 + ty * 4 + i;
            if (m >= p.M) continue;
#pragma unroll
            for (int gj = 0; gj < GN; ++gj) {
                int co = n0 + gj * 64 + tx * 4;
                if (co >= p.Cout) continue;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = acc[gi * 4 + i][gj * 4 + j];
                    if (co + j < p.Cout) {
                        if (p.bias) v[j] += __ldg(p.bias + co + j);
                        if (p.res) v[j] += __ldg(p.res + (size_t)m * p.ldres + co + j);
                        if (p.relu) v[j] = fmaxf(v[j], 0.f);
                    }
                }
                float* dst = p.y + (size_t)m * p.ldy + co;
                if (p.vec_out && co + 3 < p.Cout) {
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (co + j < p.Cout) dst[j] = v[j];
                }
            }
        }
}

// depthwise 3x3, pad 1, stride 1, with bias (aocnet.py:19 `seperate_conv`).  w: [C][3][3].
__global__ void dwconv3x3_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ bias, float* __restrict__ y,
                                 int N, int H, int W, int C) {
    int C4 = C >> 2;
    long long total = (long long)N * H * W * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c4 = (int)(i % C4);
        long long pix = i / C4;
        int wo = (int)(pix % W);
        int ho = (int)((pix / W) % H);
        int n = (int)(pix / ((long long)W * H));
        int c = c4 * 4;
        float4 acc = bias ? ldg4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            int hi = ho + r - 1;
            if (hi < 0 || hi >= H) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                int wi = wo + s - 1;
                if (wi < 0 || wi >= W) continue;
                float4 v = ldg4(x + ((size_t)(n * H + hi) * W + wi) * C + c);
                acc.x = fmaf(v.x, __ldg(w + (c + 0) * 9 + r * 3 + s), acc.x);
                acc.y = fmaf(v.y, __ldg(w + (c + 1) * 9 + r * 3 + s), acc.y);
                acc.z = fmaf(v.z, __ldg(w + (c + 2) * 9 + r * 3 + s), acc.z);
                acc.w = fmaf(v.w, __ldg(w + (c + 3) * 9 + r * 3 + s), acc.w);
            }
        }
        *reinterpret_cast<float4*>(y + (size_t)pix * C + c) = acc;
    }
}

// 3x3 stride-2 pad-1 max pool (resnet.py:113), NHWC, C % 4 == 0.
__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, float* __restrict__ y,
                                    int N, int H, int W, int C, int Ho, int Wo) {
    AOC_PDL_TRIGGER();
    int C4 = C >> 2;
    long long total = (long long)N * Ho * Wo * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C4) * 4;
        long long pix = i / C4;
        int wo = (int)(pix % Wo);
        int ho = (int)((pix / Wo) % Ho);
        int n = (int)(pix / ((long long)Wo * Ho));
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            int hi = ho * 2 - 1 + r;
            if (hi < 0 || hi >= H) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                int wi = wo * 2 - 1 + s;
                if (wi < 0 || wi >= W) continue;
                float4 v = ldg4(x + ((size_t)(n * H + hi) * W + wi) * C + c);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        *reinterpret_cast<float4*>(y + (size_t)pix * C + c) = m;
    }
}

}  // namespace aoc

using namespace aoc;

extern "C" int aoc_conv2d_nhwc_f32(const float* x, const float* w, const float* bias, const float* residual,
                                   const float* in_scale, float* y, int N, int H, int W, int Cin, int ldx,
                                   int Cout, int ldy, int ldres, int kh, int kw, int stride, int pad, int dil,
                                   int relu, cudaStream_t stream) {
    AOC_CHECK_ARG(x && w && y, "null pointer");
    AOC_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0 && dil > 0,
                  "bad dims");
    ConvP p;
    p.x = x; p.w = w; p.bias = bias; p.res = residual; p.in_scale = in_scale; p.y = y;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.ldx = ldx; p.Cout = Cout; p.ldy = ldy; p.ldres = ldres;
    p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.dil = dil; p.relu = relu;
    p.Ho = (H + 2 * pad - dil * (kh - 1) - 1) / stride + 1;
    p.Wo = (W + 2 * pad - dil * (kw - 1) - 1) / stride + 1;
    AOC_CHECK_ARG(p.Ho > 0 && p.Wo > 0, "empty output");
    p.M = N * p.Ho * p.Wo;
    p.K = kh * kw * Cin;
    p.vec_in = (Cin % 4 == 0) && (ldx % 4 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)w & 15) == 0) &&
               (!in_scale || ((uintptr_t)in_scale & 15) == 0);
    p.vec_out = (ldy % 4 == 0) && (((uintptr_t)y & 15) == 0);
    long long big = (long long)cdiv(p.M, 128) * cdiv(Cout, 128);
    if (big >= 2 * 148 && Cout > 64) {
        dim3 grid(cdiv(p.M, 128), cdiv(Cout, 128));
        conv_igemm_kernel<128, 128><<<grid, 256, 0, stream>>>(p);
    } else {
        dim3 grid(cdiv(p.M, 64), cdiv(Cout, 64));
        conv_igemm_kernel<64, 64><<<grid, 256, 0, stream>>>(p);
    }
    return launch_status("aoc_conv2d_nhwc_f32");
}

extern "C" int aoc_dwconv3x3_nhwc_f32(const float* x, const float* w, const float* bias, float* y, int N, int H,
                                      int W, int C, cudaStream_t stream) {
    AOC_CHECK_ARG(x && w && y, "null pointer");
    AOC_CHECK_ARG(C % 4 == 0 && N > 0 && H > 0 && W > 0, "C must be a multiple of 4");
    long long total = (long long)N * H * W * (C / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    dwconv3x3_kernel<<<blocks, 256, 0, stream>>>(x, w, bias, y, N, H, W, C);
    return launch_status("aoc_dwconv3x3_nhwc_f32");
}

extern "C" int aoc_maxpool3x3s2_nhwc_f32(const float* x, float* y, int N, int H, int W, int C, cudaStream_t stream) {
    AOC_CHECK_ARG(x && y, "null pointer");
    AOC_CHECK_ARG(C % 4 == 0 && N > 0 && H > 0 && W > 0, "C must be a multiple of 4");
    int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    long long total = (long long)N * Ho * Wo * (C / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    maxpool3x3s2_kernel<<<blocks, 256, 0, stream>>>(x, y, N, H, W, C, Ho, Wo);
    return launch_status("aoc_maxpool3x3s2_nhwc_f32");
}

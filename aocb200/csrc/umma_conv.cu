// Implicit-GEMM convolution on the tcgen05 tensor cores (NHWC activations, fp32-faithful 3xTF32).
//   y[m, co] = sum_k A[m, k] * W[co, k],  m = (n, ho, wo), k = (r, s, ci)
// Replaces the cuDNN nn.Conv2d calls of the ResNet101-DeepLabv3+ backbone and of the calibration decoder
// (resnet.py:23-42, deeplab/aspp.py:62-74, deeplab/decoder.py:32-41, layers/gct.py:68-91, layers/aspp.py:57-70,
// decoding_module.py:162-190) -- same contract as aoc_conv2d_nhwc_f32 (bias / residual / ReLU / per-(n,cin) input gate
// fused), checked against it in tests/test_gpu_ops.py.
//
// CTA = 128 output pixels (TMEM lanes) x TN output channels.  K is consumed in stages of 16 floats:
//   * 8 producer warps (two groups of 128 threads alternating stages) gather the im2col rows with 128-bit loads,
//     apply the optional input gate, split hi/lo and write the tcgen05 core-matrix layout into shared memory
//     (generic proxy -> fence.proxy.async -> mbarrier);
//   * warp 8 streams the pre-packed weight image with TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//   * warp 9 (one thread) issues tcgen05.mma kind::tf32, 3 MMAs per k-step, accumulating in TMEM;
//   * warps 0..7 run the epilogue out of TMEM (tcgen05.ld): bias, residual, ReLU, 128-bit stores.
#include "common.cuh"
#include "umma.cuh"

namespace aoc {
using namespace umma;

constexpr int CV_BM = 128;
constexpr int CV_KST = 2;                       // k-steps per stage (16 floats)
constexpr int CV_NST = 4;                       // pipeline stages
constexpr int CV_WRB = 128;                     // row block of the packed weight image
constexpr uint32_t CV_A_STAGE = CV_KST * 2 * CV_BM * KSTEP * 4;   // 16384

struct ConvTC {
    const float* x; const uint8_t* w; const float* bias; const float* res; const float* in_scale; float* y;
    int N, H, W, Cin, ldx, Ho, Wo, Cout, ldy, ldres, kh, kw, stride, pad, dil, relu;
    int M, K, KS;          // KS = k-steps of the weight image (K padded to 16)
    int vec_out;
};

template <int TN>
__global__ void __launch_bounds__(384, 1) conv_tc_kernel(ConvTC p) {
    constexpr uint32_t B_STAGE = CV_KST * 2 * TN * KSTEP * 4;
    constexpr uint32_t STAGE = CV_A_STAGE + B_STAGE;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CV_NST * STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CV_NST + 1);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (CV_NST + s); };
    const uint32_t TFULL = bar0 + 8u * (2 * CV_NST);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * CV_BM, n0 = blockIdx.y * TN;
    const int nkt = p.KS / CV_KST;

    if (warp == 9 && lane == 0) {
        for (int s = 0; s < CV_NST; ++s) { mbar_init(FULL(s), 129); mbar_init(EMPTY(s), 1); }
        mbar_init(TFULL, 1);
        fence_barrier_init();
    }
    if (warp == 10) {
        tmem_alloc(smem_u32(tmem_slot), TN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ===== A producers: thread <-> output pixel row of the tile =====
        const int grp = warp >> 2;
        const int row = threadIdx.x & 127;
        const int m = m0 + row;
        const bool mok = m < p.M;
        const int HoWo = p.Ho * p.Wo;
        const int mm = mok ? m : 0;
        const int n = mm / HoWo;
        const int rem = mm - n * HoWo;
        const int ho = rem / p.Wo, wo = rem - ho * p.Wo;
        const int hi0 = ho * p.stride - p.pad, wi0 = wo * p.stride - p.pad;
        const float* xn = p.x + (size_t)n * p.H * p.W * p.ldx;
        const float* sc = p.in_scale ? p.in_scale + (size_t)n * p.Cin : nullptr;
        const uint32_t row_off = (uint32_t)((row >> 3) * 256 + (row & 7) * 16);

        auto gather = [&](int kt, float4* v) {
            int k = kt * (CV_KST * KSTEP);
            int rs = k / p.Cin, ci = k - rs * p.Cin;
#pragma unroll
            for (int g = 0; g < CV_KST * 2; ++g) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (mok && k < p.K) {
                    int r = rs / p.kw, s = rs - r * p.kw;
                    int hi = hi0 + r * p.dil, wi = wi0 + s * p.dil;
                    if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) {
                        t = ldg4(xn + ((size_t)hi * p.W + wi) * p.ldx + ci);
                        if (sc) {
                            float4 q = ldg4(sc + ci);
                            t.x *= q.x; t.y *= q.y; t.z *= q.z; t.w *= q.w;
                        }
                    }
                }
                v[g] = t;
                k += 4; ci += 4;
                if (ci >= p.Cin) { ci -= p.Cin; ++rs; }
            }
        };
        float4 v[CV_KST * 2], nv[CV_KST * 2];
        if (grp < nkt) gather(grp, v);
        for (int kt = grp; kt < nkt; kt += 2) {
            const bool more = kt + 2 < nkt;
            if (more) gather(kt + 2, nv);
            const int stage = kt % CV_NST;
            const uint32_t phase = (uint32_t)(kt / CV_NST) & 1u;
            mbar_wait(EMPTY(stage), phase ^ 1u);
            uint8_t* sa = smem + (size_t)stage * STAGE;
#pragma unroll
            for (int g = 0; g < CV_KST * 2; ++g) {
                float h[4], l[4];
                split_tf32(v[g].x, h[0], l[0]); split_tf32(v[g].y, h[1], l[1]);
                split_tf32(v[g].z, h[2], l[2]); split_tf32(v[g].w, h[3], l[3]);
                // k-step (g>>1): [hi block 4096 B][lo block 4096 B]; granule (g&1) -> +128 B
                uint8_t* d = sa + (g >> 1) * (2 * CV_BM * KSTEP * 4) + (g & 1) * 128 + row_off;
                *reinterpret_cast<float4*>(d) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(d + CV_BM * KSTEP * 4) = make_float4(l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async();
            mbar_arrive(FULL(stage));
            if (more) {
#pragma unroll
                for (int g = 0; g < CV_KST * 2; ++g) v[g] = nv[g];
            }
        }
        // ===== epilogue: warps 0..3 take columns [0, TN/2), warps 4..7 take [TN/2, TN) =====
        mbar_wait(TFULL, 0);
        tc_fence_after();
        const int wq = warp & 3;
        const int erow = wq * 32 + lane;
        const int em = m0 + erow;
        const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16);
        const int cbeg = grp * (TN / 2), cend = cbeg + TN / 2;
#pragma unroll 1
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
            float a[32];
            tmem_ld32(t0 + c0, a);
            tmem_ld_wait();
            const int co0 = n0 + c0;
            if (em < p.M && co0 < p.Cout) {
                float* dst = p.y + (size_t)em * p.ldy + co0;
                const float* rsd = p.res ? p.res + (size_t)em * p.ldres + co0 : nullptr;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const int co = co0 + j4 * 4;
                    if (co >= p.Cout) break;
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        o[e] = a[j4 * 4 + e];
                        if (co + e < p.Cout) {
                            if (p.bias) o[e] += __ldg(p.bias + co + e);
                            if (rsd) o[e] += __ldg(rsd + j4 * 4 + e);
                            if (p.relu) o[e] = fmaxf(o[e], 0.f);
                        }
                    }
                    if (p.vec_out && co + 3 < p.Cout) {
                        *reinterpret_cast<float4*>(dst + j4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (co + e < p.Cout) dst[j4 * 4 + e] = o[e];
                    }
                }
            }
        }
    } else if (warp == 8) {
        if (lane == 0) {
            // ===== weight TMA producer: image chunk(rb, ks) = [hi 128x32 B][lo 128x32 B] =====
            constexpr int NRB = TN > CV_WRB ? TN / CV_WRB : 1;
            constexpr uint32_t BLK = (TN < CV_WRB ? TN : CV_WRB) * KSTEP * 4;    // bytes of one hi (or lo) block copy
            const int rb0 = n0 / CV_WRB;
            const uint32_t sub = (uint32_t)(n0 % CV_WRB) * KSTEP * 4;           // TN=64: second half of a row block
            for (int kt = 0; kt < nkt; ++kt) {
                const int stage = kt % CV_NST;
                const uint32_t phase = (uint32_t)(kt / CV_NST) & 1u;
                mbar_wait(EMPTY(stage), phase ^ 1u);
                mbar_arrive_expect_tx(FULL(stage), B_STAGE);
                const uint32_t sb = smem_u32(smem + (size_t)stage * STAGE + CV_A_STAGE);
#pragma unroll
                for (int j = 0; j < CV_KST; ++j) {
                    const int ks = kt * CV_KST + j;
#pragma unroll
                    for (int b = 0; b < NRB; ++b) {
                        const uint8_t* src = p.w + ((size_t)(rb0 + b) * p.KS + ks) * (2 * CV_WRB * KSTEP * 4) + sub;
                        const uint32_t dj = sb + j * (2 * TN * KSTEP * 4) + b * BLK;
                        bulk_g2s(dj, src, BLK, FULL(stage));                                    // hi
                        bulk_g2s(dj + TN * KSTEP * 4, src + CV_WRB * KSTEP * 4, BLK, FULL(stage));  // lo
                    }
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = idesc_tf32(CV_BM, TN);
            for (int kt = 0; kt < nkt; ++kt) {
                const int stage = kt % CV_NST;
                const uint32_t phase = (uint32_t)(kt / CV_NST) & 1u;
                mbar_wait(FULL(stage), phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE);
                const uint32_t sb = sa + CV_A_STAGE;
#pragma unroll
                for (int j = 0; j < CV_KST; ++j) {
                    const uint32_t a_hi = sa + j * (2 * CV_BM * KSTEP * 4), a_lo = a_hi + CV_BM * KSTEP * 4;
                    const uint32_t b_hi = sb + j * (2 * TN * KSTEP * 4), b_lo = b_hi + TN * KSTEP * 4;
                    const uint64_t dah = smem_desc(a_hi, LBO_BYTES, SBO_BYTES), dal = smem_desc(a_lo, LBO_BYTES, SBO_BYTES);
                    const uint64_t dbh = smem_desc(b_hi, LBO_BYTES, SBO_BYTES), dbl = smem_desc(b_lo, LBO_BYTES, SBO_BYTES);
                    mma_tf32(tmem_base, dal, dbh, idesc, (kt > 0 || j > 0) ? 1u : 0u);
                    mma_tf32(tmem_base, dah, dbl, idesc, 1u);
                    mma_tf32(tmem_base, dah, dbh, idesc, 1u);
                }
                mma_commit(EMPTY(stage));
            }
            mma_commit(TFULL);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 10) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TN);
    }
}

// w [Cout][K] fp32 -> tc image with 128-row blocks, K padded to a multiple of 16
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int K, int KS, int rows_padded,
                                    uint8_t* __restrict__ out) {
    long long total = (long long)rows_padded * KS * 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int r = (int)(i % rows_padded);
        int g = (int)(i / rows_padded);
        int ks = g >> 1, half = g & 1;
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int k = g * 4 + e;
            float v = (r < Cout && k < K) ? __ldg(w + (size_t)r * K + k) : 0.f;
            split_tf32(v, hi[e], lo[e]);
        }
        int rb = r / CV_WRB, rr = r - rb * CV_WRB;
        size_t base = ((size_t)rb * KS + ks) * chunk_bytes(CV_WRB) + elem_offset(rr, half * 4);
        *reinterpret_cast<float4*>(out + base) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(out + base + block_bytes(CV_WRB)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

template <int TN>
static int launch_conv_tc(const ConvTC& p, cudaStream_t stream) {
    constexpr uint32_t STAGE = CV_A_STAGE + CV_KST * 2 * TN * KSTEP * 4;
    constexpr uint32_t SMEM = CV_NST * STAGE + 256;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(conv_tc_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        attr = true;
    }
    dim3 grid(cdiv(p.M, CV_BM), cdiv(p.Cout, TN));
    conv_tc_kernel<TN><<<grid, 384, SMEM, stream>>>(p);
    return launch_status("aoc_conv2d_nhwc_tc");
}

}  // namespace aoc

using namespace aoc;

static inline int conv_ks(int K) { return (K + 15) / 16 * 2; }

extern "C" size_t aoc_conv_packed_weight_bytes(int Cout, int K) {
    size_t rb = (size_t)cdiv(Cout, CV_WRB) + 1;   // +1 block: TN=256 tiles may read one block past the last
    return rb * conv_ks(K) * chunk_bytes(CV_WRB);
}

extern "C" int aoc_conv_pack_weights_tf32x3(const float* w, int Cout, int K, void* w_packed, cudaStream_t stream) {
    AOC_CHECK_ARG(w && w_packed && Cout > 0 && K > 0, "bad args");
    int KS = conv_ks(K);
    int rows_padded = (cdiv(Cout, CV_WRB) + 1) * CV_WRB;
    long long total = (long long)rows_padded * KS * 2;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    pack_weights_kernel<<<blocks, 256, 0, stream>>>(w, Cout, K, KS, rows_padded, (uint8_t*)w_packed);
    return launch_status("aoc_conv_pack_weights_tf32x3");
}

extern "C" int aoc_conv2d_nhwc_tc(const float* x, const void* w_packed, const float* bias, const float* residual,
                                  const float* in_scale, float* y, int N, int H, int W, int Cin, int ldx, int Cout,
                                  int ldy, int ldres, int kh, int kw, int stride, int pad, int dil, int relu,
                                  cudaStream_t stream) {
    AOC_CHECK_ARG(x && w_packed && y, "null pointer");
    AOC_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0 && dil > 0, "bad dims");
    AOC_CHECK_ARG(Cin % 4 == 0 && ldx % 4 == 0 && (((uintptr_t)x) & 15) == 0, "Cin/ldx must be multiples of 4, x 16B aligned");
    AOC_CHECK_ARG(!in_scale || (((uintptr_t)in_scale) & 15) == 0, "in_scale must be 16B aligned");
    ConvTC p;
    p.x = x; p.w = (const uint8_t*)w_packed; p.bias = bias; p.res = residual; p.in_scale = in_scale; p.y = y;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.ldx = ldx; p.Cout = Cout; p.ldy = ldy; p.ldres = ldres;
    p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.dil = dil; p.relu = relu;
    p.Ho = (H + 2 * pad - dil * (kh - 1) - 1) / stride + 1;
    p.Wo = (W + 2 * pad - dil * (kw - 1) - 1) / stride + 1;
    AOC_CHECK_ARG(p.Ho > 0 && p.Wo > 0, "empty output");
    p.M = N * p.Ho * p.Wo;
    p.K = kh * kw * Cin;
    p.KS = conv_ks(p.K);
    p.vec_out = (ldy % 4 == 0) && (((uintptr_t)y & 15) == 0);
    // N tile: as wide as possible while keeping >= ~1 wave of CTAs
    int mt = cdiv(p.M, CV_BM);
    int tn = 256;
    while (tn > 64 && (Cout <= tn / 2 || (long long)mt * cdiv(Cout, tn) < 148)) tn >>= 1;
    if (tn == 256) return launch_conv_tc<256>(p, stream);
    if (tn == 128) return launch_conv_tc<128>(p, stream);
    return launch_conv_tc<64>(p, stream);
}

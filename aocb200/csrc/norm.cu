// Per-(sample, channel) statistics over NHWC maps and the per-(sample, channel) affine kernels built on them:
// GroupNorm, GCT gate, IA gate / conditioning-block scale (x * a[n,c]), global average pool, masked GAP.
// All are HBM-bound reduction / streaming kernels: 128-bit coalesced loads, fixed reduction order
// (deterministic), no tensor cores.
//
// Replaces: nn.GroupNorm call sites (networks/layers/gct.py:68-91, aocnet.py:19-25, decoding_module.py),
// GCT (networks/layers/gct.py:17-36), IA_gate (networks/layers/attention.py:12-17), the FiLM scale of
// conditioning_block (networks/aoc/conditioning_layer.py:63-86) and the masked global average pool of
// conditioning_layer (networks/aoc/conditioning_layer.py:24-48).
#include "common.cuh"

namespace aoc {

// part layout: [N][S][2][C] doubles (sum, sum of squares).  MASKED: weight = phi[n,p] > thr[n] (strict).
template <bool MASKED>
__global__ void __launch_bounds__(256, 6) channel_stats_partial(const float* __restrict__ x, int HW, int C, int ldx,
                                                              int PB, int CW, const float* __restrict__ phi,
                                                              const float* __restrict__ thr,
                                                              double* __restrict__ part) {
    extern __shared__ float sm[];  // [PL][2][Cc]
    const int n = blockIdx.y, s = blockIdx.x, S = gridDim.x;
    const int c0 = blockIdx.z * CW;            // channel chunk (C > 1024 is processed in chunks of CW channels)
    const int Cc = min(CW, C - c0);
    const int Q = Cc >> 2;
    const int PL = 256 / Q;
    const int tid = threadIdx.x;
    const int q = tid % Q, pl = tid / Q;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f), sq = sum;
    if (pl < PL) {
        int pend = min((s + 1) * PB, HW);
        float th = MASKED ? __ldg(thr + n) : 0.f;
        // four pixels per iteration: the loads are issued together (one 16-byte load in flight per thread left the
        // kernel at ~40 % of the HBM rate); the sums keep the sequential pixel order
        for (int p0 = s * PB + pl; p0 < pend; p0 += 4 * PL) {
            float4 v[4];
            bool on[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int p = p0 + j * PL;
                on[j] = p < pend;
                if (MASKED) on[j] = on[j] && (__ldg(phi + (size_t)n * HW + (on[j] ? p : 0)) > th);
                v[j] = on[j] ? ldg4(x + ((size_t)n * HW + p) * ldx + c0 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!on[j]) continue;
                sum.x += v[j].x; sum.y += v[j].y; sum.z += v[j].z; sum.w += v[j].w;
                sq.x = fmaf(v[j].x, v[j].x, sq.x); sq.y = fmaf(v[j].y, v[j].y, sq.y);
                sq.z = fmaf(v[j].z, v[j].z, sq.z); sq.w = fmaf(v[j].w, v[j].w, sq.w);
            }
        }
        float* d = sm + (size_t)pl * 2 * Cc;
        *reinterpret_cast<float4*>(d + q * 4) = sum;
        *reinterpret_cast<float4*>(d + Cc + q * 4) = sq;
    }
    __syncthreads();
    for (int c = tid; c < 2 * Cc; c += 256) {
        double a = 0.0;
        for (int l = 0; l < PL; ++l) a += (double)sm[(size_t)l * 2 * Cc + c];
        int which = c / Cc, cc = c - which * Cc;
        part[((size_t)(n * S + s) * 2 + which) * C + c0 + cc] = a;
    }
}

// y = x*a + b (+ res*ra) (ReLU) AND the per-(sample, channel) sum / sum of squares of y in the same pass (the GCT gate /
// global average pool of the next block then needs no pass of its own).  Same slab decomposition and fixed reduction
// order as channel_stats_partial; part layout [N][S][2][C] doubles.
__global__ void __launch_bounds__(256, 4) affine_stats_partial(const float* __restrict__ x, const float* __restrict__ a,
                                                             const float* __restrict__ b, const float* __restrict__ res,
                                                             const float* __restrict__ res_scale, float* __restrict__ y,
                                                             int HW, int C, int ldx, int ldy, int ldres, int relu, int PB,
                                                             double* __restrict__ part) {
    extern __shared__ float sm[];  // [PL][2][C]
    const int n = blockIdx.y, s = blockIdx.x, S = gridDim.x;
    const int Q = C >> 2;
    const int PL = 256 / Q;
    const int tid = threadIdx.x;
    const int q = tid % Q, pl = tid / Q;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f), sq = sum;
    if (pl < PL) {
        const int c = q * 4;
        const float4 av = ldg4(a + (size_t)n * C + c);
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), rs = make_float4(1.f, 1.f, 1.f, 1.f);
        if (b) bv = ldg4(b + (size_t)n * C + c);
        if (res && res_scale) rs = ldg4(res_scale + (size_t)n * C + c);
        const int pend = min((s + 1) * PB, HW);
        for (int p0 = s * PB + pl; p0 < pend; p0 += 2 * PL) {
            float4 vv[2], rr[2];
            bool on[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {                            // both pixels' loads in flight together
                const int p = p0 + j * PL;
                on[j] = p < pend;
                const size_t pix = (size_t)n * HW + (on[j] ? p : p0);
                vv[j] = ldg4(x + pix * ldx + c);
                rr[j] = res ? ldg4(res + pix * ldres + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (!on[j]) continue;
                const size_t pix = (size_t)n * HW + p0 + j * PL;
                const float4 v = vv[j];
                float4 o;
                if (b) {
                    o.x = fmaf(v.x, av.x, bv.x); o.y = fmaf(v.y, av.y, bv.y);
                    o.z = fmaf(v.z, av.z, bv.z); o.w = fmaf(v.w, av.w, bv.w);
                } else {
                    o.x = v.x * av.x; o.y = v.y * av.y; o.z = v.z * av.z; o.w = v.w * av.w;
                }
                if (res) {
                    float4 r = rr[j];
                    if (res_scale) { r.x *= rs.x; r.y *= rs.y; r.z *= rs.z; r.w *= rs.w; }
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                *reinterpret_cast<float4*>(y + pix * ldy + c) = o;
                sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
                sq.x = fmaf(o.x, o.x, sq.x); sq.y = fmaf(o.y, o.y, sq.y);
                sq.z = fmaf(o.z, o.z, sq.z); sq.w = fmaf(o.w, o.w, sq.w);
            }
        }
        float* d = sm + (size_t)pl * 2 * C;
        *reinterpret_cast<float4*>(d + q * 4) = sum;
        *reinterpret_cast<float4*>(d + C + q * 4) = sq;
    }
    __syncthreads();
    for (int c = tid; c < 2 * C; c += 256) {
        double acc = 0.0;
        for (int l = 0; l < PL; ++l) acc += (double)sm[(size_t)l * 2 * C + c];
        part[(size_t)(n * S + s) * 2 * C + c] = acc;
    }
}

// stats: [N][2][C] doubles.  Block = 32 values x 8 slab lanes: lane l adds the slabs l, l + 8, ... (two loads in flight),
// the eight lane sums are combined in lane order -- a fixed tree, so the result does not depend on the schedule; the
// serial chain over the slabs (101 for a 121 x 213 map) that made this tiny kernel 5-12 us long is 13 steps instead of 26 x 4.
__global__ void __launch_bounds__(256) channel_stats_final(const double* __restrict__ part, int S, int C2,
                                                           double* __restrict__ stats) {
    AOC_PDL_TRIGGER();
    __shared__ double sm[8][33];
    const int n = blockIdx.y;
    const int c = blockIdx.x * 32 + threadIdx.x, l = threadIdx.y;
    double a = 0.0;
    if (c < C2) {
        const double* p = part + (size_t)n * S * C2 + c;
        for (int s = l; s < S; s += 16) {
            const double v0 = p[(size_t)s * C2];
            const double v1 = s + 8 < S ? p[(size_t)(s + 8) * C2] : 0.0;
            a += v0;
            a += v1;
        }
    }
    sm[l][threadIdx.x] = a;
    __syncthreads();
    if (l == 0 && c < C2) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
        stats[(size_t)n * C2 + c] = t;
    }
}

// GroupNorm coefficients: y = x*a + b with a = rstd*gamma, b = beta - mean*a   (biased variance, eps inside sqrt)
__global__ void gn_coeffs_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int C, int groups, int HW, float eps,
                                 float* __restrict__ a, float* __restrict__ b) {
    AOC_PDL_TRIGGER();
    int n = blockIdx.x;
    int cpg = C / groups;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            s += stats[(size_t)n * 2 * C + c];
            q += stats[(size_t)n * 2 * C + C + c];
        }
        double cnt = (double)cpg * HW;
        double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        float rstd = (float)(1.0 / sqrt(var + (double)eps));
        float fmean = (float)mean;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            float ga = gamma[c] * rstd;
            a[(size_t)n * C + c] = ga;
            b[(size_t)n * C + c] = beta[c] - fmean * ga;
        }
    }
}

// GCT gate (gct.py:17-36, mode l2): embedding = sqrt(sum x^2 + eps)*alpha; norm = gamma/sqrt(mean_c(emb^2)+eps);
// gate = 1 + tanh(emb*norm + beta).  `pre` (optional, [N,C]) is a per-channel scale already applied to x
// analytically (x_eff = pre*x): sumsq_eff = pre^2*sumsq, and the returned gate is multiplied by pre.
__global__ void gct_coeffs_kernel(const double* __restrict__ stats, const float* __restrict__ alpha,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const float* __restrict__ pre, int C, float eps, float* __restrict__ a) {
    AOC_PDL_TRIGGER();
    __shared__ double red[32];
    __shared__ double s_mean;
    int n = blockIdx.x;
    double loc = 0.0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double sq = stats[(size_t)n * 2 * C + C + c];
        if (pre) { double pz = (double)pre[(size_t)n * C + c]; sq *= pz * pz; }
        float e = sqrtf((float)sq + eps) * alpha[c];
        loc += (double)e * (double)e;
    }
    loc = warp_sum_d(loc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) t += red[i];
        s_mean = t / C;
    }
    __syncthreads();
    float inv = 1.0f / sqrtf((float)s_mean + eps);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double sq = stats[(size_t)n * 2 * C + C + c];
        float pz = 1.f;
        if (pre) { pz = pre[(size_t)n * C + c]; sq *= (double)pz * (double)pz; }
        float e = sqrtf((float)sq + eps) * alpha[c];
        float nrm = gamma[c] * inv;
        a[(size_t)n * C + c] = pz * (1.0f + tanhf(e * nrm + beta[c]));
    }
}

__global__ void gap_from_stats_kernel(const double* __restrict__ stats, int C, float inv_hw, float* __restrict__ out) {
    AOC_PDL_TRIGGER();
    int n = blockIdx.y;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) out[(size_t)n * C + c] = (float)(stats[(size_t)n * 2 * C + c] * (double)inv_hw);
}

// y[n,p,c] = x[n,p,c]*a[n,c] + b[n,c] (+ res[n,p,c]*ra[n,c]) (ReLU)
__global__ void affine_nc_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                                 const float* __restrict__ res, const float* __restrict__ res_scale,
                                 float* __restrict__ y, int HW, int C, int ldx, int ldy, int ldres, int relu,
                                 long long total4) {
    AOC_PDL_TRIGGER();
    int C4 = C >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C4) * 4;
        long long pix = i / C4;
        int n = (int)(pix / HW);
        float4 v = ldg4(x + (size_t)pix * ldx + c);
        float4 av = ldg4(a + (size_t)n * C + c);
        float4 o;
        if (b) {
            float4 bv = ldg4(b + (size_t)n * C + c);
            o.x = fmaf(v.x, av.x, bv.x); o.y = fmaf(v.y, av.y, bv.y);
            o.z = fmaf(v.z, av.z, bv.z); o.w = fmaf(v.w, av.w, bv.w);
        } else {
            o.x = v.x * av.x; o.y = v.y * av.y; o.z = v.z * av.z; o.w = v.w * av.w;
        }
        if (res) {
            float4 r = ldg4(res + (size_t)pix * ldres + c);
            if (res_scale) {
                float4 rs = ldg4(res_scale + (size_t)n * C + c);
                r.x *= rs.x; r.y *= rs.y; r.z *= rs.z; r.w *= rs.w;
            }
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(y + (size_t)pix * ldy + c) = o;
    }
}

// tile partials [N * tiles_per_image][2][C] floats (written by the convolution epilogue) -> stats [N][2][C] doubles
__global__ void __launch_bounds__(512) tile_stats_reduce_kernel(const float* __restrict__ part, int tiles_per_image,
                                                                int C2, double* __restrict__ stats) {
    // block = 64 channels x 8 tile lanes: lane ty sums the tiles t = ty, ty + 8, ... (independent loads in flight), the
    // eight partial sums are then added in lane order -- a fixed order, so the result is deterministic
    __shared__ double sm[8][64];
    const int n = blockIdx.y;
    const int cx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * 64 + cx;
    double a = 0.0;
    if (c < C2) {
        const float* p = part + (size_t)n * tiles_per_image * C2 + c;
        for (int t = ty; t < tiles_per_image; t += 8) a += (double)__ldg(p + (size_t)t * C2);
    }
    sm[ty][cx] = a;
    __syncthreads();
    if (ty == 0 && c < C2) {
        double r = sm[0][cx];
#pragma unroll
        for (int l = 1; l < 8; ++l) r += sm[l][cx];
        stats[(size_t)n * C2 + c] = r;
    }
}

// GroupNorm coefficients straight from the convolution's per-tile partial statistics: block = (group, sample); thread
// (item = stat x channel of the group, tile lane) sums its tiles in double, the lanes are combined in lane order.
// Replaces tile_stats_reduce + gn_coeffs (two dependent 5 us launches per normalisation, ~50 per frame) by one.
__global__ void __launch_bounds__(256) gn_coeffs_tiles_kernel(const float* __restrict__ part, int tiles_per_image,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int C, int groups, int HW,
                                                              float eps, float* __restrict__ a, float* __restrict__ b) {
    AOC_PDL_TRIGGER();
    AOC_PDL_WAIT();                                   // launched as a programmatic dependent of the convolution (launch_pdl)
    __shared__ double sm[256];
    __shared__ double tot[64];
    __shared__ float s_rstd, s_mean;
    const int g = blockIdx.x, n = blockIdx.y;
    const int cpg = C / groups;
    const int items = 2 * cpg;                        // <= 64 (checked by the launcher)
    const int lanes = 256 / items;
    const int it = threadIdx.x % items, tl = threadIdx.x / items;
    double acc = 0.0;
    if (tl < lanes) {
        const int stat = it / cpg, ch = it - stat * cpg;
        const float* p = part + ((size_t)n * tiles_per_image * 2 + stat) * C + g * cpg + ch;
        // sixteen rows per iteration: independent loads in flight, added in row order (the convolution writes one row per
        // 32-pixel warp quadrant: 4 832 rows for a 121 x 213 map -- with four loads in flight the chain was 38 L2 round trips)
        constexpr int U = 16;
        for (int t = tl; t < tiles_per_image; t += U * lanes) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int tt = t + u * lanes;
                v[u] = tt < tiles_per_image ? __ldg(p + (size_t)tt * 2 * C) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += (double)v[u];
        }
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < items) {
        double r = 0.0;
        for (int l = 0; l < lanes; ++l) r += sm[l * items + threadIdx.x];
        tot[threadIdx.x] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0, q = 0.0;
        for (int c = 0; c < cpg; ++c) { s += tot[c]; q += tot[cpg + c]; }
        const double cnt = (double)cpg * HW;
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_rstd = (float)(1.0 / sqrt(var + (double)eps));
        s_mean = (float)mean;
    }
    __syncthreads();
    if (threadIdx.x < cpg) {
        const int c = g * cpg + threadIdx.x;
        const float ga = gamma[c] * s_rstd;
        a[(size_t)n * C + c] = ga;
        b[(size_t)n * C + c] = beta[c] - s_mean * ga;
    }
}

// pixels per slab (= per block) of the statistics kernels.  The blocks of a launch are equal pieces of work, so a grid a few
// blocks LARGER than what the chip holds at once runs for two block lifetimes instead of one: 6 x 101 = 606 slabs of 256
// pixels (121 x 213 map, six objects) on 148 x 4 = 592 resident blocks was the common case, and 6 x 102 = 612 slabs of 64
// pixels at half resolution likewise.  The slab is therefore sized so that the grid is a whole number of waves: `bps`
// resident blocks per SM (held by __launch_bounds__), waves = what the nominal 256-pixel slab would need, rounded DOWN
// (slabs of 256 .. 511 pixels); small launches (N = 1 maps of the backbone) get one wave of slabs of at least 32 pixels.
constexpr int STATS_BPS = 6, AFFINE_STATS_BPS = 4;      // = minBlocksPerMultiprocessor of the two partial kernels
static int stats_slab(int N, int HW, int C, int bps) {
    const long long z = cdiv(C, 1024), R = 148LL * bps;
    const long long nominal = (long long)cdiv(HW, 256) * N * z;
    const long long waves = nominal / R > 0 ? nominal / R : 1;
    long long S = waves * R / ((long long)N * z);
    if (S < 1) S = 1;
    if (S > 1024) S = 1024;
    int pb = cdiv(HW, (int)S);
    return pb < 32 ? 32 : pb;
}

}  // namespace aoc

using namespace aoc;

extern "C" size_t aoc_channel_stats_workspace_bytes(int N, int HW, int C) {
    const int Sa = cdiv(HW, stats_slab(N, HW, C, STATS_BPS)), Sb = cdiv(HW, stats_slab(N, HW, C, AFFINE_STATS_BPS));
    return (size_t)N * (size_t)(Sa > Sb ? Sa : Sb) * 2 * C * sizeof(double);       // serves both entry points
}

// stats out: [N][2][C] doubles (sum, sumsq).  phi/thr optional (masked sum: only pixels with phi > thr[n]).
extern "C" int aoc_channel_stats_f32(const float* x, int N, int HW, int C, int ldx, const float* phi,
                                     const float* thr, double* stats, void* workspace, size_t ws_bytes,
                                     cudaStream_t stream) {
    AOC_CHECK_ARG(x && stats && workspace, "null pointer");
    AOC_CHECK_ARG(C % 4 == 0 && C >= 4 && ldx % 4 == 0, "C and ldx must be multiples of 4");
    AOC_CHECK_ARG((((uintptr_t)x) & 15) == 0, "x must be 16-byte aligned");
    AOC_CHECK_ARG(ws_bytes >= aoc_channel_stats_workspace_bytes(N, HW, C), "workspace too small");
    AOC_CHECK_ARG((phi == nullptr) == (thr == nullptr), "phi and thr go together");
    int PB = stats_slab(N, HW, C, STATS_BPS);
    int S = cdiv(HW, PB);
    int CW = C < 1024 ? C : 1024;
    int PL = 256 / (CW / 4);
    size_t smem = (size_t)PL * 2 * CW * sizeof(float);
    dim3 grid(S, N, cdiv(C, CW));
    double* part = (double*)workspace;
    if (phi)
        channel_stats_partial<true><<<grid, 256, smem, stream>>>(x, HW, C, ldx, PB, CW, phi, thr, part);
    else
        channel_stats_partial<false><<<grid, 256, smem, stream>>>(x, HW, C, ldx, PB, CW, nullptr, nullptr, part);
    dim3 g2(cdiv(2 * C, 32), N);
    channel_stats_final<<<g2, dim3(32, 8), 0, stream>>>(part, S, 2 * C, stats);
    return launch_status("aoc_channel_stats_f32");
}

// y = x*a[n,c] (+ b[n,c]) (+ residual*res_scale[n,c]) (ReLU) with the statistics of y as a by-product.
// workspace: aoc_channel_stats_workspace_bytes(N, HW, C); stats: [N][2][C] doubles; C <= 1024.
extern "C" int aoc_affine_stats_nc_f32(const float* x, const float* a, const float* b, const float* residual,
                                       const float* res_scale, float* y, int N, int HW, int C, int ldx, int ldy,
                                       int ldres, int relu, double* stats, void* workspace, size_t ws_bytes,
                                       cudaStream_t stream) {
    AOC_CHECK_ARG(x && a && y && stats && workspace, "null pointer");
    AOC_CHECK_ARG(C % 4 == 0 && C >= 4 && C <= 1024 && ldx % 4 == 0 && ldy % 4 == 0 && (!residual || ldres % 4 == 0),
                  "C (<= 1024) and the row strides must be multiples of 4");
    AOC_CHECK_ARG(((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)a)) & 15) == 0, "pointers must be 16-byte aligned");
    AOC_CHECK_ARG(ws_bytes >= aoc_channel_stats_workspace_bytes(N, HW, C), "workspace too small");
    int PB = stats_slab(N, HW, C, AFFINE_STATS_BPS);
    int S = cdiv(HW, PB);
    int PL = 256 / (C / 4);
    size_t smem = (size_t)PL * 2 * C * sizeof(float);
    dim3 grid(S, N);
    double* part = (double*)workspace;
    affine_stats_partial<<<grid, 256, smem, stream>>>(x, a, b, residual, res_scale, y, HW, C, ldx, ldy, ldres, relu, PB,
                                                      part);
    dim3 g2(cdiv(2 * C, 32), N);
    channel_stats_final<<<g2, dim3(32, 8), 0, stream>>>(part, S, 2 * C, stats);
    return launch_status("aoc_affine_stats_nc_f32");
}

extern "C" int aoc_tile_stats_reduce_f32(const float* tile_stats, int N, int tiles_per_image, int C, double* stats,
                                         cudaStream_t stream) {
    AOC_CHECK_ARG(tile_stats && stats && N > 0 && tiles_per_image > 0 && C > 0, "bad args");
    dim3 g(cdiv(2 * C, 64), N);
    tile_stats_reduce_kernel<<<g, dim3(64, 8), 0, stream>>>(tile_stats, tiles_per_image, 2 * C, stats);
    return launch_status("aoc_tile_stats_reduce_f32");
}

extern "C" int aoc_gn_coeffs_tiles_f32(const float* tile_stats, int tiles_per_image, const float* gamma,
                                       const float* beta, int N, int C, int groups, int HW, float eps, float* a, float* b,
                                       cudaStream_t stream) {
    AOC_CHECK_ARG(tile_stats && gamma && beta && a && b && tiles_per_image > 0, "bad args");
    AOC_CHECK_ARG(groups > 0 && C % groups == 0 && 2 * (C / groups) <= 64, "C / groups must be <= 32");
    launch_pdl(gn_coeffs_tiles_kernel, dim3(groups, N), dim3(256), 0, stream, tile_stats, tiles_per_image, gamma, beta, C,
               groups, HW, eps, a, b);
    return launch_status("aoc_gn_coeffs_tiles_f32");
}

extern "C" int aoc_gn_coeffs_f32(const double* stats, const float* gamma, const float* beta, int N, int C, int groups,
                                 int HW, float eps, float* a, float* b, cudaStream_t stream) {
    AOC_CHECK_ARG(stats && gamma && beta && a && b, "null pointer");
    AOC_CHECK_ARG(groups > 0 && C % groups == 0, "C must be divisible by groups");
    gn_coeffs_kernel<<<N, 64, 0, stream>>>(stats, gamma, beta, C, groups, HW, eps, a, b);
    return launch_status("aoc_gn_coeffs_f32");
}

extern "C" int aoc_gct_coeffs_f32(const double* stats, const float* alpha, const float* gamma, const float* beta,
                                  const float* pre_scale, int N, int C, float eps, float* a, cudaStream_t stream) {
    AOC_CHECK_ARG(stats && alpha && gamma && beta && a, "null pointer");
    gct_coeffs_kernel<<<N, 256, 0, stream>>>(stats, alpha, gamma, beta, pre_scale, C, eps, a);
    return launch_status("aoc_gct_coeffs_f32");
}

extern "C" int aoc_gap_from_stats_f32(const double* stats, int N, int C, int HW, float* out, cudaStream_t stream) {
    AOC_CHECK_ARG(stats && out, "null pointer");
    dim3 g(cdiv(C, 256), N);
    gap_from_stats_kernel<<<g, 256, 0, stream>>>(stats, C, 1.0f / (float)HW, out);
    return launch_status("aoc_gap_from_stats_f32");
}

extern "C" int aoc_affine_nc_f32(const float* x, const float* a, const float* b, const float* residual,
                                 const float* res_scale, float* y, int N, int HW, int C, int ldx, int ldy, int ldres,
                                 int relu, cudaStream_t stream) {
    AOC_CHECK_ARG(x && a && y, "null pointer");
    AOC_CHECK_ARG(C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && (!residual || ldres % 4 == 0), "C/ld must be multiples of 4");
    AOC_CHECK_ARG(((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)a)) & 15) == 0, "pointers must be 16-byte aligned");
    long long total4 = (long long)N * HW * (C / 4);
    int blocks = (int)((total4 + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    affine_nc_kernel<<<blocks, 256, 0, stream>>>(x, a, b, residual, res_scale, y, HW, C, ldx, ldy, ldres, relu, total4);
    return launch_status("aoc_affine_nc_f32");
}

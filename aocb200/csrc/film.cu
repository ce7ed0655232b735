// FiLM-style conditioning-layer kernels (the "scale/shift modulation" of BASELINE.json's north_star).
// conditioning_layer (networks/aoc/conditioning_layer.py:24-48): phi = 1x1 conv C->1; threshold = the
// beta-th largest phi value per object slot (torch.topk(...)[..., -1]); mask = phi > threshold (STRICT);
// masked global average over ALL h*w positions; Linear(C, C).  The masked GAP itself is
// aoc_channel_stats_f32 with (phi, thr) in norm.cu.  IA_gate / conditioning_block gates are small GEMVs
// (aoc_linear_f32 with act=1: 1 + tanh) followed by aoc_affine_nc_f32.
// Warp-shuffle reductions, 128-bit coalesced loads, no tensor cores.
#include "common.cuh"

namespace aoc {

// phi[n,p] = sum_c x[n,p,c]*w[c] + b0 : one warp per pixel.
__global__ void __launch_bounds__(256) cond_phi_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, float* __restrict__ phi,
                                                        long long NP, int C, int ldx) {
    int lane = threadIdx.x & 31;
    long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    int C4 = C >> 2;
    float b0 = b ? __ldg(b) : 0.f;
    for (long long p = warp; p < NP; p += nwarps) {
        const float* row = x + (size_t)p * ldx;
        float acc = 0.f;
        for (int i = lane; i < C4; i += 32) {
            float4 v = ldg4(row + i * 4);
            float4 ww = ldg4(w + i * 4);
            acc = fmaf(v.x, ww.x, acc); acc = fmaf(v.y, ww.y, acc);
            acc = fmaf(v.z, ww.z, acc); acc = fmaf(v.w, ww.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) phi[p] = acc + b0;
    }
}

__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// k-th largest (k is 1-based) of vals[n, 0..L) per n: 4-pass MSB radix select, one block per n (exact: the threshold is
// compared with `>` against every phi value, conditioning_layer.py:34-37).  The row is read from global memory ONCE, as
// order-preserving unsigned keys, into shared memory (100 KB for a 121 x 213 map); every pass then histograms the keys
// that still match the selected prefix into per-warp histograms -- lanes of a warp that hit the same bin are combined
// with match.any first (phi values share their exponent byte: the first pass would otherwise be a 32-way conflict on one
// counter) -- and one warp finds the bin that holds the k-th key with a shuffle scan.  Round 1's version re-read the row
// from L2 in every pass and funnelled all 1024 threads into one histogram: 32-38 us per call, now bounded by ~25
// shared-memory iterations per pass.
constexpr int KTH_THREADS = 1024;
constexpr int KTH_WARPS = KTH_THREADS / 32;
__global__ void __launch_bounds__(KTH_THREADS) kth_largest_kernel(const float* __restrict__ vals, int L, int k,
                                                                   float* __restrict__ out, int keys_in_smem) {
    extern __shared__ unsigned kth_smem[];
    unsigned* whist = kth_smem;                                  // [KTH_WARPS][256]
    unsigned* hist = whist + KTH_WARPS * 256;                    // [256]
    unsigned* keys = hist + 256;                                 // [L] when keys_in_smem
    __shared__ unsigned s_prefix, s_k;
    const float* v = vals + (size_t)blockIdx.x * L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_prefix = 0u; s_k = (unsigned)k; }
    if (keys_in_smem)
        for (int i = tid; i < L; i += KTH_THREADS) keys[i] = f2ord(__ldg(v + i));
    const int Lr = (L + KTH_THREADS - 1) / KTH_THREADS * KTH_THREADS;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < KTH_WARPS * 256; i += KTH_THREADS) whist[i] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const unsigned mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < Lr; i += KTH_THREADS) {            // warp-uniform trip count (ballot / match below)
            unsigned u = 0u;
            bool in = false;
            if (i < L) {
                u = keys_in_smem ? keys[i] : f2ord(__ldg(v + i));
                in = (u & mask) == prefix;
            }
            const unsigned act = __ballot_sync(0xffffffffu, in);
            if (in) {
                const unsigned bin = (u >> shift) & 255u;
                const unsigned same = __match_any_sync(act, bin);
                if (lane == __ffs(same) - 1) atomicAdd(&whist[warp * 256 + bin], (unsigned)__popc(same));
            }
        }
        __syncthreads();
        if (tid < 256) {
            unsigned t = 0u;
#pragma unroll 8
            for (int w = 0; w < KTH_WARPS; ++w) t += whist[w * 256 + tid];
            hist[tid] = t;
        }
        __syncthreads();
        if (warp == 0) {
            // lane l owns the bins 255 - 8 l ... 248 - 8 l (descending); keys above a lane's bins = exclusive prefix
            const unsigned need = s_k;
            unsigned own = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) own += hist[255 - 8 * lane - j];
            unsigned incl = own;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, incl >= need);   // the k-th key exists: some lane reaches `need`
            const int owner = hit ? __ffs(hit) - 1 : 31;
            if (lane == owner) {
                unsigned run = incl - own;
                int b = 255 - 8 * lane;
                for (int j = 0; j < 7; ++j, --b) {
                    if (run + hist[b] >= need) break;
                    run += hist[b];
                }
                s_k = need - run;
                s_prefix = prefix | ((unsigned)b << shift);
            }
        }
        __syncthreads();
    }
    if (tid == 0) out[blockIdx.x] = ord2f(s_prefix);
}

// y[n,m] = act(sum_k x[n,k]*W[m,k] + b[m]); act 0 = identity, 1 = 1 + tanh(.)   (one warp per output)
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                      const float* __restrict__ b, float* __restrict__ y, int N,
                                                      int M, int K, int ldx, int ldy, int act) {
    AOC_PDL_TRIGGER();
    int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= N * M) return;
    int n = warp / M, m = warp - n * M;
    const float* xr = x + (size_t)n * ldx;
    const float* wr = W + (size_t)m * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(xr + k), __ldg(wr + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
        float v = acc + (b ? __ldg(b + m) : 0.f);
        if (act == 1) v = 1.0f + tanhf(v);
        y[(size_t)n * ldy + m] = v;
    }
}

// out[n, c] = sum_n' v[n', c] - v[n, c]   (conditioning_block x_delta / decoder _delta_head), written at ld/offset
__global__ void delta_sum_kernel(const float* __restrict__ v, float* __restrict__ out, int N, int C, int ldo) {
    AOC_PDL_TRIGGER();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += v[(size_t)n * C + c];
    for (int n = 0; n < N; ++n) out[(size_t)n * ldo + c] = s - v[(size_t)n * C + c];
}

}  // namespace aoc

using namespace aoc;

extern "C" int aoc_cond_phi_f32(const float* x, const float* w, const float* b, float* phi, int N, int HW, int C,
                                int ldx, cudaStream_t stream) {
    AOC_CHECK_ARG(x && w && phi, "null pointer");
    AOC_CHECK_ARG(C % 4 == 0 && ldx % 4 == 0, "C/ldx must be multiples of 4");
    long long NP = (long long)N * HW;
    int blocks = (int)((NP + 7) / 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    cond_phi_kernel<<<blocks, 256, 0, stream>>>(x, w, b, phi, NP, C, ldx);
    return launch_status("aoc_cond_phi_f32");
}

extern "C" int aoc_kth_largest_f32(const float* vals, int N, int L, int k, float* out, cudaStream_t stream) {
    AOC_CHECK_ARG(vals && out, "null pointer");
    AOC_CHECK_ARG(N > 0 && L > 0 && k >= 1 && k <= L, "k must be in [1, L]");
    const size_t fixed = (size_t)(KTH_WARPS * 256 + 256) * sizeof(unsigned);
    const int in_smem = fixed + (size_t)L * 4 <= 200 * 1024;
    const size_t smem = fixed + (in_smem ? (size_t)L * 4 : 0);
    static PerDeviceOnce attr;
    if (attr.first()) cudaFuncSetAttribute(kth_largest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    kth_largest_kernel<<<N, KTH_THREADS, smem, stream>>>(vals, L, k, out, in_smem);
    return launch_status("aoc_kth_largest_f32");
}

extern "C" int aoc_linear_f32(const float* x, const float* W, const float* b, float* y, int N, int M, int K, int ldx,
                              int ldy, int act, cudaStream_t stream) {
    AOC_CHECK_ARG(x && W && y, "null pointer");
    AOC_CHECK_ARG(N > 0 && M > 0 && K > 0 && (act == 0 || act == 1), "bad dims");
    long long warps = (long long)N * M;
    int blocks = (int)((warps * 32 + 255) / 256);
    linear_kernel<<<blocks, 256, 0, stream>>>(x, W, b, y, N, M, K, ldx, ldy, act);
    return launch_status("aoc_linear_f32");
}

extern "C" int aoc_delta_sum_f32(const float* v, float* out, int N, int C, int ldo, cudaStream_t stream) {
    AOC_CHECK_ARG(v && out, "null pointer");
    delta_sum_kernel<<<cdiv(C, 256), 256, 0, stream>>>(v, out, N, C, ldo);
    return launch_status("aoc_delta_sum_f32");
}

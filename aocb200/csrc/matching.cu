// Matching kernels of the AOC-Net per-frame path (fp32, exact path).
//
//   bank_*            : object-sorted reference bank built from (embeddings, uint8 label ids) of all memory frames.
//                       Replaces the one-hot gather of matching.py:2486-2495 / :533-545 (label 125 -> no object ->
//                       dropped).  Rows of one object keep frame-major raster order (needed by the host-drawn
//                       k-means init indices and by the centroid_avg indexing quirk, matching.py:589).
//   global_match      : matching.py:2384-2510 (+ :63-91, :27-46): per query pixel and object,
//                       min_n (|q|^2+|r_n|^2 - 2 q.r_n + 5e4*[label_n != o]), then 2*sigmoid(d+bias)-1.
//   proxy_match       : cluster level (matching.py:602-637) + k=1 proxy level (matching.py:149-197,2518-2662).
//   head_pool         : attention.py:155-189 (masked means of embeddings per object, eps=1e-5).
//   local_match       : matching.py:2710-2851 (25x25 window on the half-resolution grid, 6 nested windows).
//   prehead_assemble  : foreground2background (matching.py:9-23) + channel concat (aocnet.py:349-358).
#include "common.cuh"

namespace aoc {

constexpr int MAXO = AOC_MAX_OBJECTS;
constexpr int EMB = 100;  // cfg.MODEL_SEMANTIC_EMBEDDING_DIM
constexpr int EMB4 = 25;

// ------------------------------------------------------------------------------------------------
// bank build
// ------------------------------------------------------------------------------------------------
// blk_cnt[b][o] = number of pixels with id == o in flat pixels [b*256, b*256+256)
__global__ void __launch_bounds__(256) bank_count_kernel(const uint8_t* __restrict__ ids, int total, int O,
                                                          int* __restrict__ blk_cnt) {
    __shared__ int h[MAXO];
    if (threadIdx.x < MAXO) h[threadIdx.x] = 0;
    __syncthreads();
    int g = blockIdx.x * 256 + threadIdx.x;
    if (g < total) {
        int id = ids[g];
        if (id < O) atomicAdd(&h[id], 1);
    }
    __syncthreads();
    if (threadIdx.x < O) blk_cnt[(size_t)blockIdx.x * O + threadIdx.x] = h[threadIdx.x];
}

// One block.  In-place exclusive scan of blk_cnt over blocks (per object); nat_off[b] = #valid pixels before block b.
// meta: [0..O) counts, [MAXO..MAXO+O] segment offsets (each segment padded to `align` rows), [2*MAXO+1] = total rows
// of the sorted bank (padded), [2*MAXO+2] = total valid pixels.
__global__ void __launch_bounds__(1024) bank_scan_kernel(int* __restrict__ blk_cnt, int NB, int O, int align,
                                                          int* __restrict__ nat_off, int* __restrict__ meta) {
    __shared__ int tot[MAXO];
    __shared__ int wsum[32];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per-object scans: warp w handles objects w, w+32, ...
    for (int o = warp; o < O; o += 32) {
        int run = 0;
        for (int b0 = 0; b0 < NB; b0 += 32) {
            int b = b0 + lane;
            int v = b < NB ? blk_cnt[(size_t)b * O + o] : 0;
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (b < NB) blk_cnt[(size_t)b * O + o] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) tot[o] = run;
    }
    __syncthreads();
    // natural (all-object) offsets: nat_off[b] = sum_o excl_scan[b][o]  (each excl scan counts pixels before b)
    for (int b = threadIdx.x; b < NB; b += blockDim.x) {
        int s = 0;
        for (int o = 0; o < O; ++o) s += blk_cnt[(size_t)b * O + o];
        nat_off[b] = s;
    }
    if (threadIdx.x == 0) {
        int off = 0, valid = 0;
        for (int o = 0; o < O; ++o) {
            meta[o] = tot[o];
            meta[MAXO + o] = off;
            off += (tot[o] + align - 1) / align * align;
            valid += tot[o];
        }
        meta[MAXO + O] = off;
        meta[2 * MAXO + 1] = off;
        meta[2 * MAXO + 2] = valid;
    }
    (void)wsum;
}

// row_src[dest] = flat pixel index feeding sorted row `dest` (-1 for padding rows, pre-filled by the host wrapper);
// nat2sorted[t] = sorted row of the t-th valid pixel in natural (frame-major raster) order.
__global__ void __launch_bounds__(256) bank_scatter_kernel(const uint8_t* __restrict__ ids, int total, int O,
                                                            const int* __restrict__ blk_off,
                                                            const int* __restrict__ nat_off,
                                                            const int* __restrict__ meta, int* __restrict__ row_src,
                                                            int* __restrict__ nat2sorted) {
    __shared__ int wcnt[8][MAXO + 1];  // per-warp counts per object; [MAXO] = valid count
    int g = blockIdx.x * 256 + threadIdx.x;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int id = (g < total) ? (int)ids[g] : 255;
    bool valid = id < O;
    unsigned lt = (1u << lane) - 1u;
    int my_rank = 0;
    for (int o = 0; o < O; ++o) {
        unsigned m = __ballot_sync(0xffffffffu, id == o);
        if (id == o) my_rank = __popc(m & lt);
        if (lane == 0) wcnt[warp][o] = __popc(m);
    }
    unsigned mv = __ballot_sync(0xffffffffu, valid);
    int my_nat = __popc(mv & lt);
    if (lane == 0) wcnt[warp][MAXO] = __popc(mv);
    __syncthreads();
    if (valid) {
        int pre = 0, npre = 0;
        for (int w = 0; w < warp; ++w) { pre += wcnt[w][id]; npre += wcnt[w][MAXO]; }
        int dest = meta[MAXO + id] + blk_off[(size_t)blockIdx.x * O + id] + pre + my_rank;
        row_src[dest] = g;
        nat2sorted[nat_off[blockIdx.x] + npre + my_nat] = dest;
    }
}

// S[dest,:] = emb_all[row_src[dest],:] (zeros for padding rows), r2[dest] = |row|^2 (+inf for padding rows).
__global__ void __launch_bounds__(256) bank_gather_kernel(const float* __restrict__ emb_all,
                                                           const int* __restrict__ row_src, int rows,
                                                           float* __restrict__ S, float* __restrict__ r2) {
    int lane = threadIdx.x & 31;
    int row = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (row >= rows) return;
    int src = row_src[row];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src >= 0 && lane < EMB4) v = ldg4(emb_all + (size_t)src * EMB + lane * 4);
    if (lane < EMB4) *reinterpret_cast<float4*>(S + (size_t)row * EMB + lane * 4) = v;
    float s = v.x * v.x;
    s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    s = warp_sum(s);
    if (lane == 0) r2[row] = src >= 0 ? s : INFINITY;
}

__global__ void fill_i32_kernel(int* p, int v, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// global (pixel-level) matching, fp32 SIMT tile kernel: block = 128 queries x one object's segment.
// ------------------------------------------------------------------------------------------------
constexpr int GM_LD = 132;
__global__ void __launch_bounds__(256) global_match_simt_kernel(const float* __restrict__ q, int HW,
                                                                 const float* __restrict__ S,
                                                                 const float* __restrict__ r2,
                                                                 const int* __restrict__ meta, int O,
                                                                 float* __restrict__ mins /*[HW][O]*/) {
    extern __shared__ __align__(16) float smf[];
    float* Qs = smf;                     // [EMB][GM_LD]
    float* Rs = smf + EMB * GM_LD;       // [EMB][GM_LD]
    __shared__ float q2s[128], r2s[128];
    const int o = blockIdx.y;
    const int q0 = blockIdx.x * 128;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n_o = meta[o];
    const int seg = meta[MAXO + o];

    // load query tile transposed
    for (int i = tid; i < 128 * EMB4; i += 256) {
        int r = i / EMB4, c4 = i - r * EMB4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < HW) v = ldg4(q + (size_t)(q0 + r) * EMB + c4 * 4);
        Qs[(c4 * 4 + 0) * GM_LD + r] = v.x; Qs[(c4 * 4 + 1) * GM_LD + r] = v.y;
        Qs[(c4 * 4 + 2) * GM_LD + r] = v.z; Qs[(c4 * 4 + 3) * GM_LD + r] = v.w;
    }
    __syncthreads();
    if (tid < 128) {
        float s = 0.f;
        for (int c = 0; c < EMB; ++c) { float v = Qs[c * GM_LD + tid]; s = fmaf(v, v, s); }
        q2s[tid] = s;
    }
    float rmin[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rmin[i] = INFINITY;

    for (int t0 = 0; t0 < n_o; t0 += 128) {
        __syncthreads();
        for (int i = tid; i < 128 * EMB4; i += 256) {
            int r = i / EMB4, c4 = i - r * EMB4;
            float4 v = ldg4(S + (size_t)(seg + t0 + r) * EMB + c4 * 4);  // segments are padded to 128 rows
            Rs[(c4 * 4 + 0) * GM_LD + r] = v.x; Rs[(c4 * 4 + 1) * GM_LD + r] = v.y;
            Rs[(c4 * 4 + 2) * GM_LD + r] = v.z; Rs[(c4 * 4 + 3) * GM_LD + r] = v.w;
        }
        if (tid < 128) r2s[tid] = r2[seg + t0 + tid];
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
        for (int c = 0; c < EMB; ++c) {
            float4 a0 = *reinterpret_cast<const float4*>(Qs + c * GM_LD + ty * 4);
            float4 a1 = *reinterpret_cast<const float4*>(Qs + c * GM_LD + 64 + ty * 4);
            float4 b0 = *reinterpret_cast<const float4*>(Rs + c * GM_LD + tx * 4);
            float4 b1 = *reinterpret_cast<const float4*>(Rs + c * GM_LD + 64 + tx * 4);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
            float qq = q2s[r];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int cidx = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
                float t = qq + r2s[cidx];                 // (|q|^2 + |r|^2) - 2 q.r   (matching.py:45)
                float d = fmaf(-2.0f, acc[i][j], t);
                rmin[i] = fminf(rmin[i], d);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = rmin[i];
        v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1));
        v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 2));
        v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 4));
        v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 8));
        int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
        if (tx == 0 && q0 + r < HW) mins[(size_t)(q0 + r) * O + o] = v;
    }
}

// out[q][o] = 2*sigmoid(min(m_o, 5e4 + min_{o' != o} m_o') + bias_o) - 1; all-empty bank -> 1 (matching.py:2492-2493)
__global__ void global_match_finalize_kernel(const float* __restrict__ mins, const int* __restrict__ meta,
                                             const float* __restrict__ bias, int HW, int O, float* __restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    if (meta[2 * MAXO + 2] == 0) {
        for (int o = 0; o < O; ++o) out[(size_t)p * O + o] = 1.0f;
        return;
    }
    float m[MAXO];
    float m1 = INFINITY, m2 = INFINITY;
    int a1 = -1;
    for (int o = 0; o < O; ++o) {
        float v = meta[o] > 0 ? mins[(size_t)p * O + o] : INFINITY;
        m[o] = v;
        if (v < m1) { m2 = m1; m1 = v; a1 = o; }
        else if (v < m2) { m2 = v; }
    }
    for (int o = 0; o < O; ++o) {
        float other = (o == a1) ? m2 : m1;
        float d = fminf(m[o], other + AOC_WRONG_LABEL_PAD);
        out[(size_t)p * O + o] = sig2(d + __ldg(bias + o));
    }
}

// ------------------------------------------------------------------------------------------------
// proxy / cluster matching.  P: [O][36][EMB] (0..15 k-means centroids, 16..31 centroid_avg, 32 mean proxy),
// pvalid: [O][36] ints.  One thread per query pixel, proxies of one object at a time in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int NPX = AOC_PROXY_SLOTS;  // 36
__global__ void __launch_bounds__(128) proxy_match_kernel(const float* __restrict__ q, int HW,
                                                           const float* __restrict__ P,
                                                           const int* __restrict__ pvalid,
                                                           const float* __restrict__ bias, int O,
                                                           float* __restrict__ out_cluster /*[HW][O][2]*/,
                                                           float* __restrict__ out_proxy /*[HW][O]*/) {
    extern __shared__ __align__(16) float smf[];
    float* Qs = smf;                 // [EMB][128]
    float* Ps = smf + EMB * 128;     // [EMB][NPX]
    __shared__ float p2[NPX];
    __shared__ int pv[NPX];
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * 128;
    for (int i = tid; i < 128 * EMB4; i += 128) {
        int r = i / EMB4, c4 = i - r * EMB4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < HW) v = ldg4(q + (size_t)(q0 + r) * EMB + c4 * 4);
        Qs[(c4 * 4 + 0) * 128 + r] = v.x; Qs[(c4 * 4 + 1) * 128 + r] = v.y;
        Qs[(c4 * 4 + 2) * 128 + r] = v.z; Qs[(c4 * 4 + 3) * 128 + r] = v.w;
    }
    __syncthreads();
    float q2 = 0.f;
    for (int c = 0; c < EMB; ++c) { float v = Qs[c * 128 + tid]; q2 = fmaf(v, v, q2); }
    for (int o = 0; o < O; ++o) {
        __syncthreads();
        for (int i = tid; i < NPX * EMB; i += 128) {
            int j = i / EMB, c = i - j * EMB;
            Ps[c * NPX + j] = __ldg(P + ((size_t)o * NPX + j) * EMB + c);
        }
        if (tid < NPX) pv[tid] = pvalid[o * NPX + tid];
        __syncthreads();
        if (tid < NPX) {
            float s = 0.f;
            for (int c = 0; c < EMB; ++c) { float v = Ps[c * NPX + tid]; s = fmaf(v, v, s); }
            p2[tid] = s;
        }
        __syncthreads();
        float acc[NPX];
#pragma unroll
        for (int j = 0; j < NPX; ++j) acc[j] = 0.f;
        for (int c = 0; c < EMB; ++c) {
            float qv = Qs[c * 128 + tid];
#pragma unroll
            for (int j4 = 0; j4 < NPX / 4; ++j4) {
                float4 pp = *reinterpret_cast<const float4*>(Ps + c * NPX + j4 * 4);
                acc[j4 * 4 + 0] = fmaf(qv, pp.x, acc[j4 * 4 + 0]);
                acc[j4 * 4 + 1] = fmaf(qv, pp.y, acc[j4 * 4 + 1]);
                acc[j4 * 4 + 2] = fmaf(qv, pp.z, acc[j4 * 4 + 2]);
                acc[j4 * 4 + 3] = fmaf(qv, pp.w, acc[j4 * 4 + 3]);
            }
        }
        float m0 = INFINITY, m1 = INFINITY, dp = 0.f;
#pragma unroll
        for (int j = 0; j < 33; ++j) {
            float d = fmaf(-2.0f, acc[j], q2 + p2[j]);
            if (j < 16) { if (pv[j]) m0 = fminf(m0, d); }
            else if (j < 32) { if (pv[j]) m1 = fminf(m1, d); }
            else dp = d;
        }
        if (m0 == INFINITY) m0 = AOC_WRONG_LABEL_PAD;   // object without proxies: matching.py:619-620
        if (m1 == INFINITY) m1 = AOC_WRONG_LABEL_PAD;
        if (q0 + tid < HW) {
            float b = __ldg(bias + o);
            size_t idx = (size_t)(q0 + tid) * O + o;
            out_cluster[idx * 2 + 0] = sig2(m0 + b);
            out_cluster[idx * 2 + 1] = sig2(m1 + b);
            out_proxy[idx] = sig2(dp + b);
        }
    }
}

// The same matching for a proxy table of any width (kmax = 64: P [O][2*kmax+4][EMB]): the slots are walked in chunks of
// 32 (centroid chunks, then centroid_avg chunks, then the mean proxy), per-proxy arithmetic identical to the kernel above.
__global__ void __launch_bounds__(128) proxy_match_wide_kernel(const float* __restrict__ q, int HW,
                                                                const float* __restrict__ P,
                                                                const int* __restrict__ pvalid,
                                                                const float* __restrict__ bias, int O, int kmax,
                                                                float* __restrict__ out_cluster,
                                                                float* __restrict__ out_proxy) {
    extern __shared__ __align__(16) float smf[];
    float* Qs = smf;                 // [EMB][128]
    float* Ps = smf + EMB * 128;     // [EMB][32]
    __shared__ float p2[32];
    __shared__ int pv[32];
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * 128;
    const int slots = 2 * kmax + 4;
    for (int i = tid; i < 128 * EMB4; i += 128) {
        int r = i / EMB4, c4 = i - r * EMB4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < HW) v = ldg4(q + (size_t)(q0 + r) * EMB + c4 * 4);
        Qs[(c4 * 4 + 0) * 128 + r] = v.x; Qs[(c4 * 4 + 1) * 128 + r] = v.y;
        Qs[(c4 * 4 + 2) * 128 + r] = v.z; Qs[(c4 * 4 + 3) * 128 + r] = v.w;
    }
    __syncthreads();
    float q2 = 0.f;
    for (int c = 0; c < EMB; ++c) { float v = Qs[c * 128 + tid]; q2 = fmaf(v, v, q2); }
    for (int o = 0; o < O; ++o) {
        float m0 = INFINITY, m1 = INFINITY, dp = 0.f;
        for (int s0 = 0; s0 <= 2 * kmax; s0 += 32) {          // last chunk: the mean proxy alone
            const int n = min(32, 2 * kmax + 1 - s0);
            __syncthreads();
            for (int i = tid; i < 32 * EMB; i += 128) {
                int j = i / EMB, c = i - j * EMB;
                Ps[c * 32 + j] = j < n ? __ldg(P + ((size_t)o * slots + s0 + j) * EMB + c) : 0.f;
            }
            if (tid < 32) pv[tid] = tid < n ? pvalid[o * slots + s0 + tid] : 0;
            __syncthreads();
            if (tid < 32) {
                float s = 0.f;
                for (int c = 0; c < EMB; ++c) { float v = Ps[c * 32 + tid]; s = fmaf(v, v, s); }
                p2[tid] = s;
            }
            __syncthreads();
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.f;
            for (int c = 0; c < EMB; ++c) {
                float qv = Qs[c * 128 + tid];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float4 pp = *reinterpret_cast<const float4*>(Ps + c * 32 + j4 * 4);
                    acc[j4 * 4 + 0] = fmaf(qv, pp.x, acc[j4 * 4 + 0]);
                    acc[j4 * 4 + 1] = fmaf(qv, pp.y, acc[j4 * 4 + 1]);
                    acc[j4 * 4 + 2] = fmaf(qv, pp.z, acc[j4 * 4 + 2]);
                    acc[j4 * 4 + 3] = fmaf(qv, pp.w, acc[j4 * 4 + 3]);
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float d = fmaf(-2.0f, acc[j], q2 + p2[j]);
                if (s0 == 2 * kmax) { if (j == 0) dp = d; }
                else if (s0 < kmax) { if (pv[j]) m0 = fminf(m0, d); }
                else                { if (pv[j]) m1 = fminf(m1, d); }
            }
        }
        if (m0 == INFINITY) m0 = AOC_WRONG_LABEL_PAD;
        if (m1 == INFINITY) m1 = AOC_WRONG_LABEL_PAD;
        if (q0 + tid < HW) {
            float b = __ldg(bias + o);
            size_t idx = (size_t)(q0 + tid) * O + o;
            out_cluster[idx * 2 + 0] = sig2(m0 + b);
            out_cluster[idx * 2 + 1] = sig2(m1 + b);
            out_proxy[idx] = sig2(dp + b);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// head pooling: per-object sums of embeddings (+ total) over `total` flat pixels.
// part: [nblk][MAXO+1][EMB] floats, pcnt: [nblk][MAXO] ints.
// ------------------------------------------------------------------------------------------------
constexpr int HP_PIX = 256;
__global__ void __launch_bounds__(256) head_pool_partial_kernel(const float* __restrict__ emb,
                                                                 const uint8_t* __restrict__ ids, int total, int O,
                                                                 float* __restrict__ part, int* __restrict__ pcnt) {
    extern __shared__ __align__(16) float acc[];  // [8 warps][O+1][EMB]
    __shared__ int cnt[8][MAXO];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slots = O + 1;
    for (int i = threadIdx.x; i < 8 * slots * EMB; i += 256) acc[i] = 0.f;
    if (threadIdx.x < 8 * MAXO) (&cnt[0][0])[threadIdx.x] = 0;
    __syncthreads();
    float* my = acc + (size_t)warp * slots * EMB;
    int p0 = blockIdx.x * HP_PIX;
    int pend = min(p0 + HP_PIX, total);
    // four pixels per iteration: ids and embedding rows requested together (one dependent L2 round trip per pixel made
    // this 38 us for a 20 MB bank), accumulated in pixel order
    for (int pb = p0 + warp; pb < pend; pb += 32) {
        int id[4];
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = pb + 8 * u;
            id[u] = p < pend ? (int)ids[p] : 255;
            v[u] = (p < pend && lane < EMB4) ? ldg4(emb + (size_t)p * EMB + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (pb + 8 * u >= pend) break;
            if (lane < EMB4) {
                float4* t = reinterpret_cast<float4*>(my + O * EMB + lane * 4);
                float4 a = *t; a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; *t = a;
                if (id[u] < O) {
                    float4* d = reinterpret_cast<float4*>(my + id[u] * EMB + lane * 4);
                    float4 e = *d; e.x += v[u].x; e.y += v[u].y; e.z += v[u].z; e.w += v[u].w; *d = e;
                }
            }
            if (lane == 0 && id[u] < O) cnt[warp][id[u]] += 1;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < slots * EMB; i += 256) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += acc[(size_t)w * slots * EMB + i];
        part[(size_t)blockIdx.x * (MAXO + 1) * EMB + i] = s;
    }
    if (threadIdx.x < O) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += cnt[w][threadIdx.x];
        pcnt[(size_t)blockIdx.x * MAXO + threadIdx.x] = s;
    }
}

// pos[o][c] = sum_pos/(n_pos+eps); neg[o][c] = (sum_total-sum_pos)/((total-n_pos)+eps)    (attention.py:169-186)
// written into head[o][off_pos + c] and head[o][off_neg + c] (row stride ld_head); pos also to pos_out if given.
__global__ void __launch_bounds__(1024) head_pool_final_kernel(const float* __restrict__ part, const int* __restrict__ pcnt,
                                                                int nblk, int total, int O, float eps, float* __restrict__ head,
                                                                int ld_head, int off_pos, int off_neg,
                                                                float* __restrict__ pos_out, int ld_pos) {
    // block = 128 channel threads x 8 block lanes: lane l adds the partials of blocks l, l + 8, ... (two in flight), the eight
    // lane sums are combined in lane order (fixed tree): 13 dependent steps for a 2-frame 480p bank instead of 200
    __shared__ double s_sp[8][128], s_st[8][128];
    __shared__ long long s_np[8][128];
    const int o = blockIdx.x;
    const int c = threadIdx.x, l = threadIdx.y;
    double sp = 0.0, st = 0.0;
    long long np = 0;
    if (c < EMB) {
        for (int b = l; b < nblk; b += 16) {
            const int b1 = b + 8;
            const bool on1 = b1 < nblk;
            const float p0 = part[((size_t)b * (MAXO + 1) + o) * EMB + c], t0 = part[((size_t)b * (MAXO + 1) + O) * EMB + c];
            const float p1 = on1 ? part[((size_t)b1 * (MAXO + 1) + o) * EMB + c] : 0.f;
            const float t1 = on1 ? part[((size_t)b1 * (MAXO + 1) + O) * EMB + c] : 0.f;
            const int n0 = pcnt[(size_t)b * MAXO + o], n1 = on1 ? pcnt[(size_t)b1 * MAXO + o] : 0;
            sp += (double)p0; st += (double)t0; np += n0;
            sp += (double)p1; st += (double)t1; np += n1;
        }
    }
    s_sp[l][c] = sp; s_st[l][c] = st; s_np[l][c] = np;
    __syncthreads();
    if (l != 0 || c >= EMB) return;
    sp = 0.0; st = 0.0; np = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sp += s_sp[k][c]; st += s_st[k][c]; np += s_np[k][c]; }
    float fsp = (float)sp, fst = (float)st;
    float pos = fsp / ((float)np + eps);
    float neg = (fst - fsp) / ((float)((long long)total - np) + eps);
    head[(size_t)o * ld_head + off_pos + c] = pos;
    head[(size_t)o * ld_head + off_neg + c] = neg;
    if (pos_out) pos_out[(size_t)o * ld_pos + c] = pos;
}

// ------------------------------------------------------------------------------------------------
// local matching on the half-resolution grid: one warp per query pixel, lanes over window columns.
// ------------------------------------------------------------------------------------------------
__global__ void row_sqnorm_kernel(const float* __restrict__ x, int rows, float* __restrict__ out) {
    int lane = threadIdx.x & 31;
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    float s = 0.f;
    if (lane < EMB4) {
        float4 v = ldg4(x + (size_t)row * EMB + lane * 4);
        s = v.x * v.x; s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

constexpr int LM_R = 12;      // cfg.MODEL_MULTI_LOCAL_DISTANCE[-1]
constexpr int LM_RINGS = 7;   // ring = ceil(chebyshev/2): windows 2,4,..,12 are unions of rings 0..1, 0..2, ...
constexpr int LM_Q = 16;                  // query pixels per block: consecutive in x on one row, one warp each
constexpr int LM_W = LM_Q + 2 * LM_R;     // previous-frame pixels of one window row seen by the block (40)
constexpr int LM_LD = 101;                // staged row stride: lane <-> pixel reads are bank-conflict free

// order-preserving float <-> int map, so that a shared-memory atomicMin on ints is an exact float minimum
__device__ __forceinline__ int lm_f2o(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float lm_o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// Windowed correlation without unfold (matching.py:2710-2755): warp = query pixel, LANE = window column.  For each of
// the <= 25 window rows the block stages the 40 previous-frame pixels its 16 queries can see (one contiguous, coalesced
// 16 KB read) and every lane runs its own 100-channel dot product out of shared memory (query row broadcast as
// float4) -- no cross-lane reduction per neighbour, which cost 10 shuffle/add pairs per 4 FMAs in the first version
// (300 us per call at 61x107).  The per-(object, ring) minima are exact atomic minima in shared memory (order free).
__global__ void __launch_bounds__(LM_Q * 32) local_match_kernel(const float* __restrict__ xq, const float* __restrict__ yp,
                                                                const float* __restrict__ x2, const float* __restrict__ y2,
                                                                const uint8_t* __restrict__ ids, int hh, int ww, int O,
                                                                const float* __restrict__ bias, float* __restrict__ out,
                                                                int ld_out) {
    __shared__ __align__(16) float ys[LM_W * LM_LD];
    __shared__ __align__(16) float qs[LM_Q][EMB];
    __shared__ int mins[LM_Q][MAXO][LM_RINGS + 1];
    __shared__ float y2s[LM_W];
    __shared__ int ids_s[LM_W];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int py = blockIdx.y, px0 = blockIdx.x * LM_Q;
    const int px = px0 + warp;
    const bool qvalid = px < ww;
    const int p = py * ww + px;
    for (int i = lane; i < MAXO * (LM_RINGS + 1); i += 32) (&mins[warp][0][0])[i] = lm_f2o(AOC_WRONG_LABEL_PAD);
    if (qvalid && lane < EMB4) *reinterpret_cast<float4*>(&qs[warp][lane * 4]) = ldg4(xq + (size_t)p * EMB + lane * 4);
    const float xx = qvalid ? __ldg(x2 + p) : 0.f;
    const int y_lo = max(py - LM_R, 0), y_hi = min(py + LM_R, hh - 1);
    const int sx0 = px0 - LM_R;                                       // image x of staged column 0
    const int c_lo = max(sx0, 0), c_hi = min(px0 + LM_Q - 1 + LM_R, ww - 1);
    const int ncol = c_hi - c_lo + 1;
    const int nx = px - LM_R + lane;                                  // this lane's window column
    const bool active = qvalid && lane < 2 * LM_R + 1 && nx >= 0 && nx < ww;
    const int sidx = nx - sx0;
    const int adx = abs(nx - px);
    for (int ny = y_lo; ny <= y_hi; ++ny) {
        __syncthreads();                                              // the previous row has been consumed
        const size_t row0 = (size_t)ny * ww + c_lo;
        for (int i = threadIdx.x; i < ncol * EMB4; i += LM_Q * 32) {
            const int pix = i / EMB4, c4 = i - pix * EMB4;
            const float4 v = ldg4(yp + (row0 + pix) * EMB + c4 * 4);
            float* d = ys + (c_lo - sx0 + pix) * LM_LD + c4 * 4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        if (threadIdx.x < ncol) {
            y2s[c_lo - sx0 + threadIdx.x] = __ldg(y2 + row0 + threadIdx.x);
            ids_s[c_lo - sx0 + threadIdx.x] = ids[row0 + threadIdx.x];
        }
        __syncthreads();
        if (active) {
            const float* yr = ys + sidx * LM_LD;
            const float4* q4 = reinterpret_cast<const float4*>(&qs[warp][0]);
            float s = 0.f;
#pragma unroll 5
            for (int c4 = 0; c4 < EMB4; ++c4) {
                const float4 q = q4[c4];
                s = fmaf(q.x, yr[c4 * 4 + 0], s); s = fmaf(q.y, yr[c4 * 4 + 1], s);
                s = fmaf(q.z, yr[c4 * 4 + 2], s); s = fmaf(q.w, yr[c4 * 4 + 3], s);
            }
            const int id = ids_s[sidx];
            if (id < O) {
                const int ring = (max(abs(ny - py), adx) + 1) >> 1;
                const float d = fmaf(-2.0f, s, xx + y2s[sidx]);       // (x2 + y2) - 2 x.y   (matching.py:2754)
                atomicMin(&mins[warp][id][ring], lm_f2o(d));
            }
        }
    }
    __syncwarp();
    if (!qvalid) return;
    // channel 0 = full window (12), channels 1..5 = windows 2,4,6,8,10   (matching.py:2820-2836)
    for (int i = lane; i < O * 6; i += 32) {
        int o = i / 6, ch = i - o * 6;
        int rmax = ch == 0 ? LM_RINGS - 1 : ch;
        float m = AOC_WRONG_LABEL_PAD;
        for (int r = 0; r <= rmax; ++r) m = fminf(m, lm_o2f(mins[warp][o][r]));
        out[(size_t)p * ld_out + i] = sig2(m + __ldg(bias + o));
    }
}

// ------------------------------------------------------------------------------------------------
// pre-head input: [O][HW][24] = [global, cluster x2, proxy, local x6, local_proxy x6, prev one-hot,
//                                local bg x6, global bg]   (aocnet.py:349-358)
// ------------------------------------------------------------------------------------------------
__global__ void prehead_assemble_kernel(const float* __restrict__ g, const float* __restrict__ gc,
                                        const float* __restrict__ gp, const float* __restrict__ loc,
                                        const float* __restrict__ locp, int ld_loc,
                                        const uint8_t* __restrict__ prev_ids, int HW, int O,
                                        float* __restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    // foreground2background: min over the OTHER objects (top-2 trick); O == 1 returns the map itself.
    float m1[7], m2[7];
    int a1[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) { m1[c] = INFINITY; m2[c] = INFINITY; a1[c] = -1; }
    for (int o = 0; o < O; ++o) {
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            float v = c < 6 ? loc[(size_t)p * ld_loc + o * 6 + c] : g[(size_t)p * O + o];
            if (v < m1[c]) { m2[c] = m1[c]; m1[c] = v; a1[c] = o; }
            else if (v < m2[c]) { m2[c] = v; }
        }
    }
    int pid = prev_ids[p];
    for (int o = 0; o < O; ++o) {
        float* d = out + ((size_t)o * HW + p) * 24;
        float v[24];
        v[0] = g[(size_t)p * O + o];
        v[1] = gc[((size_t)p * O + o) * 2 + 0];
        v[2] = gc[((size_t)p * O + o) * 2 + 1];
        v[3] = gp[(size_t)p * O + o];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            v[4 + c] = loc[(size_t)p * ld_loc + o * 6 + c];
            v[10 + c] = locp[(size_t)p * ld_loc + o * 6 + c];
        }
        v[16] = (pid == o) ? 1.0f : 0.0f;
#pragma unroll
        for (int c = 0; c < 6; ++c) v[17 + c] = (O == 1) ? v[4 + c] : ((o == a1[c]) ? m2[c] : m1[c]);
        v[23] = (O == 1) ? v[0] : ((o == a1[6]) ? m2[6] : m1[6]);
#pragma unroll
        for (int c4 = 0; c4 < 6; ++c4)
            *reinterpret_cast<float4*>(d + c4 * 4) = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
    }
}

// y[n,p,0:C] = x[p,0:C] for n < N  (cat of the current embedding in front of every object slot, aocnet.py:362)
__global__ void broadcast_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int HW, int C,
                                      int ldx, int ldy) {
    int C4 = C >> 2;
    long long total = (long long)HW * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C4) * 4;
        long long p = i / C4;
        float4 v = ldg4(x + (size_t)p * ldx + c);
        for (int n = 0; n < N; ++n) *reinterpret_cast<float4*>(y + ((size_t)n * HW + p) * ldy + c) = v;
    }
}

}  // namespace aoc

using namespace aoc;

extern "C" size_t aoc_bank_workspace_bytes(int total_pixels, int O) {
    size_t NB = (size_t)cdiv(total_pixels, 256);
    return (NB * O + NB) * sizeof(int) + 256;
}

// Builds the object-sorted bank.  meta_out (device, (2*MAXO+3) int32) must be copied to the host by the caller to
// learn the counts.  row_src/nat2sorted: int32 [cap_rows]/[total_pixels].  cap_rows >= total_pixels + O*align.
extern "C" int aoc_bank_index_build(const uint8_t* ids, int total_pixels, int O, int align, int* meta_out,
                                    int* row_src, int cap_rows, int* nat2sorted, void* workspace, size_t ws_bytes,
                                    cudaStream_t stream) {
    AOC_CHECK_ARG(ids && meta_out && row_src && nat2sorted && workspace, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= MAXO, "O out of range");
    AOC_CHECK_ARG(align >= 1 && total_pixels > 0, "bad dims");
    AOC_CHECK_ARG(cap_rows >= total_pixels + O * align, "row_src capacity too small");
    AOC_CHECK_ARG(ws_bytes >= aoc_bank_workspace_bytes(total_pixels, O), "workspace too small");
    int NB = cdiv(total_pixels, 256);
    int* blk = (int*)workspace;
    int* nat_off = blk + (size_t)NB * O;
    bank_count_kernel<<<NB, 256, 0, stream>>>(ids, total_pixels, O, blk);
    bank_scan_kernel<<<1, 1024, 0, stream>>>(blk, NB, O, align, nat_off, meta_out);
    fill_i32_kernel<<<cdiv(cap_rows, 1024), 256, 0, stream>>>(row_src, -1, cap_rows);
    bank_scatter_kernel<<<NB, 256, 0, stream>>>(ids, total_pixels, O, blk, nat_off, meta_out, row_src, nat2sorted);
    return launch_status("aoc_bank_index_build");
}

extern "C" int aoc_bank_gather_f32(const float* emb_all, const int* row_src, int rows, float* S, float* r2,
                                   cudaStream_t stream) {
    AOC_CHECK_ARG(emb_all && row_src && S && r2, "null pointer");
    if (rows == 0) return AOC_OK;
    bank_gather_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, stream>>>(emb_all, row_src, rows, S, r2);
    return launch_status("aoc_bank_gather_f32");
}

extern "C" int aoc_global_match_simt_f32(const float* q, int HW, const float* S, const float* r2, const int* meta,
                                         const float* bias, int O, float* mins_ws, float* out, cudaStream_t stream) {
    AOC_CHECK_ARG(q && S && r2 && meta && bias && mins_ws && out, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= MAXO && HW > 0, "bad dims");
    static PerDeviceOnce attr_done;
    size_t smem = (size_t)2 * EMB * GM_LD * sizeof(float);
    if (attr_done.first()) {
        cudaFuncSetAttribute(global_match_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    dim3 grid(cdiv(HW, 128), O);
    global_match_simt_kernel<<<grid, 256, smem, stream>>>(q, HW, S, r2, meta, O, mins_ws);
    global_match_finalize_kernel<<<cdiv(HW, 256), 256, 0, stream>>>(mins_ws, meta, bias, HW, O, out);
    return launch_status("aoc_global_match_simt_f32");
}

extern "C" int aoc_global_match_finalize_f32(const float* mins, const int* meta, const float* bias, int HW, int O,
                                             float* out, cudaStream_t stream) {
    AOC_CHECK_ARG(mins && meta && bias && out, "null pointer");
    global_match_finalize_kernel<<<cdiv(HW, 256), 256, 0, stream>>>(mins, meta, bias, HW, O, out);
    return launch_status("aoc_global_match_finalize_f32");
}

extern "C" int aoc_proxy_match_f32(const float* q, int HW, const float* P, const int* pvalid, const float* bias,
                                   int O, int kmax, float* out_cluster, float* out_proxy, cudaStream_t stream) {
    AOC_CHECK_ARG(q && P && pvalid && bias && out_cluster && out_proxy, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= MAXO && HW > 0, "bad dims");
    AOC_CHECK_ARG(kmax == 16 || kmax == AOC_KMEANS_MAX_K, "kmax must be 16 or 64");
    if (kmax != 16) {
        static PerDeviceOnce attr_w;
        size_t smem_w = (size_t)(EMB * 128 + EMB * 32) * sizeof(float);
        if (attr_w.first())
            cudaFuncSetAttribute(proxy_match_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
        proxy_match_wide_kernel<<<cdiv(HW, 128), 128, smem_w, stream>>>(q, HW, P, pvalid, bias, O, kmax, out_cluster,
                                                                      out_proxy);
        return launch_status("aoc_proxy_match_f32");
    }
    static PerDeviceOnce attr_done;
    size_t smem = (size_t)(EMB * 128 + EMB * NPX) * sizeof(float);
    if (attr_done.first()) {
        cudaFuncSetAttribute(proxy_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    // (one object per block -- grid.y = O, 1 212 blocks instead of 202 at 9 % occupancy -- was measured: 85 -> 98 us, the
    // re-staged query tile and the third partial wave cost more than the occupancy gains)
    proxy_match_kernel<<<cdiv(HW, 128), 128, smem, stream>>>(q, HW, P, pvalid, bias, O, out_cluster, out_proxy);
    return launch_status("aoc_proxy_match_f32");
}

extern "C" size_t aoc_head_pool_workspace_bytes(int total_pixels) {
    size_t nblk = (size_t)cdiv(total_pixels, HP_PIX);
    return nblk * ((MAXO + 1) * EMB * sizeof(float) + MAXO * sizeof(int));
}

extern "C" int aoc_head_pool_f32(const float* emb, const uint8_t* ids, int total_pixels, int O, float eps,
                                 float* head, int ld_head, int off_pos, int off_neg, float* pos_out, int ld_pos,
                                 void* workspace, size_t ws_bytes, cudaStream_t stream) {
    AOC_CHECK_ARG(emb && ids && head && workspace, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= MAXO && total_pixels > 0, "bad dims");
    AOC_CHECK_ARG(ws_bytes >= aoc_head_pool_workspace_bytes(total_pixels), "workspace too small");
    int nblk = cdiv(total_pixels, HP_PIX);
    float* part = (float*)workspace;
    int* pcnt = (int*)(part + (size_t)nblk * (MAXO + 1) * EMB);
    size_t smem = (size_t)8 * (O + 1) * EMB * sizeof(float);
    static PerDeviceOnce attr_done;
    if (attr_done.first()) {
        cudaFuncSetAttribute(head_pool_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)((size_t)8 * (MAXO + 1) * EMB * sizeof(float)));
    }
    head_pool_partial_kernel<<<nblk, 256, smem, stream>>>(emb, ids, total_pixels, O, part, pcnt);
    head_pool_final_kernel<<<O, dim3(128, 8), 0, stream>>>(part, pcnt, nblk, total_pixels, O, eps, head, ld_head, off_pos,
                                                  off_neg, pos_out, ld_pos);
    return launch_status("aoc_head_pool_f32");
}

extern "C" int aoc_row_sqnorm_f32(const float* x, int rows, float* out, cudaStream_t stream) {
    AOC_CHECK_ARG(x && out && rows > 0, "bad args");
    row_sqnorm_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, stream>>>(x, rows, out);
    return launch_status("aoc_row_sqnorm_f32");
}

extern "C" int aoc_local_match_f32(const float* xq, const float* yp, const float* x2, const float* y2,
                                   const uint8_t* ids, int hh, int ww, int O, const float* bias, float* out,
                                   int ld_out, cudaStream_t stream) {
    AOC_CHECK_ARG(xq && yp && x2 && y2 && ids && bias && out, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= MAXO && hh > 0 && ww > 0 && ld_out >= 6 * O, "bad dims");
    local_match_kernel<<<dim3(cdiv(ww, LM_Q), hh), LM_Q * 32, 0, stream>>>(xq, yp, x2, y2, ids, hh, ww, O, bias, out, ld_out);
    return launch_status("aoc_local_match_f32");
}

extern "C" int aoc_prehead_assemble_f32(const float* g, const float* gc, const float* gp, const float* loc,
                                        const float* locp, int ld_loc, const uint8_t* prev_ids, int HW, int O,
                                        float* out, cudaStream_t stream) {
    AOC_CHECK_ARG(g && gc && gp && loc && locp && prev_ids && out, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= MAXO && HW > 0, "bad dims");
    prehead_assemble_kernel<<<cdiv(HW, 128), 128, 0, stream>>>(g, gc, gp, loc, locp, ld_loc, prev_ids, HW, O, out);
    return launch_status("aoc_prehead_assemble_f32");
}

extern "C" int aoc_broadcast_rows_f32(const float* x, float* y, int N, int HW, int C, int ldx, int ldy,
                                      cudaStream_t stream) {
    AOC_CHECK_ARG(x && y && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "bad args");
    long long total = (long long)HW * (C / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    broadcast_rows_kernel<<<blocks, 256, 0, stream>>>(x, y, N, HW, C, ldx, ldy);
    return launch_status("aoc_broadcast_rows_f32");
}

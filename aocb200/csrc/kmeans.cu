// Adaptive object proxies: per-object Lloyd k-means over the object's bank embeddings.
//
// Replaces scipy.cluster.vq.kmeans2(X_i, k, minit='points', iter=20) as called per object per frame at
// networks/layers/matching.py:562 (inside _nearest_neighbor_features_per_object_in_chunks_cluster2, :506-640),
// its GPU->CPU->GPU round trip, and the `centroid_avg` construction at :589.
//   * init rows come from the HOST (np.random.choice on numpy's global RNG -- same stream as scipy's _kpoints);
//   * `iters` fixed Lloyd rounds: assign = argmin_j ((-2 x.c_j) + |x|^2) + |c_j|^2, lowest index wins ties;
//     update = mean of assigned rows; an empty cluster keeps its previous centroid;
//   * returned labels are those of the last round (computed against the code book before its final update);
//   * centroid_avg[j] = mean_{t : label_i[t] == j} B[t], B = ALL-object bank in natural order (the reference's
//     indexing quirk), only for non-empty labels.
// HBM/L2-bound streaming kernels: row tiles staged in shared memory with coalesced 128-bit loads, sequential-order
// dot products per row (the order of the reference's BLAS call, so near ties break the same way), fixed-order tile sums +
// a warp-parallel second-stage reduction (deterministic); no tensor cores, no atomics.
#include "common.cuh"

namespace aoc {

constexpr int MAXO = AOC_MAX_OBJECTS;
constexpr int EMB = 100;
constexpr int EMB4 = 25;
constexpr int KM_K = AOC_KMEANS_MAX_K;   // 16

// cent: [O][KM_K][EMB]
__global__ void kmeans_init_kernel(const float* __restrict__ S, const int* __restrict__ meta,
                                   const int* __restrict__ kk, const int* __restrict__ init_idx,
                                   float* __restrict__ cent) {
    int o = blockIdx.x;
    int k = kk[o];
    int seg = meta[MAXO + o];
    for (int i = threadIdx.x; i < KM_K * EMB; i += blockDim.x) {
        int j = i / EMB, c = i - j * EMB;
        float v = 0.f;
        if (j < k) v = S[(size_t)(seg + init_idx[o * KM_K + j]) * EMB + c];
        cent[(size_t)o * KM_K * EMB + i] = v;
    }
}

// One Lloyd half-step over a tile of KM_TILE rows of one object (block = 256 threads):
//   1. the tile is staged in shared memory with coalesced 128-bit loads (row stride 101 floats: conflict-free both for
//      the row-per-thread reads of step 2 and the channel-per-thread reads of step 3);
//   2. ASSIGN: thread (row, half) accumulates the dot products with centroids 8*half .. 8*half+7 SEQUENTIALLY over the
//      channels, fma by fma -- the summation order of a BLAS sgemm element, i.e. of scipy's vq, so near-tie rows get the
//      same label as in the reference; distance (-2 x.c + |x|^2) + |c|^2; lowest index wins ties;
//   3. per-label column sums: thread (label parity, channel) walks the rows in order and adds into its own column of a
//      shared accumulator (fixed order, no atomics -> deterministic).
// INDIRECT (centroid_avg pass, matching.py:589): label t of the object selects row nat2sorted[t] of the ALL-object bank.
// part[o][b][KM_K][EMB] / pcnt[o][b][KM_K] receive the tile's per-label sums and counts.
constexpr int KM_TILE = 128;
constexpr int KM_LD = 101;
constexpr int KM_SMEM = (KM_TILE * KM_LD + EMB * KM_K + KM_K * EMB) * 4;   // tile + centroids + accumulators

template <bool ASSIGN, bool INDIRECT>
__global__ void __launch_bounds__(256) kmeans_step_kernel(const float* __restrict__ S, const int* __restrict__ meta,
                                                           const int* __restrict__ kk, const float* __restrict__ cent,
                                                           const int* __restrict__ nat2sorted,
                                                           int* __restrict__ labels /*[sorted rows]*/,
                                                           float* __restrict__ part, int* __restrict__ pcnt,
                                                           int nb_max) {
    extern __shared__ __align__(16) float km_sm[];
    float* tile = km_sm;                                   // [KM_TILE][KM_LD]
    float* Cs = tile + KM_TILE * KM_LD;                    // [EMB][KM_K]
    float* acc = Cs + EMB * KM_K;                          // [KM_K][EMB]
    __shared__ float c2[KM_K];
    __shared__ int lab[KM_TILE];
    __shared__ float bd_hi[KM_TILE];
    __shared__ int bj_hi[KM_TILE];
    const int o = blockIdx.y, b = blockIdx.x;
    const int n_o = meta[o];
    const int k = kk[o];
    const int t0 = b * KM_TILE;
    if (t0 >= n_o || k <= 0) return;
    const int seg = meta[MAXO + o];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nrow = min(KM_TILE, n_o - t0);

    // ---- 1. stage the tile: warp w loads rows w, w+8, ... (25 lanes x 16 B = one 400 B row per instruction)
    // (eight rows per batch: all loads are issued before the first store, one L2 round trip per batch instead of per row)
#pragma unroll
    for (int batch = 0; batch < KM_TILE / 64; ++batch) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = warp + 8 * (batch * 8 + i);
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrow && lane < EMB4) {
                const int srow = INDIRECT ? __ldg(nat2sorted + t0 + r) : (seg + t0 + r);
                v[i] = ldg4(S + (size_t)srow * EMB + lane * 4);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = warp + 8 * (batch * 8 + i);
            if (r < nrow && lane < EMB4) {
                float* d = tile + r * KM_LD + lane * 4;
                d[0] = v[i].x; d[1] = v[i].y; d[2] = v[i].z; d[3] = v[i].w;
            }
        }
    }
    if (ASSIGN) {
        for (int i = tid; i < KM_K * EMB; i += 256) {
            int j = i / EMB, c = i - j * EMB;
            Cs[c * KM_K + j] = cent[(size_t)o * KM_K * EMB + i];
        }
    }
    __syncthreads();
    if (ASSIGN) {
        if (tid < KM_K) {
            float s = 0.f;
            for (int c = 0; c < EMB; ++c) s = fmaf(Cs[c * KM_K + tid], Cs[c * KM_K + tid], s);
            c2[tid] = s;
        }
        __syncthreads();
        // ---- 2. assignment: thread (row, half)
        const int row = tid & (KM_TILE - 1), half = tid >> 7;
        float bd = INFINITY;
        int bj = half * 8;
        if (row < nrow) {
            const float* xr = tile + row * KM_LD;
            const float4* cr = reinterpret_cast<const float4*>(Cs + half * 8);
            float dot[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) dot[j] = 0.f;
            float x2 = 0.f;
#pragma unroll 4
            for (int c = 0; c < EMB; ++c) {
                const float x = xr[c];
                x2 = fmaf(x, x, x2);
                const float4 ca = cr[c * (KM_K / 4)], cb = cr[c * (KM_K / 4) + 1];
                dot[0] = fmaf(x, ca.x, dot[0]); dot[1] = fmaf(x, ca.y, dot[1]);
                dot[2] = fmaf(x, ca.z, dot[2]); dot[3] = fmaf(x, ca.w, dot[3]);
                dot[4] = fmaf(x, cb.x, dot[4]); dot[5] = fmaf(x, cb.y, dot[5]);
                dot[6] = fmaf(x, cb.z, dot[6]); dot[7] = fmaf(x, cb.w, dot[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int jj = half * 8 + j;
                const float d = (dot[j] * -2.0f + x2) + c2[jj];
                if (jj < k && d < bd) { bd = d; bj = jj; }
            }
        }
        if (half == 1) { bd_hi[row] = bd; bj_hi[row] = bj; }
        __syncthreads();
        if (half == 0) {
            int best = -1;
            if (row < nrow) {
                if (bd_hi[row] < bd) bj = bj_hi[row];          // strict: the lower index wins ties
                best = bj;
                labels[seg + t0 + row] = best;
            }
            lab[row] = best;
        }
    } else {
        if (tid < KM_TILE) lab[tid] = (tid < nrow) ? labels[seg + t0 + tid] : -1;
    }
    __syncthreads();
    // ---- 3. per-label column sums in row order.  The rows of the tile are first counting-sorted by label (stable: warp
    // ballots give each row its rank among the rows of its label), then thread (label j, channel c) adds the rows of
    // label j in increasing row order into a REGISTER -- the same summation order as a read-modify-write per row on a
    // shared accumulator, without its store-to-load latency chain (which was ~60 % of the kernel's time).
    {
        int* order = reinterpret_cast<int*>(acc);                    // [KM_TILE] row ids grouped by label
        int* wcnt = order + KM_TILE;                                 // [4 warps][KM_K] rows of label j in warp w
        int* start = wcnt + 4 * KM_K;                                // [KM_K + 1]
        int l = -1, rank = 0;
        if (tid < KM_TILE) {
            l = lab[tid];
#pragma unroll
            for (int j = 0; j < KM_K; ++j) {
                const unsigned m = __ballot_sync(0xffffffffu, l == j);
                if (l == j) rank = __popc(m & ((1u << lane) - 1u));
                if (lane == 0) wcnt[warp * KM_K + j] = __popc(m);
            }
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int j = 0; j < KM_K; ++j) {
                start[j] = run;
                run += wcnt[j] + wcnt[KM_K + j] + wcnt[2 * KM_K + j] + wcnt[3 * KM_K + j];
            }
            start[KM_K] = run;
        }
        __syncthreads();
        if (l >= 0) {
            int pos = start[l] + rank;
            for (int w = 0; w < warp; ++w) pos += wcnt[w * KM_K + l];
            order[pos] = tid;
        }
        __syncthreads();
        const size_t pbase = ((size_t)o * nb_max + b) * KM_K;
        for (int i = tid; i < KM_K * EMB; i += 256) {
            const int j = i / EMB, c = i - j * EMB;
            float a = 0.f;
            for (int q = start[j]; q < start[j + 1]; ++q) a += tile[order[q] * KM_LD + c];
            part[pbase * EMB + i] = a;
        }
        if (tid < KM_K) pcnt[pbase + tid] = start[tid + 1] - start[tid];
    }
}

// Second stage, one block per (cluster j, object o): warp w sums the slabs b = w, w+32, ... (lanes over channels,
// 128-bit loads), the 32 warps are combined in a fixed order in double precision.
// MODE 0: Lloyd update (cent[j] = sum/cnt, empty keeps previous).  MODE 1: write centroid_avg + validity.
// P: [O][AOC_PROXY_SLOTS][EMB], pvalid: [O][AOC_PROXY_SLOTS]
constexpr int KR_WARPS = 32;     // the slab loop is one L2 round trip per iteration: 32 warps keep it to a handful
template <int MODE>
__global__ void __launch_bounds__(KR_WARPS * 32) kmeans_reduce_kernel(const float* __restrict__ part,
                                                             const int* __restrict__ pcnt,
                                                             const int* __restrict__ meta,
                                                             const int* __restrict__ kk, int nb_max,
                                                             int rows_per_block, float* __restrict__ cent,
                                                             float* __restrict__ P, int* __restrict__ pvalid) {
    __shared__ double sm[KR_WARPS][EMB];
    __shared__ int sn[KR_WARPS];
    const int j = blockIdx.x, o = blockIdx.y;
    const int n_o = meta[o];
    const int k = kk[o];
    const int nb = (n_o + rows_per_block - 1) / rows_per_block;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int n = 0;
    if (j < k) {
        for (int b = warp; b < nb; b += KR_WARPS) {
            const size_t pbase = ((size_t)o * nb_max + b) * KM_K + j;
            if (lane < EMB4) {
                float4 v = ldg4(part + pbase * EMB + lane * 4);
                s0 += (double)v.x; s1 += (double)v.y; s2 += (double)v.z; s3 += (double)v.w;
            }
            n += __ldg(pcnt + pbase);
        }
    }
    if (lane < EMB4) {
        sm[warp][lane * 4 + 0] = s0; sm[warp][lane * 4 + 1] = s1; sm[warp][lane * 4 + 2] = s2; sm[warp][lane * 4 + 3] = s3;
    }
    if (lane == 0) sn[warp] = n;
    __syncthreads();
    const int c = threadIdx.x;
    if (c >= EMB) return;
    double s = 0.0;
    long long cntj = 0;
    for (int w = 0; w < KR_WARPS; ++w) { s += sm[w][c]; cntj += sn[w]; }
    const int i = j * EMB + c;
    if (MODE == 0) {
        if (j < k && cntj > 0) cent[(size_t)o * KM_K * EMB + i] = (float)s / (float)cntj;
    } else {
        bool ok = (j < k) && cntj > 0;
        P[((size_t)o * AOC_PROXY_SLOTS + 16 + j) * EMB + c] = ok ? (float)s / (float)cntj : 0.f;
        P[((size_t)o * AOC_PROXY_SLOTS + j) * EMB + c] = (j < k) ? cent[(size_t)o * KM_K * EMB + i] : 0.f;
        if (c == 0) {
            pvalid[o * AOC_PROXY_SLOTS + j] = (j < k) ? 1 : 0;
            pvalid[o * AOC_PROXY_SLOTS + 16 + j] = ok ? 1 : 0;
        }
    }
}

static int km_rows_per_block(int) { return KM_TILE; }

}  // namespace aoc

using namespace aoc;

extern "C" size_t aoc_kmeans_workspace_bytes(int max_rows_per_object, int O) {
    int m = max_rows_per_object > 0 ? max_rows_per_object : 1;
    size_t nb = (size_t)cdiv(m, km_rows_per_block(m));
    return (size_t)O * nb * KM_K * (EMB * sizeof(float) + sizeof(int)) + 256;
}

// S/meta/nat2sorted from aoc_bank_*; kk[o] (device int32) = clusters of object o (0 = no proxies);
// init_idx (device int32 [O][16]) = object-local row indices drawn by the host RNG.
// Outputs: cent [O][16][100], labels (int32, per sorted row), P[O][36][100] slots 0..31 + pvalid[O][36] slots 0..31.
extern "C" int aoc_kmeans_proxies_f32(const float* S, const int* meta, const int* nat2sorted, const int* kk,
                                      const int* init_idx, int O, int max_rows_per_object, int iters, float* cent,
                                      int* labels, float* P, int* pvalid, void* workspace, size_t ws_bytes,
                                      cudaStream_t stream) {
    AOC_CHECK_ARG(S && meta && nat2sorted && kk && init_idx && cent && labels && P && pvalid && workspace,
                  "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= MAXO && iters >= 1, "bad dims");
    AOC_CHECK_ARG(ws_bytes >= aoc_kmeans_workspace_bytes(max_rows_per_object, O), "workspace too small");
    const int m = max_rows_per_object > 0 ? max_rows_per_object : 1;
    const int rpb = km_rows_per_block(m);
    int nb = cdiv(m, rpb);
    float* part = (float*)workspace;
    int* pcnt = (int*)(part + (size_t)O * nb * KM_K * EMB);
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(kmeans_step_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, KM_SMEM);
        cudaFuncSetAttribute(kmeans_step_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KM_SMEM);
    }
    kmeans_init_kernel<<<O, 256, 0, stream>>>(S, meta, kk, init_idx, cent);
    dim3 gs(nb, O), gr(KM_K, O);
    for (int it = 0; it < iters; ++it) {
        kmeans_step_kernel<true, false><<<gs, 256, KM_SMEM, stream>>>(S, meta, kk, cent, nat2sorted, labels, part, pcnt, nb);
        kmeans_reduce_kernel<0><<<gr, KR_WARPS * 32, 0, stream>>>(part, pcnt, meta, kk, nb, rpb, cent, P, pvalid);
    }
    // centroid_avg: same labels, rows taken from the all-object bank in natural order (matching.py:589)
    kmeans_step_kernel<false, true><<<gs, 256, KM_SMEM, stream>>>(S, meta, kk, cent, nat2sorted, labels, part, pcnt, nb);
    kmeans_reduce_kernel<1><<<gr, KR_WARPS * 32, 0, stream>>>(part, pcnt, meta, kk, nb, rpb, cent, P, pvalid);
    return launch_status("aoc_kmeans_proxies_f32");
}

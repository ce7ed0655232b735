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
// HBM/L2-bound streaming kernels: one thread per row for the assignment (centroids broadcast from shared
// memory), fixed-order per-block partial sums + a second-stage reduction (deterministic), no tensor cores.
#include "common.cuh"

namespace aoc {

constexpr int MAXO = AOC_MAX_OBJECTS;
constexpr int EMB = 100;
constexpr int EMB4 = 25;
constexpr int KM_K = AOC_KMEANS_MAX_K;   // 16
constexpr int KM_PTS = 256;              // rows per block

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

// ASSIGN: compute labels of this block's rows against cent, store them.  Then accumulate per-label sums of
// (INDIRECT ? S[nat2sorted[t]] : S[seg + t]) in row order into part[o][b][KM_K][EMB], counts into pcnt[o][b][KM_K].
template <bool ASSIGN, bool INDIRECT>
__global__ void __launch_bounds__(256) kmeans_step_kernel(const float* __restrict__ S, const int* __restrict__ meta,
                                                           const int* __restrict__ kk,
                                                           const float* __restrict__ cent,
                                                           const int* __restrict__ nat2sorted,
                                                           int* __restrict__ labels /*[sorted rows]*/,
                                                           float* __restrict__ part, int* __restrict__ pcnt,
                                                           int nb_max) {
    __shared__ __align__(16) float Cs[EMB][KM_K];
    __shared__ float c2[KM_K];
    __shared__ int lab[KM_PTS];
    __shared__ float acc[2][KM_K][EMB];
    const int o = blockIdx.y, b = blockIdx.x;
    const int n_o = meta[o];
    const int k = kk[o];
    const int t0 = b * KM_PTS;
    if (t0 >= n_o || k <= 0) return;
    const int seg = meta[MAXO + o];
    const int tid = threadIdx.x;
    const int t = t0 + tid;

    if (ASSIGN) {
        for (int i = tid; i < KM_K * EMB; i += 256) {
            int j = i / EMB, c = i - j * EMB;
            Cs[c][j] = cent[(size_t)o * KM_K * EMB + i];
        }
        __syncthreads();
        if (tid < KM_K) {
            float s = 0.f;
            for (int c = 0; c < EMB; ++c) s = fmaf(Cs[c][tid], Cs[c][tid], s);
            c2[tid] = s;
        }
        __syncthreads();
        int best = 0;
        if (t < n_o) {
            const float* row = S + (size_t)(seg + t) * EMB;
            float dot[KM_K];
#pragma unroll
            for (int j = 0; j < KM_K; ++j) dot[j] = 0.f;
            float x2 = 0.f;
#pragma unroll 5
            for (int c4 = 0; c4 < EMB4; ++c4) {
                float4 v = ldg4(row + c4 * 4);
                float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    x2 = fmaf(xs[e], xs[e], x2);
                    const float4* cr = reinterpret_cast<const float4*>(&Cs[c4 * 4 + e][0]);
#pragma unroll
                    for (int j4 = 0; j4 < KM_K / 4; ++j4) {
                        float4 cc = cr[j4];
                        dot[j4 * 4 + 0] = fmaf(xs[e], cc.x, dot[j4 * 4 + 0]);
                        dot[j4 * 4 + 1] = fmaf(xs[e], cc.y, dot[j4 * 4 + 1]);
                        dot[j4 * 4 + 2] = fmaf(xs[e], cc.z, dot[j4 * 4 + 2]);
                        dot[j4 * 4 + 3] = fmaf(xs[e], cc.w, dot[j4 * 4 + 3]);
                    }
                }
            }
            float bd = INFINITY;
#pragma unroll
            for (int j = 0; j < KM_K; ++j) {
                float d = (dot[j] * -2.0f + x2) + c2[j];
                if (j < k && d < bd) { bd = d; best = j; }
            }
            labels[seg + t] = best;
        }
        lab[tid] = (t < n_o) ? best : -1;
    } else {
        lab[tid] = (t < n_o) ? labels[seg + t] : -1;
    }
    for (int i = tid; i < 2 * KM_K * EMB; i += 256) (&acc[0][0][0])[i] = 0.f;
    __syncthreads();
    // fixed-order accumulation: thread (half, c) walks its half of the block's rows in order
    {
        int half = tid >> 7, c = tid & 127;
        if (c < EMB) {
            int pbeg = half * (KM_PTS / 2), pend = pbeg + KM_PTS / 2;
            for (int p = pbeg; p < pend; ++p) {
                int l = lab[p];
                if (l < 0) break;
                int srow = INDIRECT ? nat2sorted[t0 + p] : (seg + t0 + p);
                acc[half][l][c] += __ldg(S + (size_t)srow * EMB + c);
            }
        }
    }
    __syncthreads();
    size_t pbase = ((size_t)o * nb_max + b) * KM_K;
    for (int i = tid; i < KM_K * EMB; i += 256) {
        int j = i / EMB, c = i - j * EMB;
        part[pbase * EMB + i] = acc[0][j][c] + acc[1][j][c];
    }
    if (tid < KM_K) {
        int n = 0;
        for (int p = 0; p < KM_PTS; ++p) n += (lab[p] == tid);
        pcnt[pbase + tid] = n;
    }
}

// MODE 0: Lloyd update (cent[j] = sum/cnt, empty keeps previous).  MODE 1: write centroid_avg + validity.
// P: [O][AOC_PROXY_SLOTS][EMB], pvalid: [O][AOC_PROXY_SLOTS]
template <int MODE>
__global__ void __launch_bounds__(256) kmeans_reduce_kernel(const float* __restrict__ part,
                                                             const int* __restrict__ pcnt,
                                                             const int* __restrict__ meta,
                                                             const int* __restrict__ kk, int nb_max,
                                                             float* __restrict__ cent, float* __restrict__ P,
                                                             int* __restrict__ pvalid) {
    const int o = blockIdx.y;
    const int n_o = meta[o];
    const int k = kk[o];
    const int nb = (n_o + KM_PTS - 1) / KM_PTS;
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= KM_K * EMB) return;
    int j = i / EMB, c = i - j * EMB;
    double s = 0.0;
    long long n = 0;
    if (j < k) {
        for (int b = 0; b < nb; ++b) {
            size_t pbase = ((size_t)o * nb_max + b) * KM_K;
            s += (double)part[pbase * EMB + i];
            n += pcnt[pbase + j];
        }
    }
    if (MODE == 0) {
        if (j < k && n > 0) cent[(size_t)o * KM_K * EMB + i] = (float)s / (float)n;
    } else {
        bool ok = (j < k) && n > 0;
        P[((size_t)o * AOC_PROXY_SLOTS + 16 + j) * EMB + c] = ok ? (float)s / (float)n : 0.f;
        P[((size_t)o * AOC_PROXY_SLOTS + j) * EMB + c] = (j < k) ? cent[(size_t)o * KM_K * EMB + i] : 0.f;
        if (c == 0) {
            pvalid[o * AOC_PROXY_SLOTS + j] = (j < k) ? 1 : 0;
            pvalid[o * AOC_PROXY_SLOTS + 16 + j] = ok ? 1 : 0;
        }
    }
}

}  // namespace aoc

using namespace aoc;

extern "C" size_t aoc_kmeans_workspace_bytes(int max_rows_per_object, int O) {
    size_t nb = (size_t)cdiv(max_rows_per_object > 0 ? max_rows_per_object : 1, KM_PTS);
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
    int nb = cdiv(max_rows_per_object > 0 ? max_rows_per_object : 1, KM_PTS);
    float* part = (float*)workspace;
    int* pcnt = (int*)(part + (size_t)O * nb * KM_K * EMB);
    kmeans_init_kernel<<<O, 256, 0, stream>>>(S, meta, kk, init_idx, cent);
    dim3 gs(nb, O), gr(cdiv(KM_K * EMB, 256), O);
    for (int it = 0; it < iters; ++it) {
        kmeans_step_kernel<true, false><<<gs, 256, 0, stream>>>(S, meta, kk, cent, nat2sorted, labels, part, pcnt, nb);
        kmeans_reduce_kernel<0><<<gr, 256, 0, stream>>>(part, pcnt, meta, kk, nb, cent, P, pvalid);
    }
    // centroid_avg: same labels, rows taken from the all-object bank in natural order (matching.py:589)
    kmeans_step_kernel<false, true><<<gs, 256, 0, stream>>>(S, meta, kk, cent, nat2sorted, labels, part, pcnt, nb);
    kmeans_reduce_kernel<1><<<gr, 256, 0, stream>>>(part, pcnt, meta, kk, nb, cent, P, pvalid);
    return launch_status("aoc_kmeans_proxies_f32");
}

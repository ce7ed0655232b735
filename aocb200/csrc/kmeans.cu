// Adaptive object proxies: per-object Lloyd k-means over the object's bank embeddings -- ONE persistent cooperative kernel.
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
// Schedule (round 1 ran 2 * iters + 3 dependent launches of 128-row tiles, ~13 us each, latency bound):
//   the whole call is one launch; CTAs own row tiles (256 rows x 100 floats staged in shared memory with coalesced 128-bit
//   loads) and walk the Lloyd rounds with two grid-wide barriers per round (assign + per-tile partial sums | per-centroid
//   reduction of the tile sums).  A CTA that owns a single tile keeps it RESIDENT in shared memory for all rounds -- the
//   rows of a bank of <= 3 frames at 480p are read from L2/HBM exactly once per frame instead of once per round.
// Arithmetic order (what makes the labels bit-identical to the reference's on identical inputs): sequential-order dot
// products per row, fma by fma (the order of a BLAS sgemm element); per-label tile sums in row order after a stable
// counting sort (warp ballots); tile sums added in tile order in double.  No tensor cores, no atomics on data.
#include "common.cuh"

namespace aoc {

constexpr int MAXO = AOC_MAX_OBJECTS;
constexpr int EMB = 100;
constexpr int EMB4 = 25;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// two independent IEEE fp32 FMAs in one issue slot: (d0, d1) <- x * (c0, c1) + (d0, d1)
__device__ __forceinline__ void ffma2_bcast(float& d0, float& d1, float x, float c0, float c1) {
    asm("{\n\t.reg .b64 d, a, b;\n\tmov.b64 d, {%0, %1};\n\tmov.b64 a, {%2, %2};\n\tmov.b64 b, {%3, %4};\n\t"
        "fma.rn.f32x2 d, a, b, d;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "+f"(d0), "+f"(d1) : "f"(x), "f"(c0), "f"(c1));
}

struct KmP {
    const float* S; const int* meta; const int* nat2sorted; const int* kk; const int* init_idx;
    float* cent; int* labels; float* P; int* pvalid;
    long long* acc;     // [3][O][K][EMB] fixed-point (2^-32) per-label column sums, rotating over the Lloyd rounds
    int* cnt;           // [3][O][K] rows per label
    unsigned* bar;      // grid barrier counter (zeroed by the launcher); bar[1]: sticky range-error flag
    int O, iters;
};

// NP = K / 8 "parts": thread (row, part) owns the dot products with centroids 8*part .. 8*part+7 of its row
template <int NP, int TILE>
struct KmCfg {
    static constexpr int K = 8 * NP;
    static constexpr int THREADS = TILE * NP;
    static constexpr int NW = THREADS / 32;                 // warps of the CTA
    static constexpr int TW = TILE / 32;                    // warps holding one thread per row
    static constexpr int SLOTS = 2 * K + 4;                 // proxy slots per object: K centroids, K centroid_avg, mean, pad
    // shared memory (bytes).  Row r of the tile starts at float r*EMB + (r >> 3): the one-float skew per 8 rows makes the
    // row-per-thread reads of the assignment conflict-free (bank = 4 r + (r >> 3) + c mod 32 is a bijection over a warp).
    static constexpr int TILE_B = (TILE * EMB + TILE / 8) * 4;
    static constexpr int CS_B = EMB * K * 4;                // centroids [EMB][K]
    static constexpr int C2_B = K * 4;                      // |c_j|^2
    static constexpr int BD_B = (NP - 1) * TILE * 4;        // best distance of parts 1.. per row
    static constexpr int BJ_B = (NP - 1) * TILE;            // their argmin (uint8)
    static constexpr int LAB_B = TILE;                      // int8 label per row (-1: no row)
    static constexpr int ORD_B = TILE * 2;                  // uint16 row ids grouped by label
    static constexpr int WCNT_B = TW * K * 2;               // uint16 rows of label j in warp w
    static constexpr int START_B = (K + 1) * 4;
    static constexpr int TS_B = (MAXO + 1) * 4;
    static constexpr int OFF_CS = (TILE_B + 15) / 16 * 16;
    static constexpr int OFF_C2 = OFF_CS + CS_B;
    static constexpr int OFF_BD = OFF_C2 + C2_B;
    static constexpr int OFF_BJ = OFF_BD + BD_B;
    static constexpr int OFF_LAB = OFF_BJ + BJ_B;
    static constexpr int OFF_ORD = OFF_LAB + LAB_B;
    static constexpr int OFF_WCNT = OFF_ORD + ORD_B;
    static constexpr int OFF_START = (OFF_WCNT + WCNT_B + 3) / 4 * 4;
    static constexpr int OFF_TS = OFF_START + START_B;
    static constexpr int SMEM = OFF_TS + TS_B;
    static constexpr int MINB = NP <= 2 ? 3 : 1;            // CTAs per SM the register budget is held to
};

__device__ __forceinline__ void km_grid_sync(unsigned* bar, unsigned& epoch, unsigned G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                   // this CTA's global writes / reductions before the arrival
        atomicAdd(bar, 1u);
        ++epoch;
        const unsigned target = epoch * G;                 // the counter only grows: no reset between barriers
        while (ld_acquire_u32(bar) < target) { }
        __threadfence();
    }
    __syncthreads();
}

// Cross-tile sums.  Every tile adds its per-label column sums into ONE accumulator per (object, label, channel) with
// integer atomics on a 2^-32 fixed-point image of the fp32 tile sum: integer addition is associative, so the result does
// not depend on the order the tiles arrive in (bit-reproducible run to run, unlike floating-point atomics), there is no
// per-tile workspace and no second reduction phase -- one grid barrier per Lloyd round.  Resolution 2.3e-10 absolute (a
// tile sum of embedding values is O(1..1e3): far below one fp32 ulp of the result), range +-2.1e9; a tile sum beyond
// 2^30 raises the sticky error word instead of wrapping.
constexpr double KM_FIX = 4294967296.0;
__device__ __forceinline__ long long km_to_fix(float v) { return __double2ll_rn((double)v * KM_FIX); }
__device__ __forceinline__ double km_from_fix(long long a) { return (double)a * (1.0 / KM_FIX); }

template <int NP, int TILE>
__global__ void __launch_bounds__(KmCfg<NP, TILE>::THREADS, KmCfg<NP, TILE>::MINB) kmeans_persistent_kernel(KmP p) {
    using Cfg = KmCfg<NP, TILE>;
    constexpr int K = Cfg::K, THREADS = Cfg::THREADS, NW = Cfg::NW, TW = Cfg::TW;
    extern __shared__ __align__(16) unsigned char km_raw[];
    float* tile = reinterpret_cast<float*>(km_raw);
    float* Cs = reinterpret_cast<float*>(km_raw + Cfg::OFF_CS);             // [EMB][K]
    float* c2 = reinterpret_cast<float*>(km_raw + Cfg::OFF_C2);
    float* bd_s = reinterpret_cast<float*>(km_raw + Cfg::OFF_BD);
    uint8_t* bj_s = km_raw + Cfg::OFF_BJ;
    int8_t* lab = reinterpret_cast<int8_t*>(km_raw + Cfg::OFF_LAB);
    uint16_t* order = reinterpret_cast<uint16_t*>(km_raw + Cfg::OFF_ORD);
    uint16_t* wcnt = reinterpret_cast<uint16_t*>(km_raw + Cfg::OFF_WCNT);
    int* start = reinterpret_cast<int*>(km_raw + Cfg::OFF_START);
    int* tstart = reinterpret_cast<int*>(km_raw + Cfg::OFF_TS);             // first tile of object o; [O] = total

    __shared__ int ncnt[K];                                                  // rows per label of the previous round
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x, G = gridDim.x;
    const int O = p.O;
    const size_t ACC_N = (size_t)O * K * EMB, CNT_N = (size_t)O * K;
    if (tid == 0) {
        int run = 0;
        for (int o = 0; o < O; ++o) {
            tstart[o] = run;
            const int n_o = p.meta[o];
            if (p.kk[o] > 0 && n_o > 0) run += (n_o + TILE - 1) / TILE;
        }
        for (int o = O; o <= MAXO; ++o) tstart[o] = run;
    }
    __syncthreads();
    const int total = tstart[O];
    const bool resident = g + G >= total;                  // this CTA owns at most one tile: it stays in shared memory
    unsigned epoch = 0;
    bool loaded = false;

    // Code book of object o for round `it` into Cs (transposed) -- every CTA derives it itself from the accumulators of
    // round it-1: mean of the assigned rows, an empty cluster keeps its previous centroid (global `cent`), round 0 = the
    // host-drawn rows.  The CTA that owns the object's first tile also writes it to `cent` (the empty-cluster fallback of
    // the next round and, after the last round, the result).
    auto load_codebook = [&](int o, int it, bool owner) {
        const int k = p.kk[o];
        const long long* acc = p.acc + (size_t)((it + 2) % 3) * ACC_N + (size_t)o * K * EMB;    // buffer of round it-1
        const int* cnt = p.cnt + (size_t)((it + 2) % 3) * CNT_N + (size_t)o * K;
        // two L2 round trips per call instead of one per element: the label counts first, then every accumulator word
        // of this thread in flight before the first conversion (the call is on the critical path between two grid barriers)
        if (tid < K) ncnt[tid] = (it > 0 && tid < k) ? __ldcg(cnt + tid) : 0;
        __syncthreads();
        constexpr int PER = (K * EMB + THREADS - 1) / THREADS;
        long long a[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = tid + u * THREADS;
            a[u] = 0;
            if (it > 0 && i < K * EMB && ncnt[i / EMB] > 0) a[u] = __ldcg(acc + i);
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = tid + u * THREADS;
            if (i >= K * EMB) break;
            const int j = i / EMB, c = i - j * EMB;
            const size_t ci = (size_t)o * K * EMB + i;
            float v = 0.f;
            if (j < k) {
                if (it == 0) {
                    v = __ldg(p.S + (size_t)(p.meta[MAXO + o] + p.init_idx[o * K + j]) * EMB + c);
                } else {
                    const int n = ncnt[j];
                    v = n > 0 ? (float)km_from_fix(a[u]) / (float)n : __ldcg(p.cent + ci);
                }
            }
            Cs[c * K + j] = v;
            if (owner) p.cent[ci] = v;
        }
    };

    // ---- one pass over this CTA's tiles.  ASSIGN: Lloyd assignment (round `it`) + per-label sums of the object's own rows;
    //      !ASSIGN: per-label sums (same labels) of the rows nat2sorted[t] of the all-object bank (matching.py:589)
    auto tile_phase = [&](const bool ASSIGN, const int it) {
        long long* acc_out = p.acc + (size_t)(it % 3) * ACC_N;
        int* cnt_out = p.cnt + (size_t)(it % 3) * CNT_N;
        for (int T = g; T < total; T += G) {
            int o = 0;
            while (T >= tstart[o + 1]) ++o;
            const int b = T - tstart[o];
            const int n_o = p.meta[o], seg = p.meta[MAXO + o], k = p.kk[o];
            const int t0 = b * TILE;
            const int nrow = min(TILE, n_o - t0);
            // -- 1. stage the tile: warp w loads rows w, w + NW, ... (25 lanes x 16 B = one 400 B row per instruction),
            //       a batch of loads issued before the first store (one L2 round trip per batch)
            if (!(ASSIGN && resident && loaded)) {
                constexpr int RPW = TILE / NW;
                constexpr int BATCH = RPW < 8 ? RPW : 8;
#pragma unroll
                for (int r0 = 0; r0 < RPW; r0 += BATCH) {
                    float4 v[BATCH];
#pragma unroll
                    for (int i = 0; i < BATCH; ++i) {
                        const int r = warp + NW * (r0 + i);
                        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r < nrow && lane < EMB4) {
                            const int srow = ASSIGN ? (seg + t0 + r) : __ldg(p.nat2sorted + t0 + r);
                            v[i] = ldg4(p.S + (size_t)srow * EMB + lane * 4);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < BATCH; ++i) {
                        const int r = warp + NW * (r0 + i);
                        if (r < nrow && lane < EMB4) {
                            float* d = tile + r * EMB + (r >> 3) + lane * 4;
                            d[0] = v[i].x; d[1] = v[i].y; d[2] = v[i].z; d[3] = v[i].w;
                        }
                    }
                }
                loaded = ASSIGN;
            }
            if (ASSIGN) load_codebook(o, it, b == 0);
            __syncthreads();
            if (ASSIGN) {
                if (tid < K) {
                    float s = 0.f;
                    for (int c = 0; c < EMB; ++c) s = fmaf(Cs[c * K + tid], Cs[c * K + tid], s);
                    c2[tid] = s;
                }
                __syncthreads();
                // -- 2. assignment: thread (row, part); dot products accumulated SEQUENTIALLY over the channels, fma by
                //       fma -- the summation order of a BLAS sgemm element, i.e. of scipy's vq, so near-tie rows get the
                //       same label as in the reference; distance (-2 x.c + |x|^2) + |c|^2; lowest index wins ties
                const int row = tid & (TILE - 1), part = tid / TILE;
                float bd = INFINITY;
                int bj = part * 8;
                if (row < nrow) {
                    const float* xr = tile + row * EMB + (row >> 3);
                    const float4* cr = reinterpret_cast<const float4*>(Cs + part * 8);
                    float dot[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) dot[j] = 0.f;
                    float x2 = 0.f;
#pragma unroll 4
                    for (int c = 0; c < EMB; ++c) {
                        const float x = xr[c];
                        x2 = fmaf(x, x, x2);
                        const float4 ca = cr[c * (K / 4)], cb = cr[c * (K / 4) + 1];
                        ffma2_bcast(dot[0], dot[1], x, ca.x, ca.y);
                        ffma2_bcast(dot[2], dot[3], x, ca.z, ca.w);
                        ffma2_bcast(dot[4], dot[5], x, cb.x, cb.y);
                        ffma2_bcast(dot[6], dot[7], x, cb.z, cb.w);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int jj = part * 8 + j;
                        const float d = (dot[j] * -2.0f + x2) + c2[jj];
                        if (jj < k && d < bd) { bd = d; bj = jj; }
                    }
                }
                if (part > 0) { bd_s[(part - 1) * TILE + row] = bd; bj_s[(part - 1) * TILE + row] = (uint8_t)bj; }
                __syncthreads();
                if (part == 0) {
                    int best = -1;
                    if (row < nrow) {
#pragma unroll
                        for (int q = 0; q < NP - 1; ++q) {                   // strict: the lower index wins ties
                            const float d = bd_s[q * TILE + row];
                            if (d < bd) { bd = d; bj = bj_s[q * TILE + row]; }
                        }
                        best = bj;
                        p.labels[seg + t0 + row] = best;
                    }
                    lab[row] = (int8_t)best;
                }
            } else {
                if (tid < TILE) lab[tid] = (tid < nrow) ? (int8_t)p.labels[seg + t0 + tid] : (int8_t)-1;
            }
            __syncthreads();
            // -- 3. per-label column sums.  The rows are first counting-sorted by label (stable: warp ballots give each row
            //       its rank among the rows of its label), then thread (label j, channel c) adds the rows of label j in a
            //       fixed order (four interleaved chains: rows q, q+4, ... of the sorted list) into registers.
            int l = -1, rank = 0;
            if (tid < TILE) {
                l = lab[tid];
                for (int j = 0; j < K; ++j) {
                    const unsigned m = __ballot_sync(0xffffffffu, l == j);
                    if (l == j) rank = __popc(m & ((1u << lane) - 1u));
                    if (lane == 0) wcnt[warp * K + j] = (uint16_t)__popc(m);
                }
            }
            __syncthreads();
            if (tid == 0) {
                int run = 0;
                for (int j = 0; j < K; ++j) {
                    start[j] = run;
                    for (int w = 0; w < TW; ++w) run += wcnt[w * K + j];
                }
                start[K] = run;
            }
            __syncthreads();
            if (l >= 0) {
                int pos = start[l] + rank;
                for (int w = 0; w < warp; ++w) pos += wcnt[w * K + l];
                order[pos] = (uint16_t)tid;
            }
            __syncthreads();
            for (int i = tid; i < K * EMB; i += THREADS) {
                const int j = i / EMB, c = i - j * EMB;
                const int q0 = start[j], q1 = start[j + 1];
                if (q1 == q0) continue;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                int q = q0;
                for (; q + 3 < q1; q += 4) {
                    const int r0 = order[q], r1 = order[q + 1], r2 = order[q + 2], r3 = order[q + 3];
                    a0 += tile[r0 * EMB + (r0 >> 3) + c];
                    a1 += tile[r1 * EMB + (r1 >> 3) + c];
                    a2 += tile[r2 * EMB + (r2 >> 3) + c];
                    a3 += tile[r3 * EMB + (r3 >> 3) + c];
                }
                for (; q < q1; ++q) {
                    const int r = order[q];
                    a0 += tile[r * EMB + (r >> 3) + c];
                }
                const float a = (a0 + a1) + (a2 + a3);
                if (!(fabsf(a) < 1073741824.f)) p.bar[1] = 1u;      // outside the fixed-point range (or NaN)
                atomicAdd(reinterpret_cast<unsigned long long*>(acc_out + (size_t)o * K * EMB + i),
                          (unsigned long long)km_to_fix(a));
            }
            if (tid < K && start[tid + 1] > start[tid]) atomicAdd(cnt_out + o * K + tid, start[tid + 1] - start[tid]);
            __syncthreads();                               // the scratch arrays and (streaming) the tile are reused
        }
        // the accumulators of the NEXT round (last read before the previous barrier) are cleared here, before this
        // CTA arrives at this round's barrier
        long long* accz = p.acc + (size_t)((it + 1) % 3) * ACC_N;
        int* cntz = p.cnt + (size_t)((it + 1) % 3) * CNT_N;
        for (size_t i = (size_t)g * THREADS + tid; i < ACC_N; i += (size_t)G * THREADS) accz[i] = 0;
        for (size_t i = (size_t)g * THREADS + tid; i < CNT_N; i += (size_t)G * THREADS) cntz[i] = 0;
    };

    for (int it = 0; it < p.iters; ++it) {
        tile_phase(true, it);
        km_grid_sync(p.bar, epoch, G);
    }
    // final code book (the update after the last assignment) -> cent and the centroid slots of the proxy table
    for (int o = g; o < O; o += G) {
        __syncthreads();
        if (tstart[o + 1] > tstart[o]) load_codebook(o, p.iters, true);
        __syncthreads();
        const bool has = tstart[o + 1] > tstart[o];
        for (int i = tid; i < K * EMB; i += THREADS) {
            const int j = i / EMB, c = i - j * EMB;
            p.P[((size_t)o * Cfg::SLOTS + j) * EMB + c] = (has && j < p.kk[o]) ? Cs[c * K + j] : 0.f;
            if (!has) p.cent[(size_t)o * K * EMB + i] = 0.f;
        }
        if (tid < K) p.pvalid[o * Cfg::SLOTS + tid] = (has && tid < p.kk[o]) ? 1 : 0;
    }
    // centroid_avg: same labels, rows taken from the all-object bank in natural order (matching.py:589)
    tile_phase(false, p.iters);
    km_grid_sync(p.bar, epoch, G);
    {
        const long long* acc = p.acc + (size_t)(p.iters % 3) * ACC_N;
        const int* cnt = p.cnt + (size_t)(p.iters % 3) * CNT_N;
        for (size_t i = (size_t)g * THREADS + tid; i < ACC_N; i += (size_t)G * THREADS) {
            const int o = (int)(i / (K * EMB)), r = (int)(i - (size_t)o * K * EMB);
            const int j = r / EMB, c = r - j * EMB;
            const int n = __ldcg(cnt + o * K + j);
            const bool ok = j < p.kk[o] && n > 0 && tstart[o + 1] > tstart[o];
            float v = ok ? (float)km_from_fix(__ldcg(acc + i)) / (float)n : 0.f;
            if (__ldcg(p.bar + 1)) v = __int_as_float(0x7fc00000);     // a tile sum left the fixed-point range: poison, do not guess
            p.P[((size_t)o * Cfg::SLOTS + K + j) * EMB + c] = v;
            if (c == 0) p.pvalid[o * Cfg::SLOTS + K + j] = ok ? 1 : 0;
        }
    }
}

template <int NP, int TILE>
static int km_launch(const KmP& p, int max_tiles, cudaStream_t stream) {
    using Cfg = KmCfg<NP, TILE>;
    static int occ[AOC_MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= AOC_MAX_DEVICES) dev = 0;
    if (occ[dev] == 0) {
        cudaFuncSetAttribute(kmeans_persistent_kernel<NP, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        int n = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kmeans_persistent_kernel<NP, TILE>, Cfg::THREADS, Cfg::SMEM);
        occ[dev] = n > 0 ? n : -1;
    }
    if (occ[dev] < 0) {
        set_error("aoc_kmeans_proxies_f32: the persistent kernel does not fit an SM (%d B shared memory)", Cfg::SMEM);
        return AOC_ELAUNCH;
    }
    int grid = occ[dev] * device_sms();                   // every CTA must be resident: they meet at grid barriers
    const int want = max_tiles > p.O ? max_tiles : p.O;
    if (grid > want) grid = want;
    // barrier word, error word and the accumulators of rounds 0 and 2 (= "round -1": never read) start at zero
    cudaMemsetAsync(p.bar, 0, 256 + 3 * ((size_t)p.O * Cfg::K * EMB * sizeof(long long) + (size_t)p.O * Cfg::K * sizeof(int)), stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kmeans_persistent_kernel<NP, TILE>, p);
    return launch_status("aoc_kmeans_proxies_f32");
}

constexpr int KM_TILE16 = 128, KM_TILE64 = 128;
static int km_tile(int kmax) { return kmax <= 16 ? KM_TILE16 : KM_TILE64; }

}  // namespace aoc

using namespace aoc;

extern "C" size_t aoc_kmeans_workspace_bytes(int max_rows_per_object, int O, int kmax) {
    (void)max_rows_per_object;
    return 256 + 3 * ((size_t)O * kmax * EMB * sizeof(long long) + (size_t)O * kmax * sizeof(int));
}

// S/meta/nat2sorted from aoc_bank_*; kk[o] (device int32) = clusters of object o (0 = no proxies);
// init_idx (device int32 [O][kmax]) = object-local row indices drawn by the host RNG.  kmax = 16 or 64 (cluster_num <= kmax).
// Outputs: cent [O][kmax][100], labels (int32, per sorted row), P[O][2*kmax+4][100] slots 0..2*kmax-1 + pvalid likewise.
extern "C" int aoc_kmeans_proxies_f32(const float* S, const int* meta, const int* nat2sorted, const int* kk,
                                      const int* init_idx, int O, int max_rows_per_object, int iters, int kmax,
                                      float* cent, int* labels, float* P, int* pvalid, void* workspace, size_t ws_bytes,
                                      cudaStream_t stream) {
    AOC_CHECK_ARG(S && meta && nat2sorted && kk && init_idx && cent && labels && P && pvalid && workspace,
                  "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= MAXO && iters >= 1, "bad dims");
    AOC_CHECK_ARG(kmax == 16 || kmax == AOC_KMEANS_MAX_K, "kmax must be 16 or 64");
    AOC_CHECK_ARG(ws_bytes >= aoc_kmeans_workspace_bytes(max_rows_per_object, O, kmax), "workspace too small");
    AOC_CHECK_ARG((((uintptr_t)workspace) & 15) == 0, "workspace must be 16-byte aligned");
    const int m = max_rows_per_object > 0 ? max_rows_per_object : 1;
    const int tiles = O * cdiv(m, km_tile(kmax));
    KmP p;
    p.S = S; p.meta = meta; p.nat2sorted = nat2sorted; p.kk = kk; p.init_idx = init_idx;
    p.cent = cent; p.labels = labels; p.P = P; p.pvalid = pvalid;
    p.bar = (unsigned*)workspace;
    p.acc = (long long*)((char*)workspace + 256);
    p.cnt = (int*)(p.acc + 3 * (size_t)O * kmax * EMB);
    p.O = O; p.iters = iters;
    return kmax <= 16 ? km_launch<2, KM_TILE16>(p, tiles, stream) : km_launch<8, KM_TILE64>(p, tiles, stream);
}

// Pixel-level global matching on the tcgen05 tensor cores (the dominant dense contraction of the path):
//   for every query pixel and object:  min_n (|q|^2 + |r_n|^2 - 2 q.r_n)   over the object's bank rows,
// reference: networks/layers/matching.py:2384-2510 (-> :200-249, :63-91, :27-46), 132.85 GFLOP per reference frame
// at 480p.  The N_q x N_ref distance matrix is never materialised: a CTA owns 128 query rows (TMEM lanes), streams
// the object-sorted bank through a 6-stage TMA (cp.async.bulk) pipeline, accumulates q.r in TMEM with 3xTF32
// (fp32-faithful), and the epilogue warps keep one running minimum per thread straight out of tcgen05.ld -- no
// cross-thread reduction, no chunking (the reference's n_chunks loop disappears).
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..7 = epilogue (warp w reads TMEM lanes 32*(w%4)..+31).  Accumulators are double-buffered in TMEM
// (2 x 256 columns) so the epilogue of bank block i overlaps the MMAs of block i+1.
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace aoc {
using namespace umma;

constexpr int MAXO_ = AOC_MAX_OBJECTS;
constexpr int TC_K = 104;              // embedding width 100 zero-padded to a multiple of 8 (3xTF32 images)
constexpr int TC_KS = TC_K / KSTEP;    // 13 k-steps
constexpr int QB = 128;                // query rows per CTA (= TMEM lanes)
constexpr int RBK = 256;               // bank rows per MMA (N)
constexpr int NST = 6;                 // smem pipeline stages (one (row block, k-step) chunk each)
constexpr uint32_t B_STAGE = 2 * RBK * KSTEP * 4;            // 16384: hi + lo block of one k-step (32 B per row and term)
// operand format of the contraction: 3xTF32 (13 k-steps of 8 floats) or split-fp16 (7 k-steps of 16 halves; hi = fp16(x),
// lo = fp16(x - hi) UNscaled: centred embeddings are O(1), so lo keeps >= 2^-24 absolute accuracy and all three products
// share one accumulator).  Same bytes per k-step and term, so both formats use the same pipeline; fp16 needs 21 instead
// of 39 tensor instructions per 256 bank rows and streams 54 % of the bank bytes from L2.
template <bool F16> struct MatchFmt {
    static constexpr int KS = F16 ? 7 : TC_KS;
    static constexpr uint32_t A_BYTES = KS * 2 * QB * KSTEP * 4;
    static constexpr uint32_t SMEM = A_BYTES + NST * B_STAGE + 256;
};
constexpr uint32_t A_BYTES = MatchFmt<false>::A_BYTES;       // 106496 (self-test path)
constexpr uint32_t SMEM_MATCH = MatchFmt<false>::SMEM;

// x [R][ld] fp32 (first K_valid columns used, zero beyond; rows >= R zero) -> tc image with row blocks of RB rows,
// K = ksteps*8 columns.  One thread per (row, float4 granule).
//
// center (optional, [>= K_valid]): the image holds x - center (the matching distances are invariant to a common
// translation of queries and bank rows; centred operands keep the truncating TMEM accumulation near zero, see
// aoc_global_match_tc).  valid_r2 (optional, [R]): rows whose entry is +inf are padding and stay zero.
__global__ void pack_tc_image_kernel(const float* __restrict__ x, int R, int K_valid, int ld, int RB, int ksteps,
                                     long long rows_padded, uint8_t* __restrict__ out,
                                     const float* __restrict__ center, const float* __restrict__ valid_r2) {
    long long total = rows_padded * ksteps * 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i % rows_padded;            // consecutive threads -> consecutive rows (coalesced 16 B stores)
        int g = (int)(i / rows_padded);           // granule index: k = 4*g
        int ks = g >> 1, half = g & 1;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < R && !(valid_r2 && isinf(__ldg(valid_r2 + r)))) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int k = g * 4 + e;
                if (k < K_valid) v[e] = __ldg(x + (size_t)r * ld + k) - (center ? __ldg(center + k) : 0.f);
            }
        }
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32(v[e], hi[e], lo[e]);
        long long rb = r / RB;
        int rr = (int)(r - rb * RB);
        size_t base = ((size_t)rb * ksteps + ks) * chunk_bytes(RB) + elem_offset(rr, half * 4);
        *reinterpret_cast<float4*>(out + base) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(out + base + block_bytes(RB)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// split-fp16 image of x - center: one thread per (row, granule of 8 halves); chunk (rb, ks) = [hi block][lo block]
__global__ void pack_tc_image_f16_kernel(const float* __restrict__ x, int R, int K_valid, int ld, int RB, int ksteps,
                                         long long rows_padded, uint8_t* __restrict__ out,
                                         const float* __restrict__ center, const float* __restrict__ valid_r2) {
    long long total = rows_padded * ksteps * 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i % rows_padded;
        int g = (int)(i / rows_padded);           // granule index: k = 8*g
        int ks = g >> 1, half = g & 1;
        uint32_t hi[4], lo[4];
        const bool live = r < R && !(valid_r2 && isinf(__ldg(valid_r2 + r)));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v[2] = {0.f, 0.f};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int k = g * 8 + e * 2 + h;
                if (live && k < K_valid) v[h] = __ldg(x + (size_t)r * ld + k) - (center ? __ldg(center + k) : 0.f);
                v[h] = fminf(fmaxf(v[h], -65504.f), 65504.f);
            }
            const __half2 hh = __floats2half2_rn(v[0], v[1]);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(v[0] - hf.x, v[1] - hf.y);
            hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
            lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        long long rb = r / RB;
        int rr = (int)(r - rb * RB);
        size_t base = ((size_t)rb * ksteps + ks) * chunk_bytes(RB) + elem_offset16(rr, half * 8);
        *reinterpret_cast<uint4*>(out + base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(out + base + block_bytes(RB)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// Bank-sharded matching of ONE sequence on several GPUs: rank r contracts the queries with its range of
// bank row blocks only; min over bank rows = min over the ranks' partial minima (associative and commutative: the
// result is bit-identical to the single-GPU one).  The exchange is part of the matching kernel: as soon as a CTA has
// the per-object minima of its 128 query rows it stores them into its slot of every peer's receive buffer (plain
// stores to peer-mapped memory, i.e. NVLink writes), fences at system scope and bumps the peer's arrival counter --
// the transfer of a tile overlaps the contraction of the tiles still running; no collective call, no extra launch.
constexpr int MATCH_MAX_PEERS = 8;
struct MatchPeers {
    float* recv[MATCH_MAX_PEERS];        // peer p's receive slot for THIS rank, parity 0: [HW][O]
    unsigned* flag[MATCH_MAX_PEERS];     // peer p's arrival counter, parity 0 (parity 1: + 16 words)
    int n;                               // number of peers (world - 1); 0 = not sharded
    const unsigned* epoch;               // device word: frames exchanged so far; parity = epoch & 1 (double-buffered slots:
                                         // a rank can run at most one frame ahead of a peer, see DESIGN)
    long long par_stride;                // floats between the parity-0 and parity-1 slots (same layout on every rank)
    int fast;                            // 1: "fast" precision mode -- only the hi*hi term of the split product (see g_match_fast)
    int collector;                       // 1: hi(A) is kept in the tensor core's collector between the two products it is in
};

// MODE 0: matching epilogue (running min per object -> mins[split][HW][O]);  MODE 1: raw C = A*B^T (self-test);
// MODE 2: MODE 0 restricted to the row blocks [rb_begin, rb_end) (grid.y = 1) + the peer exchange above.
// Simg / r2 start at row block img_rb0.
template <int MODE, bool F16>
__global__ void __launch_bounds__(256, 1) match_tc_kernel(const uint8_t* __restrict__ Qimg,
                                                          const uint8_t* __restrict__ Simg,
                                                          const float* __restrict__ q2, const float* __restrict__ r2,
                                                          const int* __restrict__ meta, int O, int HW, int rb_begin,
                                                          int rb_end, int img_rb0,
                                                          int nsplit, float* __restrict__ mins, float* __restrict__ C,
                                                          int ldc, uint32_t lbo, uint32_t sbo, MatchPeers peers) {
    constexpr int TC_KS = MatchFmt<F16>::KS;
    constexpr uint32_t A_BYTES = MatchFmt<F16>::A_BYTES;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + NST * B_STAGE);
    // bars: [0..NST) full, [NST..2NST) empty, 2NST = a_full, 2NST+1.. tmem_full[2], 2NST+3.. tmem_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 5);
    __shared__ int seg[MAXO_ + 1];

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int qt = blockIdx.x, sp = blockIdx.y;
    long long par_off = 0;
    if (MODE == 2) {                                       // this frame's parity: read at run time, a captured graph follows it
        par_off = (long long)(*peers.epoch & 1u) * peers.par_stride;
        mins += par_off;
    }
    const int rb0 = rb_begin + (int)((long long)(rb_end - rb_begin) * sp / nsplit);
    const int rb1 = rb_begin + (int)((long long)(rb_end - rb_begin) * (sp + 1) / nsplit);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
    const uint32_t A_FULL = bar0 + 8u * (2 * NST);
    auto TFULL = [&](int s) { return bar0 + 8u * (2 * NST + 1 + s); };
    auto TEMPTY = [&](int s) { return bar0 + 8u * (2 * NST + 3 + s); };

    if (threadIdx.x <= O && MODE != 1) seg[threadIdx.x] = meta[MAXO_ + threadIdx.x];
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(A_FULL, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(TFULL(s), 1); mbar_init(TEMPTY(s), 128); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0 && rb1 > rb0) {
            // ===== TMA producer =====
            mbar_arrive_expect_tx(A_FULL, A_BYTES);
            for (int ks = 0; ks < TC_KS; ++ks)
                bulk_g2s(smem_u32(sA) + ks * (A_BYTES / TC_KS), Qimg + ((size_t)qt * TC_KS + ks) * (A_BYTES / TC_KS),
                         A_BYTES / TC_KS, A_FULL);
            int stage = 0;
            uint32_t phase = 0;
            for (int rb = rb0; rb < rb1; ++rb)
                for (int ks = 0; ks < TC_KS; ++ks) {
                    mbar_wait(EMPTY(stage), phase ^ 1u);
                    mbar_arrive_expect_tx(FULL(stage), B_STAGE);
                    bulk_g2s(smem_u32(sB) + stage * B_STAGE, Simg + ((size_t)(rb - img_rb0) * TC_KS + ks) * B_STAGE,
                             B_STAGE, FULL(stage));
                    if (++stage == NST) { stage = 0; phase ^= 1u; }
                }
        }
    } else if (warp == 1) {
        if (rb1 > rb0) {
            // ===== MMA issuer: warp-uniform loop, one elected lane issues (operands stay in uniform registers) =====
            const uint32_t idesc = F16 ? idesc_f16(QB, RBK) : idesc_tf32(QB, RBK);
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            mbar_wait(A_FULL, 0);
            tc_fence_after();
            int stage = 0, as = 0;
            uint32_t phase = 0, aph0 = 0, aph1 = 0;
            for (int rb = rb0; rb < rb1; ++rb) {
                if (as == 0) { mbar_wait(TEMPTY(0), aph0 ^ 1u); aph0 ^= 1u; }
                else         { mbar_wait(TEMPTY(1), aph1 ^ 1u); aph1 ^= 1u; }
                tc_fence_after();
                const uint32_t d = tb + (uint32_t)as * RBK;
                for (int ks = 0; ks < TC_KS; ++ks) {
                    mbar_wait(FULL(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_hi = smem_u32(sA) + ks * (A_BYTES / TC_KS);
                        const uint32_t a_lo = a_hi + QB * KSTEP * 4;
                        const uint32_t b_hi = smem_u32(sB) + stage * B_STAGE;
                        const uint32_t b_lo = b_hi + RBK * KSTEP * 4;
                        const uint64_t dah = smem_desc(a_hi, lbo, sbo), dal = smem_desc(a_lo, lbo, sbo);
                        const uint64_t dbh = smem_desc(b_hi, lbo, sbo), dbl = smem_desc(b_lo, lbo, sbo);
                        if (F16 && peers.fast) {
                            mma_f16(d, dah, dbh, idesc, ks > 0 ? 1u : 0u);         // fast mode: fp16-rounded operands, one MMA
                        } else if (F16 && peers.collector) {
                            mma_f16(d, dal, dbh, idesc, ks > 0 ? 1u : 0u);
                            mma_f16_afill(d, dah, dbl, idesc, 1u);             // hi(A) fetched once for the two products it is in
                            mma_f16_alast(d, dah, dbh, idesc, 1u);
                        } else if (F16) {
                            mma_f16(d, dal, dbh, idesc, ks > 0 ? 1u : 0u);
                            mma_f16(d, dah, dbl, idesc, 1u);
                            mma_f16(d, dah, dbh, idesc, 1u);
                        } else {
                            mma_tf32(d, dal, dbh, idesc, ks > 0 ? 1u : 0u);
                            mma_tf32(d, dah, dbl, idesc, 1u);
                            mma_tf32(d, dah, dbh, idesc, 1u);
                        }
                        mma_commit(EMPTY(stage));
                        if (ks == TC_KS - 1) mma_commit(TFULL(as));
                    }
                    __syncwarp();
                    if (++stage == NST) { stage = 0; phase ^= 1u; }
                }
                as ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 128 threads <-> 128 TMEM lanes (query rows) =====
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const int qi = qt * QB + row;
        const float qq = (MODE != 1 && qi < HW) ? __ldg(q2 + qi) : 0.f;
        int as = 0;
        uint32_t aphase = 0;
        int cur = 0;
        float m0 = INFINITY, m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;   // four independent chains of the running minimum
        if (MODE != 1) {
            while (cur < O - 1 && rb0 * RBK >= seg[cur + 1]) ++cur;
        }
        for (int rb = rb0; rb < rb1; ++rb) {
            if (MODE != 1) {
                if (rb * RBK >= seg[cur + 1]) {   // crossed into the next object's segment: flush
                    if (qi < HW) mins[((size_t)sp * HW + qi) * O + cur] = fminf(fminf(m0, m1), fminf(m2, m3));
                    m0 = m1 = m2 = m3 = INFINITY;
                    while (cur < O - 1 && rb * RBK >= seg[cur + 1]) ++cur;
                }
            }
            mbar_wait(TFULL(as), aphase);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)as * RBK;
#pragma unroll 1
            for (int c0 = 0; c0 < RBK; c0 += 64) {
                float v[64];                                      // two tensor-memory loads in flight per wait
                tmem_ld32(t0 + c0, v);
                tmem_ld32(t0 + c0 + 32, v + 32);
                tmem_ld_wait();
                if (MODE != 1) {
                    const float4* rr = reinterpret_cast<const float4*>(r2 + (size_t)(rb - img_rb0) * RBK + c0);
#pragma unroll
                    for (int j4 = 0; j4 < 16; ++j4) {
                        float4 r4 = __ldg(rr + j4);
                        m0 = fminf(m0, fmaf(-2.0f, v[j4 * 4 + 0], qq + r4.x));   // (|q|^2+|r|^2) - 2 q.r  (matching.py:45)
                        m1 = fminf(m1, fmaf(-2.0f, v[j4 * 4 + 1], qq + r4.y));
                        m2 = fminf(m2, fmaf(-2.0f, v[j4 * 4 + 2], qq + r4.z));
                        m3 = fminf(m3, fmaf(-2.0f, v[j4 * 4 + 3], qq + r4.w));
                    }
                } else {
                    float* dst = C + (size_t)qi * ldc + (size_t)rb * RBK + c0;
#pragma unroll
                    for (int j = 0; j < 64; ++j) dst[j] = v[j];
                }
            }
            tc_fence_before();
            mbar_arrive(TEMPTY(as));
            as ^= 1;
            if (as == 0) aphase ^= 1u;
        }
        if (MODE != 1 && rb1 > rb0 && qi < HW) mins[((size_t)sp * HW + qi) * O + cur] = fminf(fminf(m0, m1), fminf(m2, m3));
        if (MODE == 2) {
            // ---- peer exchange: this CTA's [rows][O] block of the local slot (objects outside this rank's row range keep
            // the +inf the slot was filled with) goes to the same place in every peer's slot for this rank
            asm volatile("bar.sync 3, 128;" ::: "memory");                  // the block is complete (the four epilogue warps)
            const int te = threadIdx.x - 128;
            const int nval = min(QB, HW - qt * QB) * O;
            const size_t off = (size_t)qt * QB * O;
            for (int pi = 0; pi < peers.n; ++pi) {
                float* dst = peers.recv[pi] + par_off + off;
                for (int i = te; i < nval; i += 128) dst[i] = __ldcg(mins + off + i);
            }
            __threadfence_system();
            asm volatile("bar.sync 3, 128;" ::: "memory");
            if (te == 0)
                for (int pi = 0; pi < peers.n; ++pi) atomicAdd_system(peers.flag[pi] + (par_off ? 16 : 0), 1u);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// column means of x [rows][100]: part[b][100] partial sums over a fixed row slab (deterministic), then mu[k]
__global__ void __launch_bounds__(256) col_sum_partial_kernel(const float* __restrict__ x, int rows, int slab,
                                                               float* __restrict__ part) {
    __shared__ float sm[8][104];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    const int r0 = blockIdx.x * slab, r1 = min(rows, r0 + slab);
    if (lane < 25)
        for (int r = r0 + warp; r < r1; r += 8) {
            float4 v = ldg4(x + (size_t)r * 100 + lane * 4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
    if (lane < 25) *reinterpret_cast<float4*>(&sm[warp][lane * 4]) = a;
    __syncthreads();
    if (threadIdx.x < 100) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];
        part[(size_t)blockIdx.x * 100 + threadIdx.x] = t;
    }
}
__global__ void col_mean_final_kernel(const float* __restrict__ part, int nb, int rows, float* __restrict__ mu) {
    int k = threadIdx.x;
    if (k >= 104) return;
    double t = 0.0;
    if (k < 100)
        for (int b = 0; b < nb; ++b) t += (double)part[(size_t)b * 100 + k];
    mu[k] = (float)(t / (double)rows);
}
// out[row] = |x[row] - center|^2 (warp per row); rows flagged +inf in valid_r2 (padding) keep +inf
__global__ void centered_sqnorm_kernel(const float* __restrict__ x, int rows, const float* __restrict__ center,
                                       const float* __restrict__ valid_r2, float* __restrict__ out) {
    int lane = threadIdx.x & 31;
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    float s = 0.f;
    if (lane < 25) {
        float4 v = ldg4(x + (size_t)row * 100 + lane * 4);
        float4 c = ldg4(center + lane * 4);
        v.x -= c.x; v.y -= c.y; v.z -= c.z; v.w -= c.w;
        s = v.x * v.x; s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) out[row] = (valid_r2 && isinf(valid_r2[row])) ? INFINITY : s;
}

__global__ void fill_f32_kernel(float* p, float v, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}

// mins[split][HW][O] -> min over splits, in place into split 0
__global__ void min_over_splits_kernel(float* __restrict__ mins, long long n, int nsplit) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m = mins[i];
    for (int s = 1; s < nsplit; ++s) m = fminf(m, mins[(size_t)s * n + i]);
    mins[i] = m;
}

static int pick_splits(int nqt, int nrb) {
    int s = 1;
    while (nqt * s < 2 * 148 && s * 2 <= nrb && s < 16) s *= 2;
    return s;
}

static PerDeviceOnce g_attr0, g_attr1, g_attr2;

}  // namespace aoc

using namespace aoc;

extern "C" size_t aoc_tc_image_bytes(long long rows, int K_img, int RB) {
    long long rb = (rows + RB - 1) / RB;
    int ks = (K_img + KSTEP - 1) / KSTEP;
    return (size_t)rb * ks * chunk_bytes(RB);
}

// x [rows][ld] (first K columns valid) -> tc image with K_img >= K columns (zero padded), row blocks of RB rows
extern "C" int aoc_pack_tc_image_f32(const float* x, long long rows, int K, int ld, int RB, int K_img, void* out,
                                     cudaStream_t stream) {
    AOC_CHECK_ARG(x && out && rows > 0 && K > 0 && K_img >= K, "bad args");
    AOC_CHECK_ARG(RB % 8 == 0 && RB >= 8, "RB must be a multiple of 8");
    int ks = (K_img + KSTEP - 1) / KSTEP;
    long long rows_padded = (rows + RB - 1) / RB * RB;
    long long total = rows_padded * ks * 2;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    pack_tc_image_kernel<<<blocks, 256, 0, stream>>>(x, (int)rows, K, ld, RB, ks, rows_padded, (uint8_t*)out, nullptr,
                                                     nullptr);
    return launch_status("aoc_pack_tc_image_f32");
}

namespace aoc {
int g_match_f16 = 1;     // aoc_set_option("match_f16", 0/1): split-fp16 operands in the global matching contraction
// aoc_set_option("match_fast", 0/1): FAST precision mode of the global matching (off by default; not bit-faithful): the
// contraction keeps only the hi*hi term of the split-fp16 product, i.e. a plain fp16 tensor-core GEMM with fp32
// accumulation (11-bit operands, ~1e-3 absolute on the squared distances of the centred embeddings).  One third of the
// tensor work of the exact mode; reported by bench.py --fast-match as frames/s next to IoU / argmax agreement with the oracle.
int g_match_fast = 0;
// aoc_set_option("match_collector", 0/1): collector hints on the A operand of the split product (same arithmetic, same
// order: hi(A) is fetched from shared memory once for the two products it is in -- UTCHMMA ... .A_REUSE in the SASS).
// Measured on B200 (bench.py, three runs on one box): 358 TFLOP/s fp32-equivalent without the hints, 316-317 with them --
// a kept collector serialises the two MMAs behind each other's operand fetch instead of saving one; off by default.
int g_match_collector = 0;
}

static int pack_centered(const float* x, long long rows, int RB, const float* center, const float* valid_r2, void* out,
                         bool f16, cudaStream_t stream) {
    int ks = f16 ? MatchFmt<true>::KS : MatchFmt<false>::KS;
    long long rows_padded = (rows + RB - 1) / RB * RB;
    long long total = rows_padded * ks * 2;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (f16)
        pack_tc_image_f16_kernel<<<blocks, 256, 0, stream>>>(x, (int)rows, 100, 100, RB, ks, rows_padded, (uint8_t*)out,
                                                             center, valid_r2);
    else
        pack_tc_image_kernel<<<blocks, 256, 0, stream>>>(x, (int)rows, 100, 100, RB, ks, rows_padded, (uint8_t*)out,
                                                         center, valid_r2);
    return launch_status("aoc_global_match_tc(pack)");
}

constexpr int MU_SLAB = 512;

extern "C" size_t aoc_global_match_tc_workspace_bytes(int HW, int rows_padded) {
    // query image (RB=128) + bank image (RB=256) + |q-mu|^2 + |r-mu|^2 + mu + column-sum partials + mins[16][HW][MAXO]
    size_t img = aoc_tc_image_bytes(HW, TC_K, QB) + aoc_tc_image_bytes(rows_padded > 0 ? rows_padded : 1, TC_K, RBK);
    size_t vec = (size_t)((HW + QB + 3) & ~3) + (size_t)(rows_padded + RBK) + 128 + (size_t)(cdiv(HW, MU_SLAB) + 1) * 100;
    return img + vec * sizeof(float) + (size_t)16 * HW * MAXO_ * sizeof(float) + 4096;
}

// q [HW][100]; S [rows_padded][100] = object-sorted bank rows (aoc_bank_gather_f32, align = 256) with r2 = |row|^2
// (+inf on padding rows); meta (device) as written by aoc_bank_index_build; rows_padded = meta[2*MAXO+1] (host copy).
// Queries and bank rows are translated by mu = column mean of q before the contraction: |q - r|^2 is unchanged, but
// the dot products q'.r' hover around zero instead of growing monotonically (embeddings are post-ReLU, all >= 0), which
// removes most of the truncation bias of the TMEM accumulation and most of the cancellation in |q|^2+|r|^2-2q.r.
extern "C" int aoc_global_match_tc(const float* q, int HW, const float* S, const float* r2, const int* meta_dev,
                                   int rows_padded, const float* bias, int O, void* workspace, size_t ws_bytes,
                                   float* out, cudaStream_t stream) {
    AOC_CHECK_ARG(q && S && r2 && meta_dev && bias && workspace && out, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= MAXO_ && HW > 0 && rows_padded % RBK == 0, "bad dims");
    AOC_CHECK_ARG(ws_bytes >= aoc_global_match_tc_workspace_bytes(HW, rows_padded), "workspace too small");
    uint8_t* ws = (uint8_t*)workspace;
    size_t imgq = aoc_tc_image_bytes(HW, TC_K, QB), imgs = aoc_tc_image_bytes(rows_padded > 0 ? rows_padded : 1, TC_K, RBK);
    uint8_t* Qimg = ws;
    uint8_t* Simg = ws + imgq;
    float* q2 = (float*)(ws + imgq + imgs);
    float* r2c = q2 + ((HW + QB + 3) & ~3);            // keeps r2c / mu 16-byte aligned (float4 loads)
    float* mu = r2c + (rows_padded + RBK);
    float* part = mu + 128;
    float* mins = part + (size_t)(cdiv(HW, MU_SLAB) + 1) * 100;
    int nqt = cdiv(HW, QB), nrb = rows_padded / RBK;
    int nb = cdiv(HW, MU_SLAB);
    col_sum_partial_kernel<<<nb, 256, 0, stream>>>(q, HW, MU_SLAB, part);
    col_mean_final_kernel<<<1, 128, 0, stream>>>(part, nb, HW, mu);
    const bool f16 = g_match_f16 != 0;      // (the fp16 images are smaller: the TF32-sized workspace layout is kept)
    int rc = pack_centered(q, HW, QB, mu, nullptr, Qimg, f16, stream);
    if (rc) return rc;
    centered_sqnorm_kernel<<<cdiv((long long)HW * 32, 256), 256, 0, stream>>>(q, HW, mu, nullptr, q2);
    if (nrb > 0) {
        rc = pack_centered(S, rows_padded, RBK, mu, r2, Simg, f16, stream);
        if (rc) return rc;
        centered_sqnorm_kernel<<<cdiv((long long)rows_padded * 32, 256), 256, 0, stream>>>(S, rows_padded, mu, r2, r2c);
    }
    int nsplit = nrb > 0 ? pick_splits(nqt, nrb) : 1;
    long long n = (long long)HW * O;
    fill_f32_kernel<<<cdiv(n * nsplit, 1024), 256, 0, stream>>>(mins, INFINITY, n * nsplit);
    if (nrb > 0) {
        if (g_attr0.first()) {
            cudaFuncSetAttribute(match_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MatchFmt<false>::SMEM);
            cudaFuncSetAttribute(match_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MatchFmt<true>::SMEM);
        }
        dim3 grid(nqt, nsplit);
        MatchPeers pr0 = {};
        pr0.fast = g_match_fast;
        pr0.collector = g_match_collector;
        if (f16)
            match_tc_kernel<0, true><<<grid, 256, MatchFmt<true>::SMEM, stream>>>(Qimg, Simg, q2, r2c, meta_dev, O, HW, 0, nrb, 0,
                                                                                 nsplit, mins, nullptr, 0, LBO_BYTES, SBO_BYTES,
                                                                                 pr0);
        else
            match_tc_kernel<0, false><<<grid, 256, MatchFmt<false>::SMEM, stream>>>(Qimg, Simg, q2, r2c, meta_dev, O, HW, 0, nrb, 0,
                                                                                   nsplit, mins, nullptr, 0, LBO_BYTES, SBO_BYTES,
                                                                                   pr0);
        if (nsplit > 1) min_over_splits_kernel<<<cdiv(n, 256), 256, 0, stream>>>(mins, n, nsplit);
    }
    rc = aoc_global_match_finalize_f32(mins, meta_dev, bias, HW, O, out, stream);
    if (rc) return rc;
    return launch_status("aoc_global_match_tc");
}

// ---------------------------------------------------------------------------------------------- bank-sharded matching
namespace aoc {
// Per-rank exchange area (one cudaMalloc'ed, IPC-exported allocation per rank, see peer.cu):
//   [0, 256)                 arrival counters: flag[parity] at byte 64 * parity (monotone, bumped by the peers' CTAs)
//   [256, ...)               recv[parity][world][cap_hw][MAXO] floats: slot (parity, g) = rank g's partial minima
// Local state (device, 64 bytes, zeroed once): {epoch, expected[2], done counter}.
__host__ __device__ inline size_t match_slot_floats(int cap_hw) { return (size_t)cap_hw * MAXO_; }

struct ShardState { unsigned epoch, expected[2], done; };

// waits for every peer's partial minima of this frame, reduces over the ranks, then the same tail as
// global_match_finalize_kernel (matching.py:83-90, :2505-2508).  The last block advances the epoch.
__global__ void __launch_bounds__(256) match_sharded_finalize_kernel(const float* __restrict__ recv_all /*[2][world][cap][MAXO]*/,
                                                                    const unsigned* __restrict__ flags, ShardState* st,
                                                                    int world, int cap_hw, unsigned add_expected,
                                                                    const int* __restrict__ meta,
                                                                    const float* __restrict__ bias, int HW, int O,
                                                                    float* __restrict__ out) {
    __shared__ unsigned s_par;
    if (threadIdx.x == 0) {
        const unsigned e = st->epoch;
        const unsigned par = e & 1u;
        const unsigned want = st->expected[par] + add_expected;
        const unsigned* f = flags + 16 * par;
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        } while ((int)(v - want) < 0);
        s_par = par;
    }
    __syncthreads();
    const unsigned par = s_par;
    const float* base = recv_all + (size_t)par * world * match_slot_floats(cap_hw);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < HW) {
        if (meta[2 * MAXO_ + 2] == 0) {
            for (int o = 0; o < O; ++o) out[(size_t)p * O + o] = 1.0f;
        } else {
            float m[MAXO_];
            float m1 = INFINITY, m2 = INFINITY;
            int a1 = -1;
            for (int o = 0; o < O; ++o) {
                float v = INFINITY;
                if (meta[o] > 0)
                    for (int g = 0; g < world; ++g)
                        v = fminf(v, __ldcg(base + (size_t)g * match_slot_floats(cap_hw) + (size_t)p * O + o));
                m[o] = v;
                if (v < m1) { m2 = m1; m1 = v; a1 = o; }
                else if (v < m2) { m2 = v; }
            }
            for (int o = 0; o < O; ++o) {
                const float other = (o == a1) ? m2 : m1;
                const float d = fminf(m[o], other + AOC_WRONG_LABEL_PAD);
                out[(size_t)p * O + o] = sig2(d + __ldg(bias + o));
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(&st->done, 1u);
        if (done == gridDim.x - 1) {
            st->expected[par] += add_expected;
            st->done = 0;
            __threadfence();
            st->epoch += 1;
        }
    }
}

}  // namespace aoc

extern "C" size_t aoc_match_shard_area_bytes(int world, int cap_hw) {
    return 256 + (size_t)2 * world * match_slot_floats(cap_hw) * sizeof(float);
}

// Row blocks [nrb * rank / world, nrb * (rank + 1) / world) of the object-sorted bank are this rank's share.
extern "C" int aoc_match_shard_range(int rows_padded, int rank, int world, int* rb_begin, int* rb_end) {
    AOC_CHECK_ARG(rows_padded % RBK == 0 && world >= 1 && rank >= 0 && rank < world && rb_begin && rb_end, "bad args");
    const int nrb = rows_padded / RBK;
    *rb_begin = (int)((long long)nrb * rank / world);
    *rb_end = (int)((long long)nrb * (rank + 1) / world);
    return AOC_OK;
}

namespace aoc {
__global__ void fill_slot_kernel(float* base, long long par_stride, const unsigned* epoch, float v, long long n) {
    float* p = base + (long long)(*epoch & 1u) * par_stride;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}
}  // namespace aoc

// aoc_global_match_tc for ONE sequence whose bank is sharded over `world` GPUs (one process per GPU, all running the same
// frames in lockstep).  areas[g] = rank g's exchange area (aoc_match_shard_area_bytes, allocated / exported / opened with
// the aoc_peer_* calls) as mapped into this process, areas[rank] the local one; state = 64 zeroed device bytes (local).
// The call contracts q with this rank's row-block range only, ships the partial minima to every peer from inside the
// matching kernel and reduces the `world` partial results; `out` is bit-identical to aoc_global_match_tc's on every rank.
extern "C" int aoc_global_match_tc_sharded(const float* q, int HW, const float* S, const float* r2, const int* meta_dev,
                                           int rows_padded, const float* bias, int O, int rank, int world,
                                           void* const* areas, int cap_hw, void* state, void* workspace,
                                           size_t ws_bytes, float* out, cudaStream_t stream) {
    AOC_CHECK_ARG(q && S && r2 && meta_dev && bias && workspace && out && areas && state, "null pointer");
    AOC_CHECK_ARG(O >= 1 && O <= MAXO_ && HW > 0 && HW <= cap_hw && rows_padded % RBK == 0, "bad dims");
    AOC_CHECK_ARG(world >= 2 && world <= MATCH_MAX_PEERS + 1 && rank >= 0 && rank < world, "bad rank / world");
    AOC_CHECK_ARG(ws_bytes >= aoc_global_match_tc_workspace_bytes(HW, rows_padded), "workspace too small");
    uint8_t* ws = (uint8_t*)workspace;
    size_t imgq = aoc_tc_image_bytes(HW, TC_K, QB), imgs = aoc_tc_image_bytes(rows_padded > 0 ? rows_padded : 1, TC_K, RBK);
    uint8_t* Qimg = ws;
    uint8_t* Simg = ws + imgq;
    float* q2 = (float*)(ws + imgq + imgs);
    float* r2c = q2 + ((HW + QB + 3) & ~3);
    float* mu = r2c + (rows_padded + RBK);
    float* part = mu + 128;
    const int nqt = cdiv(HW, QB), nb = cdiv(HW, MU_SLAB);
    int rb_lo = 0, rb_hi = 0;
    aoc_match_shard_range(rows_padded, rank, world, &rb_lo, &rb_hi);
    col_sum_partial_kernel<<<nb, 256, 0, stream>>>(q, HW, MU_SLAB, part);
    col_mean_final_kernel<<<1, 128, 0, stream>>>(part, nb, HW, mu);
    const bool f16 = g_match_f16 != 0;
    int rc = pack_centered(q, HW, QB, mu, nullptr, Qimg, f16, stream);
    if (rc) return rc;
    centered_sqnorm_kernel<<<cdiv((long long)HW * 32, 256), 256, 0, stream>>>(q, HW, mu, nullptr, q2);
    const long long my_rows = (long long)(rb_hi - rb_lo) * RBK;
    if (my_rows > 0) {                                   // only this rank's share of the bank is packed
        const float* S0 = S + (size_t)rb_lo * RBK * 100;
        rc = pack_centered(S0, my_rows, RBK, mu, r2 + (size_t)rb_lo * RBK, Simg, f16, stream);
        if (rc) return rc;
        centered_sqnorm_kernel<<<cdiv(my_rows * 32, 256), 256, 0, stream>>>(S0, (int)my_rows, mu, r2 + (size_t)rb_lo * RBK, r2c);
    }
    ShardState* st = (ShardState*)state;
    const size_t slot = match_slot_floats(cap_hw);
    MatchPeers pr = {};
    pr.n = 0;
    pr.fast = g_match_fast;
    pr.collector = g_match_collector;
    for (int g = 0; g < world; ++g) {
        if (g == rank) continue;
        AOC_CHECK_ARG(areas[g], "peer area not mapped");
        pr.recv[pr.n] = (float*)((char*)areas[g] + 256) + (size_t)rank * slot;       // my slot on peer g, parity 0
        pr.flag[pr.n] = (unsigned*)areas[g];
        ++pr.n;
    }
    pr.epoch = &st->epoch;
    pr.par_stride = (long long)world * slot;
    float* local_slot = (float*)((char*)areas[rank] + 256) + (size_t)rank * slot;     // parity 0
    const long long n = (long long)HW * O;
    fill_slot_kernel<<<cdiv(n, 1024), 256, 0, stream>>>(local_slot, pr.par_stride, &st->epoch, INFINITY, n);
    if (g_attr2.first()) {
        cudaFuncSetAttribute(match_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MatchFmt<false>::SMEM);
        cudaFuncSetAttribute(match_tc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MatchFmt<true>::SMEM);
    }
    dim3 grid(nqt, 1);
    if (f16)
        match_tc_kernel<2, true><<<grid, 256, MatchFmt<true>::SMEM, stream>>>(Qimg, Simg, q2, r2c, meta_dev, O, HW, rb_lo, rb_hi,
                                                                             rb_lo, 1, local_slot, nullptr, 0, LBO_BYTES,
                                                                             SBO_BYTES, pr);
    else
        match_tc_kernel<2, false><<<grid, 256, MatchFmt<false>::SMEM, stream>>>(Qimg, Simg, q2, r2c, meta_dev, O, HW, rb_lo,
                                                                               rb_hi, rb_lo, 1, local_slot, nullptr, 0,
                                                                               LBO_BYTES, SBO_BYTES, pr);
    match_sharded_finalize_kernel<<<cdiv(HW, 256), 256, 0, stream>>>(
        (const float*)((char*)areas[rank] + 256), (const unsigned*)areas[rank], st, world, cap_hw,
        (unsigned)((world - 1) * nqt), meta_dev, bias, HW, O, out);
    return launch_status("aoc_global_match_tc_sharded");
}

// Self-test of the tcgen05 pipeline: C[M][N] = A[M][K] * B[N][K]^T, K <= 104 (zero padded), M % 128 == 0, N % 256 == 0.
// variant 0: descriptors as designed; variant 1: LBO/SBO swapped (diagnostic only).  ws: packed images.
extern "C" int aoc_gemm_tf32x3_test(const float* A, const float* B, float* C, int M, int N, int K, int variant,
                                    void* workspace, size_t ws_bytes, cudaStream_t stream) {
    AOC_CHECK_ARG(A && B && C && workspace, "null pointer");
    AOC_CHECK_ARG(M % QB == 0 && N % RBK == 0 && K > 0 && K <= TC_K, "M%128, N%256, K<=104 required");
    size_t ia = aoc_tc_image_bytes(M, TC_K, QB), ib = aoc_tc_image_bytes(N, TC_K, RBK);
    AOC_CHECK_ARG(ws_bytes >= ia + ib, "workspace too small");
    uint8_t* Ai = (uint8_t*)workspace;
    uint8_t* Bi = Ai + ia;
    int rc = aoc_pack_tc_image_f32(A, M, K, K, QB, TC_K, Ai, stream);
    if (rc) return rc;
    rc = aoc_pack_tc_image_f32(B, N, K, K, RBK, TC_K, Bi, stream);
    if (rc) return rc;
    if (g_attr1.first()) {
        cudaFuncSetAttribute(match_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MATCH);
    }
    dim3 grid(M / QB, 1);
    uint32_t lbo = variant == 1 ? SBO_BYTES : LBO_BYTES, sbo = variant == 1 ? LBO_BYTES : SBO_BYTES;
    match_tc_kernel<1, false><<<grid, 256, SMEM_MATCH, stream>>>(Ai, Bi, nullptr, nullptr, nullptr, 1, M, 0, N / RBK, 0, 1, nullptr,
                                                         C, N, lbo, sbo, MatchPeers{});
    return launch_status("aoc_gemm_tf32x3_test");
}

// Implicit-GEMM convolution on the tcgen05 tensor cores, second generation (NHWC fp32 activations, fp32-faithful).
//   y[m, co] = sum_k A[m, k] * W[co, k],   m = (n, ho, wo) in a th x tw pixel patch,  k = (r, s, ci)
// Replaces the cuDNN nn.Conv2d call sites of the ResNet101-DeepLabv3+ backbone and the calibration decoder
// (resnet.py:23-42, deeplab/aspp.py:62-74, deeplab/decoder.py:32-41, layers/gct.py:68-91, layers/aspp.py:57-70,
// decoding_module.py:162-190,228-240), with the per-(sample, channel) affine that precedes them in the reference
// (GroupNorm apply + ReLU, GCT / IA gate) fused into the operand path.
//
// Data flow of one persistent CTA (work items = 128 output pixels x TN output channels [x a K slice]; K consumed in
// stages of 16 input channels of one filter tap; 20 warps, register file re-partitioned per role with setmaxnreg):
//   warp 16    TMA: cp.async.bulk.tensor.4d of the raw fp32 input patch [th][tw][32 channels] (zero fill = conv padding,
//              element strides = conv stride, 128B swizzle) into a 6-deep ring; one box feeds two operand stages
//   warps 0-7  transform, two sets of four owning alternate raw boxes = pairs of stages (thread = pixel = TMEM lane): raw ->
//              (optional a*x+b, ReLU, padding mask) -> operand split -> tcgen05.st into an operand ring in TENSOR memory.
//              Operand formats:
//              split-fp16 (default): hi = fp16(x), lo = fp16((x - hi) * 2^11), one K = 16 kind::f16 MMA per term;
//              3xTF32: hi = rna_tf32(x), lo = rna_tf32(x - hi), two K = 8 kind::tf32 MMAs per term
//   warp 17    TMA bulk copies of the pre-split weight image into the shared-memory operand ring
//   warps 18/19  warp-uniform loops, one elected lane issues tcgen05.mma with the A operand from TMEM:
//              hi*hi -> MAIN accumulator (warp 18), lo*hi + hi*lo -> CORR accumulator (warp 19); the operand ring is
//              synchronised per pair of stages (one full barrier for both operands, one empty barrier) and an issuer
//              iteration covers a pair
//   warps 8-15 drain: every `chunk` stages the MAIN accumulator (double buffered in TMEM) is read with tcgen05.ld and
//              added to fp32 registers with round-to-nearest; the epilogue adds CORR (x 2^-11 for split-fp16), transposes
//              through shared memory, applies bias, residual, ReLU, writes 128-byte lines and per-tile statistics.
// Shared memory only carries the raw patch and the weights: with both operands in shared memory the 128 B/clk port was
// the limit (measured 821 us vs 764 us on the largest layer), and the per-MMA operand set-up must come from uniform
// registers (a divergent single-thread issue loop cost ~14 instructions per MMA: 764 us -> 509 us once warp-uniform).
//
// Why the chunked accumulation: the tensor core adds into its fp32 accumulator with truncation, a systematic
// -0.5 ulp per MMA.  Over K = 2304 ... 18432 that bias reaches 1e-4 relative and broke the 1e-3 logit parity
// (measured: 7e-4 abs on a 2048->256 3x3 conv, 1.6e-2 on the final logits).  Short TMEM chains (K = 128) summed
// in registers, and the 2^-11-scaled correction terms kept in their own accumulator, bring the result back to
// fp32-FMA quality.
#include <cuda.h>

#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace aoc {
using namespace umma;

constexpr int C2_BM = 128;
constexpr int C2_KC = 16;                               // input channels per operand stage (2 k-steps of 8)
constexpr int C2_RKC = 32;                              // input channels per TMA box (128-byte rows)
constexpr uint32_t C2_RAW_BYTES = C2_BM * C2_RKC * 4;   // 16384
constexpr int C2_WRB = 128;                             // rows per block of the packed weight image
// bytes of one (row block, stage) weight chunk: 128 rows x 16 channels x {hi, lo}: 3xTF32 16384, split-fp16 8192
__host__ __device__ constexpr uint32_t c2_wchunk(bool f16) { return C2_WRB * C2_KC * (f16 ? 2u : 4u) * 2u; }
constexpr int C2_MAX_AFFINE_C = 1024;
constexpr int C2_MAX_KSPLIT = 8;                        // split-K slices per output tile
// setmaxnreg targets.  The CTA's registers are fixed at launch (20 warps x 96); an increase can only take what decreases
// of the same CTA have released, so 8 x 80 (transform) + 8 x 136 (drain) + 4 x 40 (TMA / MMA issue) <= 20 x 96.
#ifndef AOC_CONV_NS
#define AOC_CONV_NS 2
#endif
// transform sets (four warps each).  3 sets measured (tooling build -DAOC_CONV_NS=3): no gain on the large layers, which
// are then bound by L2 -> SM bandwidth (16 KB of weights + raw patch per stage and SM), see DESIGN section 4
constexpr int C2_NS = AOC_CONV_NS;
constexpr int C2_XW = 4 * C2_NS;                        // transform warps = first drain warp
constexpr int C2_XT = 32 * C2_XW;                       // transform threads = first drain thread
constexpr int C2_LAUNCH_REGS = C2_NS == 2 ? 96 : 80;    // what ptxas gives a thread under __launch_bounds__(32 * (C2_XW + 12), 1)
constexpr int C2_REGS_XFORM = C2_NS == 2 ? 80 : 64;
constexpr int C2_REGS_DRAIN = C2_NS == 2 ? 136 : 128;
constexpr int C2_REGS_MISC = C2_NS == 2 ? 40 : 32;
static_assert(C2_XW * C2_REGS_XFORM + 8 * C2_REGS_DRAIN + 4 * C2_REGS_MISC <= (C2_XW + 12) * C2_LAUNCH_REGS,
              "setmaxnreg.inc can only take what the CTA's own warps released: an over-subscribed budget blocks forever");

// Division of a non-negative int (< 2^31) by a launch constant as multiply-high + shift (the host computes the pair):
// the tile decode below is ten divisions by kernel parameters, inlined into the loop of every warp role -- as hardware-less
// integer divisions they were 1 157 of the 4 096 instructions of the split-K kernel, a quarter of a code image that is
// already twice the 32 KB instruction cache and is fetched cold at the start of every launch.
struct FastDiv { uint32_t mul, shr; };
inline FastDiv fastdiv_make(int d) {
    FastDiv f = {0u, 0u};
    if (d > 1) {
        uint32_t l = 0;
        while ((1u << l) < (uint32_t)d) ++l;                 // ceil(log2 d)
        const uint64_t pw = 1ull << (31 + l);
        f.mul = (uint32_t)((pw + (uint64_t)d - 1) / (uint64_t)d);
        f.shr = l - 1;
    }
    return f;
}
__device__ __forceinline__ int fastdiv(int x, FastDiv f) {       // f.mul == 0: divisor 1
    return f.mul ? (int)(__umulhi((uint32_t)x, f.mul) >> f.shr) : x;
}

struct Conv2P {
    const uint8_t* w; const float* bias; const float* res; const float* in_a; const float* in_b; float* y;
    float* tile_stats;      // optional [pixel tiles][2][Cout]: per-tile sum / sum of squares of the stored output
    int N, H, W, Cin, Ho, Wo, Cout, ldy, ldres, kw, stride, pad, dil, relu, in_relu;
    int tw_log2, th, tiles_x, tiles_y;
    int ncc, nIt, chunk, taps;
    int tiles_n, total_tiles;
    int ksplit;             // split-K factor: work item = (tile, K slice); slices write raw partial sums to `ws`
    int tail_mode, n_plain; // tail splitting (see decode): items below n_plain are whole tiles, the rest K slices of the tail tiles
    float* ws;              // [ksplit][N*Ho*Wo][Cout] partial sums (ksplit > 1)
    int* ws_cnt;            // [total_tiles] arrival counters of the K slices (zero between launches)
    unsigned long long* trace;   // tooling only (aoc_conv_trace): clock64 of pipeline events of CTA 0, [event < 16][stage < 256]
    int vec_out;
    int* overflow;          // optional device word: set to 1 when a split-fp16 activation operand reaches the fp16 range (|x| >= 6e4)
    int dbg;                // tooling build only (tools/conv_attrib.py): ablation bits, see C2_DBG
    FastDiv fd_ksplit, fd_nrc, fd_tiles_n, fd_tpi, fd_tiles_x;    // set by c2_set_fastdiv() once the launcher has fixed the schedule
};
static void c2_set_fastdiv(Conv2P& q) {
    q.fd_ksplit = fastdiv_make(q.ksplit);
    q.fd_nrc = fastdiv_make((q.ncc + 1) >> 1);
    q.fd_tiles_n = fastdiv_make(q.tiles_n);
    q.fd_tpi = fastdiv_make(q.tiles_x * q.tiles_y);
    q.fd_tiles_x = fastdiv_make(q.tiles_x);
}

#ifdef AOC_CONV_TRACE   // tooling build only (tools/conv_trace.py): keeps the production kernel's code small
// stage index of an event = running stage count of CTA 0 (tile ordinal x stages per tile + stage): short-K layers show several tiles
#define C2_TRACE(ev, st) do { if (p.trace && blockIdx.x == 0) { \
                                  const int st_ = __float2int_rd(((float)t + 0.5f) * c2_inv_grid) * nIt + (st); \
                                  if (st_ < 256) p.trace[(ev) * 256 + st_] = clock64(); } } while (0)
// ablation switches of the tooling build (aoc_set_option("conv_dbg", bits); results are garbage, the timing is the point):
// 1 no weight copies (the barrier is completed by a plain arrive), 2 no activation TMA, 4 no correction MMAs,
// 8 no transform arithmetic (zeros are stored), 16 no main MMAs
#define C2_DBG(bit) ((p.dbg & (bit)) != 0)
// whole-kernel milestones of CTA 0 (event row 15): 0 kernel entry, 1 prologue done, 2 previous grid complete (transform warp 0),
// 3 first activation TMA issued, 4 accumulators drained, 5 partial tile written + arrived, 6 all K slices arrived, 7 finished
#define C2_MARK(k) do { if (p.trace && blockIdx.x == 0) p.trace[15 * 256 + (k)] = clock64(); } while (0)
#else
#define C2_TRACE(ev, st) do { } while (0)
#define C2_DBG(bit) false
#define C2_MARK(k) do { } while (0)
#endif

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// Packed fp32 arithmetic (two IEEE round-to-nearest operations per issue slot; the transform warps are bound by issue
// slots, not by the FMA pipe): (x0, x1) <- (x0, x1) * (a0, a1) + (b0, b1) and (x0, x1) <- (x0, x1) * c.
__device__ __forceinline__ void ffma2(float& x0, float& x1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 x, a, b;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\t"
        "fma.rn.f32x2 x, x, a, b;\n\tmov.b64 {%0, %1}, x;\n\t}"
        : "+f"(x0), "+f"(x1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// (d0, d1) += (a0, a1) and (d0, d1) += (a0 * a0, a1 * a1): the per-channel statistics of the epilogue
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {
    asm("{\n\t.reg .b64 d, a;\n\tmov.b64 d, {%0, %1};\n\tmov.b64 a, {%2, %3};\n\t"
        "add.rn.f32x2 d, d, a;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void fsqacc2(float& d0, float& d1, float a0, float a1) {
    asm("{\n\t.reg .b64 d, a;\n\tmov.b64 d, {%0, %1};\n\tmov.b64 a, {%2, %3};\n\t"
        "fma.rn.f32x2 d, a, a, d;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void fmul2(float& x0, float& x1, float c) {
    asm("{\n\t.reg .b64 x, c;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 c, {%2, %2};\n\t"
        "mul.rn.f32x2 x, x, c;\n\tmov.b64 {%0, %1}, x;\n\t}"
        : "+f"(x0), "+f"(x1) : "f"(c));
}

// What bounds a stage (clock64 trace of the roles, tools/conv_trace.py; ablation runs, tools/conv_attrib.py): the
// iteration of the MMA-issuing warps, not the tensor pipe, the copies or the transform arithmetic (83 % of the time is
// left with all of those switched off).  First finding (3xTF32 days): ONE thread issuing 6 tcgen05.mma + 2-3
// tcgen05.commit per stage needed 400-550 cycles per stage, so the issue work is split over TWO warps -- one feeds the
// MAIN accumulator, one the CORR accumulator.  Second finding (split-fp16, 3 MMAs per stage = 192-222 tensor cycles):
// the MAIN issuer's iteration -- barrier waits (~100 cycles per try_wait, not overlappable), tcgen05 fence, elect, the
// uniform-register chain in front of each tcgen05 instruction, on a sub-partition shared with four busy warps -- took
// ~385 cycles per stage, hence one full barrier per PAIR of stages for both operands and two stages per iteration.
// Halo variant (3x3, stride 1, pad = dilation <= 2, split-fp16): output tiles of 16 rows x 8 columns; the raw halo patch
// (18 x 10 pixels x 32 channels; 20 x 12 for dilation 2) of a channel box is loaded ONCE (not once per filter tap), transformed ONCE into
// split-fp16 operand tiles in SHARED memory, and the nine taps of the box read that tile through nine descriptor start
// addresses: K-major SWIZZLE_NONE core matrices = 8 consecutive pixels of a halo row x 8 channels (16 B per pixel), the
// 8-row-group stride (SBO) is one halo row, the k-group stride (LBO) one pixel plane.
constexpr int C2H_TH = 16, C2H_TW = 8;                           // output tile (rows x columns) = 128 pixels
constexpr int C2H_D = 2;                                         // largest dilation (= padding) the shared-memory budget is sized for
constexpr int C2H_P = (C2H_TH + 2 * C2H_D) * (C2H_TW + 2 * C2H_D);      // halo pixels (180)
constexpr uint32_t C2H_RAW_BYTES = (C2H_P * 128 + 1023) / 1024 * 1024;  // raw halo box: P rows of 32 fp32, 1 KB aligned (swizzle)
constexpr int C2H_NRAW = 2;                                      // raw halo ring depth (a box lasts 18 stages: one in use, one in flight)
constexpr uint32_t C2H_A_HALF = 4 * C2H_P * 16;                  // hi (or lo) operand tile of a box: 4 k-groups x P pixels x 16 B
constexpr uint32_t C2H_A_BUF = 2 * C2H_A_HALF;                   // hi + lo

template <int TN, bool F16, bool HALO = false>
struct C2Cfg {
    // raw activation ring depth (128 pixels x 32 channels each); halo: barrier index space -- slots 0..2 the raw halo ring,
    // 3..4 the two operand-tile buffers
    static constexpr int NR = HALO ? 5 : (C2_NS == 2 ? 6 : 7);
    // operand ring (activation half in TMEM, 32 columns per stage; weight half in shared memory): a weight chunk is
    // requested when the stage it replaces retires and needs an L2 round trip (~1.5k cycles) to land, so the period of
    // a stage cannot drop below (round trip + MMA time) / depth: 8 stages where TMEM has room (TN = 64), else 4
    static constexpr int NO = (F16 || TN <= 64) ? 8 : 4;
    static constexpr uint32_t B_BYTES = TN * C2_KC * (F16 ? 2 : 4) * 2;
    static constexpr uint32_t A_COLS = F16 ? 16 : 32;        // TMEM columns of one activation operand stage
    static constexpr uint32_t RAW_OFF = 0;
    static constexpr uint32_t A_OFF = C2H_NRAW * C2H_RAW_BYTES;                // halo: operand tiles [2 buffers][hi | lo]
    static constexpr uint32_t OP_OFF = HALO ? A_OFF + 2 * C2H_A_BUF : NR * C2_RAW_BYTES;
    static constexpr uint32_t TAB_OFF = OP_OFF + NO * B_BYTES;
    static constexpr uint32_t STG_OFF = TAB_OFF + 2 * C2_MAX_AFFINE_C * 4;     // epilogue staging: 8 warps x 32 rows x 128 B
    static constexpr uint32_t BAR_OFF = STG_OFF + 8 * 4096;
    static constexpr uint32_t SMEM = BAR_OFF + 512 + 1024;    // + alignment slack
    static constexpr uint32_t TMEM_COLS = 512;
    static constexpr int NCB = (TN <= 64 || HALO) ? 2 : 1;   // CORR accumulators: double buffered across tiles when they fit
    // correction issuers: one warp for both terms.  (Round 1 attributed the slow issue loops to a per-instruction cost of
    // ~75 cycles per issuing thread and tried one warp per term; tools/microbench/umma_issue.cu shows otherwise: ONE thread
    // sustains 3 MMAs of M = N = 128 plus a commit every other stage in exactly the 192 tensor cycles -- an MMA costs its
    // thread <= 39 cycles, a commit 9.  What slows the real loops is waiting: on operands in the per-tap kernel, on the
    // shared-memory port in the halo variant.)
    static constexpr int NCI = 1;   // (a third issuer for TN = 64 was measured: no gain once the weight latency is the limit)
    static constexpr int THREADS = (C2_XW + 8 + 2 + 1 + NCI) * 32;
    static constexpr uint32_t A_TMEM_COL = (2 + NCB * NCI) * TN;   // ring of C2_NO stages x 32 columns behind the accumulators
};

// work item -> output tile and K range.  With split-K (SK; layers that cannot fill the chip with tiles) a work item is
// (tile, K slice): the slice covers the raw stages [r0, r1) of the (tap, 32-channel box) sequence, i.e. the operand
// stages [it0, it1).  Without it the K fields are the whole range and fold away (the issue loops of the plain kernel must
// not carry them: a Tile kept in local memory cost 18 % on the large layers).
struct Tile { int n, ho0, wo0, n0, ks, r0, r1, it0, it1, tap0, cc0, tile, nsl; };
#ifndef AOC_CONV_NO_HOIST     // (tooling build for the A/B measurement of the hoisted first-tile decode)
#define C2_FIRST_TILE(t) ((t) == (int)blockIdx.x)
#else
#define C2_FIRST_TILE(t) false
#endif   // tile: flat tile index; nsl: K slices of it

// SK: 0 = every work item is a whole tile, 1 = split-K of every tile (small maps), 2 = tail splitting (K slices for the tiles
// of the partial last wave only)
template <int TN, int SK, bool F16, bool HALO = false>
__global__ void __launch_bounds__(C2Cfg<TN, F16, HALO>::THREADS, 1) conv2_kernel(const __grid_constant__ CUtensorMap tmapA, Conv2P p) {
    static_assert(!HALO || (F16 && !SK && C2_NS == 2), "the halo variant is split-fp16, unsplit, two transform warpgroups");
    using Cfg = C2Cfg<TN, F16, HALO>;
    constexpr uint32_t C2_WCHUNK = c2_wchunk(F16);
    constexpr int C2_NR = Cfg::NR;
    constexpr int C2_NO = Cfg::NO, C2_NB = Cfg::NO;
    constexpr int NCB = Cfg::NCB;                            // CORR accumulator buffers (2 when TMEM allows)
    constexpr int NCI = Cfg::NCI;                            // correction issuers / accumulators per buffer
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* tab_a = reinterpret_cast<float*>(smem + Cfg::TAB_OFF);
    float* tab_b = tab_a + C2_MAX_AFFINE_C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 60);
    const uint32_t bar0 = smem_u32(bars);
    auto RAW_FULL = [&](int s) { return bar0 + 8u * s; };
    auto RAW_EMPTY = [&](int s) { return bar0 + 8u * (C2_NR + s); };
    // The operand ring is synchronised per PAIR of stages (slots 2q, 2q+1) and with ONE "full" barrier for both operands:
    // a try_wait occupies its warp for ~100 cycles even on a completed barrier and try_waits do not overlap (measured:
    // four back-to-back try_waits on completed barriers = 460 cycles), and the MMA-issuing warps -- the pace setters of
    // the kernel: ~385 cycles per stage with or without MMAs, copies or arithmetic (tools/conv_attrib.py) -- spent two
    // of them per stage.  OP_FULL(q) collects the four transform warps' arrivals and the weight copy's
    // arrive.expect_tx of both stages (count 10); OP_EMPTY(q) the issuers' commits after the pair's second stage.
    auto OP_FULL = [&](int q) { return bar0 + 8u * (2 * C2_NR + q); };
    auto OP_EMPTY = [&](int q) { return bar0 + 8u * (2 * C2_NR + C2_NO + q); };
    auto CORR_FULL = [&](int s) { return bar0 + 8u * (2 * C2_NR + 2 * C2_NO + C2_NB + s); };
    auto MAIN_FULL = [&](int b) { return bar0 + 8u * (2 * C2_NR + 2 * C2_NO + 2 * C2_NB + b); };
    auto MAIN_EMPTY = [&](int b) { return bar0 + 8u * (2 * C2_NR + 2 * C2_NO + 2 * C2_NB + 2 + b); };
    auto CORR_EMPTY = [&](int b) { return bar0 + 8u * (2 * C2_NR + 2 * C2_NO + 2 * C2_NB + 4 + b); };
    const uint32_t raw0 = smem_u32(smem + Cfg::RAW_OFF);
    const uint32_t op0 = smem_u32(smem + Cfg::OP_OFF);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
#ifdef AOC_CONV_TRACE
    const float c2_inv_grid = 1.0f / (float)gridDim.x;       // tile ordinal of CTA 0 without an integer division per event
#endif
    const int tw_mask = (1 << p.tw_log2) - 1;
    const int tpi = p.tiles_x * p.tiles_y;
    const bool affine = p.in_a != nullptr || p.in_b != nullptr || p.in_relu;
    const int nIt = p.nIt;
    // persistent CTA: tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; t -> (pixel tile, output-channel tile) with the
    // channel tile fastest, so CTAs running side by side share the activation patch in L2.  Every role walks the same
    // tile sequence and the rings keep flowing across tile boundaries (the producers run ahead into the next tile).
    const int nrc = (p.ncc + 1) >> 1;                        // raw (32-channel) stages per tap
    const int n_items = SK == 1 ? p.total_tiles * p.ksplit : SK == 2 ? p.n_plain + (p.total_tiles - p.n_plain) * p.ksplit : p.total_tiles;
    auto decode = [&](int w) {
        Tile tl;
        int t = w;
        if (SK) {
            int nsl = p.ksplit;
            if (SK == 2) {
                // tail splitting: the whole waves of tiles run unsplit, the tiles of the last, partial wave are cut into
                // `ksplit` K slices each so that the wave fills the chip (items n_plain .. : tile-major, slice fastest)
                if (w < p.n_plain) { t = w; tl.ks = 0; nsl = 1; }
                else { const int u = w - p.n_plain, uq = fastdiv(u, p.fd_ksplit); t = p.n_plain + uq; tl.ks = u - uq * p.ksplit; }
            } else {
                t = fastdiv(w, p.fd_ksplit);
                tl.ks = w - t * p.ksplit;
            }
            tl.nsl = nsl;
            const int R = p.taps * nrc;
            tl.r0 = nsl == 1 ? 0 : fastdiv(R * tl.ks, p.fd_ksplit);
            tl.r1 = nsl == 1 ? R : fastdiv(R * (tl.ks + 1), p.fd_ksplit);
            int tp = fastdiv(tl.r0, p.fd_nrc);
            tl.tap0 = tp;
            tl.cc0 = 2 * (tl.r0 - tp * nrc);
            tl.it0 = tp * p.ncc + tl.cc0;
            tp = fastdiv(tl.r1, p.fd_nrc);
            tl.it1 = tp * p.ncc + min(2 * (tl.r1 - tp * nrc), p.ncc);
        } else {
            tl.ks = 0; tl.r0 = 0; tl.r1 = p.taps * nrc; tl.tap0 = 0; tl.cc0 = 0; tl.it0 = 0; tl.it1 = nIt; tl.nsl = 1;
        }
        tl.tile = t;
        const int mt = fastdiv(t, p.fd_tiles_n);
        tl.n0 = (t - mt * p.tiles_n) * TN;
        tl.n = fastdiv(mt, p.fd_tpi);
        const int trem = mt - tl.n * tpi;
        const int tyi = fastdiv(trem, p.fd_tiles_x), txi = trem - tyi * p.tiles_x;
        tl.ho0 = tyi * p.th;
        tl.wo0 = txi << p.tw_log2;
        return tl;
    };

    // Programmatic dependent launch: the next kernel of the stream may be scheduled onto SMs as this grid's CTAs retire
    // and run its prologue (barrier init, TMEM allocation, descriptor fetch) under our tail; every role that touches
    // global memory first executes griddepcontrol.wait (= the previous grid has completed and its writes are visible).
    if (threadIdx.x == 0) C2_MARK(0);
    asm volatile("griddepcontrol.launch_dependents;");
    // The first work item of this CTA is decoded HERE, before the wait for the previous grid: ten integer divisions on
    // kernel parameters that miss the constant cache on first touch were ~1 500-3 000 cycles between "previous grid complete"
    // and the first activation TMA (tools/conv_marks.py), paid again by every role at the top of its loop -- and a layer of
    // the 31 x 54 backbone maps has exactly one work item per CTA.  (The empty asm pins the values: without it the
    // compiler re-materialises the decode inside each role, after the wait.)
    Tile tile0 = decode((int)blockIdx.x);
    asm volatile("" : "+r"(tile0.n), "+r"(tile0.ho0), "+r"(tile0.wo0), "+r"(tile0.n0), "+r"(tile0.ks), "+r"(tile0.r0),
                 "+r"(tile0.r1), "+r"(tile0.it0), "+r"(tile0.it1), "+r"(tile0.tap0), "+r"(tile0.cc0), "+r"(tile0.tile),
                 "+r"(tile0.nsl));
    if (warp == C2_XW + 9 && lane == 0) {
        if (HALO) {
            // RAW_FULL(0..2) / RAW_EMPTY(0..2): raw halo ring (released by the eight transform warps); RAW_FULL(3 + b) /
            // RAW_EMPTY(3 + b): operand-tile buffer b full (eight transform warps) / empty (both issuers' commits)
            for (int s = 0; s < C2H_NRAW; ++s) { mbar_init(RAW_FULL(s), 1); mbar_init(RAW_EMPTY(s), 8); }
            for (int b = 0; b < 2; ++b) { mbar_init(RAW_FULL(3 + b), 8); mbar_init(RAW_EMPTY(3 + b), 1 + NCI); }
            for (int q = 0; q < C2_NO / 2; ++q) { mbar_init(OP_FULL(q), 2); mbar_init(OP_EMPTY(q), 1 + NCI); }   // weights only
        } else {
            for (int s = 0; s < C2_NR; ++s) { mbar_init(RAW_FULL(s), 1); mbar_init(RAW_EMPTY(s), 4); }
            for (int q = 0; q < C2_NO / 2; ++q) { mbar_init(OP_FULL(q), 10); mbar_init(OP_EMPTY(q), 1 + NCI); }
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(MAIN_FULL(b), 1); mbar_init(MAIN_EMPTY(b), 8); mbar_init(CORR_EMPTY(b), 8); mbar_init(CORR_FULL(b), NCI);
        }
        fence_barrier_init();
    }
    if (warp == C2_XW + 10) {
        tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    // the activation tensor map is a kernel parameter: its descriptor fetch (the first cp.async.bulk.tensor would pay it
    // after the wait below) is started here, under the previous grid's tail
    if (warp == C2_XW + 8 && lane == 0)
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) C2_MARK(1);
    // transform, drain and the activation TMA producer wait for the previous grid; the weight producer (warp C2_XW + 9) does
    // NOT: the packed weight image is a constant of the model (complete before the first launch that uses it -- contract of
    // aoc_conv2d_nhwc_tc), so its first ring of stages is fetched from HBM while the previous layer's last CTAs finish
    // Nobody waits for the previous grid HERE.  Each role executes griddepcontrol.wait immediately in front of its first
    // access to global memory, so that its way there -- role dispatch, register re-partitioning, loop set-up: code that is
    // cold in the instruction cache at every launch (the kernel image is twice the 32 KB L1.5; measured at ~2 400 cycles
    // between "previous grid complete" and the first activation copy, tools/conv_marks.py) -- overlaps the previous grid:
    //   activation producer   in front of its first cp.async.bulk.tensor
    //   transform warps       in front of the coefficient table loads (in_a / in_b); without an input affine they touch shared
    //                         and tensor memory only and do not wait at all (the sticky overflow word is write-only)
    //   drain warps           after their register increase (residual / split-K partial reads, every global write)
    //   weight producer       never (constants), MMA issuers never (no global memory)

    if (warp < C2_XW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C2_REGS_XFORM));
        bool waited = false;                     // griddepcontrol.wait executed by this thread
        if constexpr (HALO) {
            // ===== halo transform: all eight warps convert the raw halo patch of a channel box (P pixels x 32 channels) ONCE
            // into the split-fp16 operand tiles [k-group of 8 channels][pixel][16 B] (hi and lo); item = (pixel, k-group) =====
            const int d = p.dil;
            const int HWd = C2H_TW + 2 * d, P = HWd * (C2H_TH + 2 * d);
            const uint32_t a0 = smem_u32(smem + Cfg::A_OFF);
            const uint32_t tab_s = smem_u32(tab_a);
            const bool need_mask = p.in_b != nullptr;
            int sr = 0, ab = 0;
            uint32_t pr = 0, pa0 = 0, pa1 = 0;
            int tab_n = -1;
            float amax = 0.f;
            for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
                const Tile tl = C2_FIRST_TILE(t) ? tile0 : decode(t);
                if (affine && tl.n != tab_n) {
                    if (!waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); waited = true; }
                    asm volatile("bar.sync 1, %0;" ::"n"(C2_XT) : "memory");
                    const int cpad = p.ncc * C2_KC;
                    for (int c = threadIdx.x; c < cpad; c += C2_XT) {
                        const bool ok = c < p.Cin;
                        tab_a[c] = ok ? (p.in_a ? __ldg(p.in_a + (size_t)tl.n * p.Cin + c) : 1.f) : 0.f;
                        tab_b[c] = (ok && p.in_b) ? __ldg(p.in_b + (size_t)tl.n * p.Cin + c) : 0.f;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(C2_XT) : "memory");
                    tab_n = tl.n;
                }
                const int hy0 = tl.ho0 - d, hx0 = tl.wo0 - d;             // image position of halo pixel (0, 0): stride 1, pad = d
                for (int cb = 0; cb < nrc; ++cb) {
                    const int nq = (2 * cb + 1 < p.ncc) ? 4 : 2;          // k-groups of this box (an odd stage count: half a box)
                    mbar_wait(RAW_FULL(sr), pr);
                    if (ab == 0) { mbar_wait(RAW_EMPTY(3), pa0 ^ 1u); pa0 ^= 1u; }
                    else         { mbar_wait(RAW_EMPTY(4), pa1 ^ 1u); pa1 ^= 1u; }
                    const uint32_t rawb = raw0 + sr * C2H_RAW_BYTES;
                    const uint32_t ahi = a0 + ab * C2H_A_BUF, alo = ahi + C2H_A_HALF;
                    for (int i = threadIdx.x; i < nq * P; i += C2_XT) {
                        const int q = (i >= P) + (i >= 2 * P) + (i >= 3 * P);
                        const int pp = i - q * P;
                        float v[8];
                        const uint32_t row = rawb + (uint32_t)(pp * 128), sw = (uint32_t)(pp & 7);
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(v[4 * j]), "=f"(v[4 * j + 1]), "=f"(v[4 * j + 2]), "=f"(v[4 * j + 3])
                                         : "r"(row + ((((uint32_t)(2 * q + j)) ^ sw) << 4)));
                        if (affine) {
                            const uint32_t tc = tab_s + (uint32_t)((cb * C2_RKC + q * 8) * 4);
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                float4 a4, b4;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                             : "=f"(a4.x), "=f"(a4.y), "=f"(a4.z), "=f"(a4.w) : "r"(tc + j * 16));
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                             : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(tc + C2_MAX_AFFINE_C * 4 + j * 16));
                                ffma2(v[4 * j], v[4 * j + 1], a4.x, a4.y, b4.x, b4.y);
                                ffma2(v[4 * j + 2], v[4 * j + 3], a4.z, a4.w, b4.z, b4.w);
                            }
                            if (p.in_relu) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
                            }
                            if (need_mask) {                                 // zero padding applies AFTER the affine
                                const int py = pp / HWd, px = pp - py * HWd;
                                const bool ok = (unsigned)(hy0 + py) < (unsigned)p.H && (unsigned)(hx0 + px) < (unsigned)p.W;
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] = ok ? v[e] : 0.f;
                            }
                        }
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float d0, d1;
                            asm("max.abs.f32 %0, %0, %1, %2;" : "+f"(amax) : "f"(v[2 * j]), "f"(v[2 * j + 1]));
                            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
                            asm("{\n\t.reg .b16 e0, e1;\n\tmov.b32 {e0, e1}, %2;\n\t"
                                "sub.rn.f32.f16 %0, e0, %3;\n\tsub.rn.f32.f16 %1, e1, %4;\n\t}"
                                : "=f"(d0), "=f"(d1) : "r"(h[j]), "f"(v[2 * j]), "f"(v[2 * j + 1]));
                            fmul2(d0, d1, -2048.f);
                            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l[j]) : "f"(d1), "f"(d0));
                        }
                        const uint32_t off = (uint32_t)((q * P + pp) * 16);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ahi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(alo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
                    }
                    fence_proxy_async();                                     // generic-proxy stores -> tcgen05.mma operand reads
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(RAW_FULL(3 + ab)); mbar_arrive(RAW_EMPTY(sr)); }
                    if (++sr == C2H_NRAW) { sr = 0; pr ^= 1u; }
                    ab ^= 1;
                }
            }
            if (p.overflow && amax >= 6.0e4f) *p.overflow = 1;
        } else {
        // ===== transform warps: two sets of four (set g owns the operand stages with an even / odd running index, so a
        // set's per-stage instruction stream -- ~550 cycles -- has two stage-times to complete); thread <-> pixel row <->
        // TMEM lane; 16 raw channels -> hi/lo -> tcgen05.st =====
        const int g = warp >> 2;
        const int pp = (warp & 3) * 32 + lane;
        const int ty = pp >> p.tw_log2, tx = pp & tw_mask;
        const uint32_t src_row = (uint32_t)(pp * 128);
        const uint32_t sw = (uint32_t)(pp & 7);                              // TMA SWIZZLE_128B: 16 B chunk ^= row % 8
        const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + Cfg::A_TMEM_COL;
        const bool need_mask = p.in_b != nullptr;
        const uint32_t tab_s = smem_u32(tab_a);
        int sr = 0, so = 0;
        uint32_t pr = 0, po = 0;
        int tab_n = -1;
        int bi = 0;                                                          // owner set of the next raw box (running index mod C2_NS)
        int last_owner = -1;                                                 // owner of the most recent box
        int pend = -1;                                                       // operand slot stored but not yet published
        float amax = 0.f;                                                    // largest |operand| this thread converted to fp16
        for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
            const Tile tl = C2_FIRST_TILE(t) ? tile0 : decode(t);
            const int hb = (tl.ho0 + ty) * p.stride - p.pad, wb = (tl.wo0 + tx) * p.stride - p.pad;
            if (affine && tl.n != tab_n) {
                // per-(sample, channel) coefficient table of this image; only the transform warps touch it
                if (!waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); waited = true; }
                asm volatile("bar.sync 1, %0;" ::"n"(C2_XT) : "memory");
                const int cpad = p.ncc * C2_KC;
                for (int c = threadIdx.x; c < cpad; c += C2_XT) {
                    const bool ok = c < p.Cin;
                    tab_a[c] = ok ? (p.in_a ? __ldg(p.in_a + (size_t)tl.n * p.Cin + c) : 1.f) : 0.f;
                    tab_b[c] = (ok && p.in_b) ? __ldg(p.in_b + (size_t)tl.n * p.Cin + c) : 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(C2_XT) : "memory");
                tab_n = tl.n;
            }
            int tap = tl.tap0, cc = tl.cc0;
            int tr = 0, ts = 0;                                              // filter row / column of `tap` (padding mask)
            bool ok = true;
            if (need_mask) {
                tr = tap / p.kw; ts = tap - tr * p.kw;
                ok = (unsigned)(hb + tr * p.dil) < (unsigned)p.H && (unsigned)(wb + ts * p.dil) < (unsigned)p.W;
            }
            // A set owns whole raw boxes (32 channels = two operand stages, one for an odd tail), alternating with the
            // other set; the box's RAW_FULL / RAW_EMPTY barriers are touched by the owner only (RAW_EMPTY counts its four
            // warps), so a set spends a handful of counter updates on a box it does not own.
            for (int rb = tl.r0; rb < tl.r1; ++rb) {
                const int nst = (cc + 1 < p.ncc) ? 2 : 1;
                if (bi == g) {
                    mbar_wait(RAW_FULL(sr), pr);
                    const uint32_t rawb = raw0 + sr * C2_RAW_BYTES + src_row;
                    int so_s = so;
                    uint32_t po_s = po;
#pragma unroll 1
                    for (int hs = 0; hs < nst; ++hs) {
                        const int ccs = cc + hs;
                        if (threadIdx.x == 0 || threadIdx.x == 128 || threadIdx.x == 256) C2_TRACE(1, tap * p.ncc + ccs);
                        float v[16];
                        const uint32_t c8 = (uint32_t)hs * 4u;
                        if (C2_DBG(8)) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) v[e] = 0.f;
                        } else
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(v[4 * j]), "=f"(v[4 * j + 1]), "=f"(v[4 * j + 2]), "=f"(v[4 * j + 3])
                                         : "r"(rawb + (((c8 + j) ^ sw) << 4)));
                        if (affine && !C2_DBG(8)) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float4 a4, b4;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                             : "=f"(a4.x), "=f"(a4.y), "=f"(a4.z), "=f"(a4.w)
                                             : "r"(tab_s + (uint32_t)((ccs * C2_KC + j * 4) * 4)));
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                             : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                                             : "r"(tab_s + (uint32_t)((C2_MAX_AFFINE_C + ccs * C2_KC + j * 4) * 4)));
                                ffma2(v[4 * j], v[4 * j + 1], a4.x, a4.y, b4.x, b4.y);
                                ffma2(v[4 * j + 2], v[4 * j + 3], a4.z, a4.w, b4.z, b4.w);
                            }
                            if (p.in_relu) {
#pragma unroll
                                for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
                            }
                            if (!__all_sync(0xffffffffu, ok)) {              // warp-uniform: interior tiles skip the selects
#pragma unroll
                                for (int e = 0; e < 16; ++e) v[e] = ok ? v[e] : 0.f;
                            }
                        }
                        float o[Cfg::A_COLS];
                        if (F16) {
                            // split-fp16 operand: x = hi + 2^-11 * lo with hi = fp16(x), lo = fp16((x - hi) * 2^11): 22
                            // mantissa bits like the TF32 pair, but one K = 16 MMA per term instead of two K = 8 ones.
                            // Columns of the stage: [hi: k0..15, two per column][lo: k0..15]; |x| saturates at the fp16
                            // range (65504).  hi - x comes from the mixed-precision subtract (one FHADD on the packed
                            // half, exact in fp32) instead of a half -> float conversion and an FADD; the affine and the
                            // 2^11 scale are packed f32x2 operations.
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                uint32_t h, l;
                                float d0, d1;
                                // range guard: one three-input FMNMX per two operands; checked once at the end of the kernel
                                asm("max.abs.f32 %0, %0, %1, %2;" : "+f"(amax) : "f"(v[2 * j]), "f"(v[2 * j + 1]));
                                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
                                asm("{\n\t.reg .b16 e0, e1;\n\tmov.b32 {e0, e1}, %2;\n\t"
                                    "sub.rn.f32.f16 %0, e0, %3;\n\tsub.rn.f32.f16 %1, e1, %4;\n\t}"
                                    : "=f"(d0), "=f"(d1) : "r"(h), "f"(v[2 * j]), "f"(v[2 * j + 1]));
                                fmul2(d0, d1, -2048.f);
                                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(d1), "f"(d0));
                                o[j] = __uint_as_float(h);
                                o[8 + j] = __uint_as_float(l);
                            }
                        } else {
                            // columns of the stage: [ks0: hi k0..7 | lo k0..7][ks1: hi k8..15 | lo k8..15]
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                float h, l;
                                split_tf32(v[e], h, l);
                                o[(e >> 3) * 16 + (e & 7)] = h;
                                o[(e >> 3) * 16 + 8 + (e & 7)] = l;
                            }
                        }
                        if (threadIdx.x == 0 || threadIdx.x == 128 || threadIdx.x == 256) C2_TRACE(2, tap * p.ncc + ccs);
                        // the tensor-memory store of the set's previous stage completes under this stage's ALU work: it
                        // is waited for and published only now (the MMA side runs stages behind, nothing waits on it)
                        if (pend >= 0) {
                            tmem_st_wait();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(OP_FULL(pend >> 1));
                        }
                        // the slot's pair was released as a whole: one wait per pair the box touches
                        if (hs == 0 || (so_s & 1) == 0) mbar_wait(OP_EMPTY(so_s >> 1), po_s ^ 1u);
                        if (threadIdx.x == 0 || threadIdx.x == 128 || threadIdx.x == 256) C2_TRACE(3, tap * p.ncc + ccs);
                        tc_fence_after();
                        if (F16) tmem_st16(a_lane + (uint32_t)(so_s * Cfg::A_COLS), o);
                        else     tmem_st32(a_lane + (uint32_t)(so_s * Cfg::A_COLS), o);
                        pend = so_s;
                        if (threadIdx.x == 0 || threadIdx.x == 128 || threadIdx.x == 256) C2_TRACE(4, tap * p.ncc + ccs);
                        if (++so_s == C2_NO) { so_s = 0; po_s ^= 1u; }
                    }
                    // the box's last stage is published before the set moves on (deferred to the set's next box, two
                    // stage times later, it would hold back the pair the MMA side waits for)
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();                                            // (every lane has also read its raw row)
                    if (lane == 0) { mbar_arrive(OP_FULL(pend >> 1)); mbar_arrive(RAW_EMPTY(sr)); }
                    pend = -1;
                }
                last_owner = bi;
                if (++bi == C2_NS) bi = 0;
                if (++sr == C2_NR) { sr = 0; pr ^= 1u; }
                so += nst;
                if (so >= C2_NO) { so -= C2_NO; po ^= 1u; }
                cc += nst;
                if (cc == p.ncc) {
                    cc = 0; ++tap;
                    if (need_mask) {
                        if (++ts == p.kw) { ts = 0; ++tr; }
                        ok = (unsigned)(hb + tr * p.dil) < (unsigned)p.H && (unsigned)(wb + ts * p.dil) < (unsigned)p.W;
                    }
                }
            }
        }
        // an odd number of stages leaves the last pair half filled: the issuers wait for whole pairs, so one set (and the
        // weight producer) supply the missing stage's arrivals
        // (the set that owned the last box: its real arrival for that pair is already in, so the phase is the right one)
        if ((so & 1) && last_owner == g && lane == 0) mbar_arrive(OP_FULL(so >> 1));
        // cvt.rn.satfinite clamps at 65504: a clamped operand means a wrong result, so it is reported (sticky word; the
        // host re-runs the frame with 3xTF32 operands, which have the fp32 exponent range)
        if (F16 && p.overflow && amax >= 6.0e4f) *p.overflow = 1;
        }
    } else if (warp < C2_XW + 8) {
        // ===== drain warps: MAIN accumulator chunks -> fp32 registers (round-to-nearest adds), then epilogue =====
        // (the register file is re-partitioned between the warpgroups: these two hold TN/2 accumulators + a 32-wide
        // tcgen05.ld per thread; the producer / issuer warpgroup gives its share back)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C2_REGS_DRAIN));
        asm volatile("griddepcontrol.wait;" ::: "memory");
        constexpr int NC = TN / 2;
        const int dwp = warp - C2_XW;
        const int q = dwp & 3, half = dwp >> 2;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * NC);
        const uint32_t stg = smem_u32(smem + Cfg::STG_OFF) + (uint32_t)(dwp * 4096);
        int b = 0, cb = 0;
        uint32_t ph0 = 0u, ph1 = 0u, pcf0 = 0u, pcf1 = 0u;
        for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
            const Tile tl = C2_FIRST_TILE(t) ? tile0 : decode(t);
            float acc[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c] = 0.f;
            const int nchunks = (tl.it1 - tl.it0 + p.chunk - 1) / p.chunk;
            for (int ch = 0; ch < nchunks; ++ch) {
                if (b == 0) { mbar_wait(MAIN_FULL(0), ph0); ph0 ^= 1u; }
                else        { mbar_wait(MAIN_FULL(1), ph1); ph1 ^= 1u; }
                if (threadIdx.x == C2_XT) C2_TRACE(11, tl.it0 + ch * p.chunk);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < NC; c0 += 32) {
                    float v[32];
                    tmem_ld32(tlane + (uint32_t)(b * TN + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc[c0 + e] += v[e];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(MAIN_EMPTY(b));
                if (threadIdx.x == C2_XT) C2_TRACE(12, tl.it0 + ch * p.chunk);
                b ^= 1;
            }
            // the correction terms are issued by their own warp: wait for its end-of-tile commit
            if (cb == 0) { mbar_wait(CORR_FULL(0), pcf0); pcf0 ^= 1u; }
            else         { mbar_wait(CORR_FULL(1), pcf1); pcf1 ^= 1u; }
            tc_fence_after();
#pragma unroll
            for (int ci = 0; ci < NCI; ++ci)
#pragma unroll
                for (int c0 = 0; c0 < NC; c0 += 32) {
                    float v[32];
                    tmem_ld32(tlane + (uint32_t)((2 + ci * NCB + cb) * TN + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc[c0 + e] = F16 ? fmaf(v[e], 1.0f / 2048.f, acc[c0 + e]) : acc[c0 + e] + v[e];
                }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(CORR_EMPTY(cb));          // the issuer may start the next tile on this buffer
            if (threadIdx.x == C2_XT) C2_TRACE(13, tl.it0);
            if (threadIdx.x == C2_XT && t == (int)blockIdx.x) C2_MARK(4);
            if (SK == 2 && tl.nsl > 1) {
                // ---- tail splitting: slices 1.. park their partial accumulators (this thread's row, its NC channels) in the
                // workspace and move on; slice 0 waits for them, adds them in slice order (deterministic) and runs the normal
                // epilogue -- bias, residual, ReLU, statistics, all as for an unsplit tile.  Slices never wait for slice 0 and all
                // slices of the tail are in one wave (launcher), so nothing can block.
                const int tt = tl.tile - p.n_plain;
                int* cnt = p.ws_cnt + tt;
                float* mine = p.ws + ((size_t)tt * (tl.nsl - 1)) * (C2_BM * TN) + (size_t)(q * 32 + lane) * TN + half * NC;
                if (tl.ks > 0) {
                    float* dst = mine + (size_t)(tl.ks - 1) * (C2_BM * TN);
#pragma unroll
                    for (int c = 0; c < NC; c += 4)
                        __stcg(reinterpret_cast<float4*>(dst + c), make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]));
                    __threadfence();
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    if (threadIdx.x == C2_XT) atomicAdd(cnt, 1);
                    continue;                                        // no epilogue for this slice
                }
                if (threadIdx.x == C2_XT) {
                    int v;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
                    } while (v < tl.nsl - 1);
                    *cnt = 0;                                        // ready for the next launch (every slice has arrived)
                    __threadfence();
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                for (int sl = 0; sl < tl.nsl - 1; ++sl) {
                    const float* src = mine + (size_t)sl * (C2_BM * TN);
#pragma unroll
                    for (int c = 0; c < NC; c += 4) {
                        const float4 v4 = __ldcg(reinterpret_cast<const float4*>(src + c));
                        acc[c] += v4.x; acc[c + 1] += v4.y; acc[c + 2] += v4.z; acc[c + 3] += v4.w;
                    }
                }
            }
            if (NCB == 2) cb ^= 1;
            // ---- epilogue.  The accumulators sit one pixel row per thread; stored that way every warp store would touch 32
            // different lines (and the residual loads likewise).  Each warp therefore transposes 32 pixels x 32 channels
            // through a private, XOR-swizzled 4 KB staging block: on the way back 8 lanes cover the 128 contiguous bytes
            // of one pixel, so residual loads and output stores are whole 128-byte lines, the bias is one float4 per lane,
            // and the GroupNorm / GCT statistics need two shuffles per value instead of a 32-lane butterfly.
            // The 32-channel chunks of a warp run through ONE copy of this code (rolled loop, the upper chunk rotated into
            // acc[0..31]): the unrolled version was 3 000 instructions per kernel, and on the short-K layers -- where the
            // epilogue is the critical path -- a third of the drain warps' stalls were instruction-cache misses.
            const int cbase = tl.n0 + half * NC;
            constexpr bool partial = SK == 1;                        // split-K: raw partial sums of this K slice to `ws`
            float* const ybase = partial ? p.ws + (size_t)tl.ks * ((size_t)p.N * p.Ho * p.Wo) * p.Cout : p.y;
            const int ldo = partial ? p.Cout : p.ldy;
            const float* const bias = partial ? nullptr : p.bias;
            const float* const resb = partial ? nullptr : p.res;
            const bool relu = !partial && p.relu;
            const int r_sub = lane >> 3, ch4 = lane & 7;
            // statistics row of this warp: (pixel tile, 32-pixel quadrant) -- no cross-warp step, no block barrier
            float* const strow = (!partial && p.tile_stats) ? p.tile_stats + (size_t)((tl.tile / p.tiles_n) * 4 + q) * 2 * p.Cout : nullptr;
            int pixi[8], rowoff[8];                                  // output pixel of row i * 4 + r_sub (-1: outside), its offset in y
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = q * 32 + i * 4 + r_sub;
                const int ho = tl.ho0 + (m >> p.tw_log2), wo = tl.wo0 + (m & tw_mask);
                pixi[i] = (ho < p.Ho && wo < p.Wo) ? (tl.n * p.Ho + ho) * p.Wo + wo : -1;
                rowoff[i] = pixi[i] * ldo;                           // (the launcher checks that the output fits 2^31 elements)
            }
#pragma unroll 1
            for (int c0 = 0; c0 < NC; c0 += 32) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4))),
                                 "f"(acc[4 * j]), "f"(acc[4 * j + 1]), "f"(acc[4 * j + 2]), "f"(acc[4 * j + 3])
                                 : "memory");
                if (NC > 32) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc[e] = acc[(32 + e) % NC];
                }
                __syncwarp();
                const int co = cbase + c0 + ch4 * 4;
                const bool vec = p.vec_out && co + 3 < p.Cout;
                float b4[4] = {0.f, 0.f, 0.f, 0.f};
                if (bias) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (co + e < p.Cout) b4[e] = __ldg(bias + co + e);
                }
                float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
                if (vec) {
                    // four rows at a time: their staging reads (and residual lines) are requested before the first use;
                    // bias / residual / ReLU / statistics cost instructions only where the layer has them (warp-uniform
                    // branches), the statistics are packed f32x2 operations
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float o[4][4];
                        float4 r4[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int i = h * 4 + k, r = i * 4 + r_sub;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(o[k][0]), "=f"(o[k][1]), "=f"(o[k][2]), "=f"(o[k][3])
                                         : "r"(stg + (uint32_t)(r * 128 + ((ch4 ^ (r & 7)) << 4))));
                            if (resb) r4[k] = pixi[i] >= 0 ? ldg4(resb + (size_t)pixi[i] * p.ldres + co) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int i = h * 4 + k;
                            if (pixi[i] >= 0) {
                                if (resb) { o[k][0] += r4[k].x; o[k][1] += r4[k].y; o[k][2] += r4[k].z; o[k][3] += r4[k].w; }
                                if (bias) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) o[k][e] += b4[e];
                                }
                                if (relu) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) o[k][e] = fmaxf(o[k][e], 0.f);
                                }
                                if (strow) {
                                    fadd2(s1[0], s1[1], o[k][0], o[k][1]);
                                    fadd2(s1[2], s1[3], o[k][2], o[k][3]);
                                    fsqacc2(s2[0], s2[1], o[k][0], o[k][1]);
                                    fsqacc2(s2[2], s2[3], o[k][2], o[k][3]);
                                }
                                *reinterpret_cast<float4*>(ybase + (size_t)(rowoff[i] + co)) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
                            }
                        }
                    }
                } else {
#pragma unroll 1
                    for (int i = 0; i < 8; ++i) {                    // ragged channel count / unaligned rows: scalar accesses
                        const int r = i * 4 + r_sub;
                        float o[4];
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3])
                                     : "r"(stg + (uint32_t)(r * 128 + ((ch4 ^ (r & 7)) << 4))));
                        int pix = -1;
#pragma unroll
                        for (int k = 0; k < 8; ++k) pix = (k == i) ? pixi[k] : pix;
                        if (pix >= 0) {
                            float* dst = ybase + (size_t)pix * ldo + co;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (co + e < p.Cout) {
                                    if (resb) o[e] += __ldg(resb + (size_t)pix * p.ldres + co + e);
                                    o[e] += b4[e];
                                    if (relu) o[e] = fmaxf(o[e], 0.f);
                                    dst[e] = o[e];
                                    s1[e] += o[e]; s2[e] = fmaf(o[e], o[e], s2[e]);
                                }
                            }
                        }
                    }
                }
                __syncwarp();                                        // the staging block is rewritten by the next chunk
                if (strow) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 8);
                        s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 8);
                        s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
                        s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
                    }
                    if (lane < 8) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (co + e < p.Cout) { strow[co + e] = s1[e]; strow[p.Cout + co + e] = s2[e]; }
                    }
                }
            }
            if (threadIdx.x == C2_XT) C2_TRACE(14, tl.it0);
            if (partial) {
                // ---- split-K finish, fused (the separate finish kernel -- 65 launches per frame on the 31x54 backbone maps --
                // is gone).  Every K slice of a tile runs on its own CTA at the same time (the launcher splits only when
                // tiles x slices <= SMs, one work item per CTA), so the slices MEET at the tile's counter: each writes its
                // partial tile, arrives, waits until all `ksplit` slices have arrived, and then finishes its share of the
                // tile's 128 pixel rows -- the partial tiles are added in slice order (deterministic), bias / residual /
                // ReLU applied, y written.  The other slices' sums are read with ld.global.cg (written by other SMs during
                // this launch).  A second round of arrivals tells the last slice to clear the counter for the next launch.
                int* cnt = p.ws_cnt + tl.tile;
                __threadfence();
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (threadIdx.x == C2_XT) {
                    atomicAdd(cnt, 1);
                    C2_MARK(5);
                    int v;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
                    } while (v < p.ksplit);
                    __threadfence();
                    C2_MARK(6);
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                {
                    const size_t Mtot = (size_t)p.N * p.Ho * p.Wo;
                    constexpr int C4 = TN / 4;
                    const int m_lo = C2_BM * tl.ks / p.ksplit, m_hi = C2_BM * (tl.ks + 1) / p.ksplit;
                    for (int i = m_lo * C4 + (threadIdx.x - C2_XT); i < m_hi * C4; i += 256) {
                        const int m = i / C4, c = (i - m * C4) * 4;
                        const int ho = tl.ho0 + (m >> p.tw_log2), wo = tl.wo0 + (m & tw_mask);
                        const int co = tl.n0 + c;
                        if (ho >= p.Ho || wo >= p.Wo || co >= p.Cout) continue;   // (Cout % 4 == 0 on this path)
                        const size_t pix = ((size_t)tl.n * p.Ho + ho) * p.Wo + wo;
                        // every slice's partial sum is requested before the first add (one L2 round trip instead of
                        // ksplit dependent ones); the adds keep the slice order
                        constexpr int NF = 5;                            // slices in flight (the 31x54 backbone maps split 5 ways)
                        float4 v[NF];
#pragma unroll
                        for (int k = 0; k < NF; ++k)
                            if (k < p.ksplit)
                                v[k] = __ldcg(reinterpret_cast<const float4*>(p.ws + ((size_t)k * Mtot + pix) * p.Cout + co));
                        float4 a = v[0];
#pragma unroll
                        for (int k = 1; k < NF; ++k)
                            if (k < p.ksplit) { a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w; }
                        for (int k = NF; k < p.ksplit; ++k) {
                            const float4 u = __ldcg(reinterpret_cast<const float4*>(p.ws + ((size_t)k * Mtot + pix) * p.Cout + co));
                            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
                        }
                        if (p.bias) { const float4 b = ldg4(p.bias + co); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
                        if (p.res) {
                            const float4 r = ldg4(p.res + pix * p.ldres + co);
                            a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
                        }
                        if (p.relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
                        *reinterpret_cast<float4*>(p.y + pix * p.ldy + co) = a;
                    }
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");          // every thread of this slice has read the partial sums
                if (threadIdx.x == C2_XT) C2_MARK(7);
                if (threadIdx.x == C2_XT && atomicAdd(cnt, 1) == 2 * p.ksplit - 1) *cnt = 0;
            }
        }
    } else if (warp == C2_XW + 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C2_REGS_MISC));
        if (lane == 0) {
            // ===== activation TMA producer: one 128-pixel x 32-channel box per two operand stages =====
            int sr = 0;
            uint32_t pr = 0;
            bool waited = false;                 // griddepcontrol.wait executed (see the prologue)
            if constexpr (HALO) {
                // halo: ONE box per 32-channel group -- the (16 + 2d) x (8 + 2d) pixel patch all nine taps read
                const uint32_t bytes = (uint32_t)((C2H_TW + 2 * p.dil) * (C2H_TH + 2 * p.dil) * 128);
                for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
                    const Tile tl = C2_FIRST_TILE(t) ? tile0 : decode(t);
                    for (int cb = 0; cb < nrc; ++cb) {
                        mbar_wait(RAW_EMPTY(sr), pr ^ 1u);
                        C2_TRACE(0, 18 * cb);
                        mbar_arrive_expect_tx(RAW_FULL(sr), bytes);
                        if (!waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); waited = true; }
                        tma_load_4d(raw0 + sr * C2H_RAW_BYTES, &tmapA, cb * C2_RKC, tl.wo0 - p.dil, tl.ho0 - p.dil, tl.n, RAW_FULL(sr));
                        if (++sr == C2H_NRAW) { sr = 0; pr ^= 1u; }
                    }
                }
            } else
            for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
                const Tile tl = C2_FIRST_TILE(t) ? tile0 : decode(t);
                const int wbase = tl.wo0 * p.stride - p.pad, hbase = tl.ho0 * p.stride - p.pad;
                int tap = tl.tap0, rc = tl.cc0 >> 1;
                int r = tap / p.kw, s = tap - r * p.kw;
                for (int rr = tl.r0; rr < tl.r1; ++rr) {
                    mbar_wait(RAW_EMPTY(sr), pr ^ 1u);
                    C2_TRACE(0, 2 * rr);
                    if (C2_DBG(2)) {
                        mbar_arrive(RAW_FULL(sr));
                    } else {
                        mbar_arrive_expect_tx(RAW_FULL(sr), C2_RAW_BYTES);
                        if (!waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); waited = true; C2_MARK(2); }
                        tma_load_4d(raw0 + sr * C2_RAW_BYTES, &tmapA, rc * C2_RKC, wbase + s * p.dil, hbase + r * p.dil,
                                    tl.n, RAW_FULL(sr));
                    }
                    if (++sr == C2_NR) { sr = 0; pr ^= 1u; }
                    if (++rc == nrc) {
                        rc = 0; ++tap;
                        if (++s == p.kw) { s = 0; ++r; }
                    }
                }
            }
        }
    } else if (warp == C2_XW + 9) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C2_REGS_MISC));
        if (lane == 0) {
            // ===== weight TMA producer: chunk (row block, stage) = [ks0: hi | lo][ks1: hi | lo], 4096 B blocks =====
            int sb_ = 0;
            uint32_t pb = 0;
            for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
                const Tile tl = C2_FIRST_TILE(t) ? tile0 : decode(t);
                const int rb = tl.n0 / C2_WRB;
                const uint8_t* wsrc = p.w + (size_t)rb * nIt * C2_WCHUNK;
                const uint32_t sub = (uint32_t)(tl.n0 % C2_WRB) * 32u;       // byte offset of row n0 inside a 4096 B block
                // halo: the K loop runs (channel box, tap, half box) instead of (tap, stage); the packed image keeps its order
                int h_cb = 0, h_tap = 0, h_half = 0;
                for (int it_ = tl.it0; it_ < tl.it1; ++it_) {
                    int it = it_;
                    if (HALO) {
                        it = h_tap * p.ncc + 2 * h_cb + h_half;
                        const int nh = (2 * h_cb + 1 < p.ncc) ? 2 : 1;
                        if (++h_half == nh) { h_half = 0; if (++h_tap == 9) { h_tap = 0; ++h_cb; } }
                    }
                    if ((sb_ & 1) == 0) mbar_wait(OP_EMPTY(sb_ >> 1), pb ^ 1u);
                    C2_TRACE(8, it);
                    const uint32_t bfull = OP_FULL(sb_ >> 1);
                    if (C2_DBG(1)) {
                        mbar_arrive(bfull);
                        if (++sb_ == C2_NB) { sb_ = 0; pb ^= 1u; }
                        continue;
                    }
                    mbar_arrive_expect_tx(bfull, Cfg::B_BYTES);
                    const uint32_t sb = op0 + sb_ * Cfg::B_BYTES;
                    const uint8_t* src = wsrc + (size_t)it * C2_WCHUNK;
                    if (TN == C2_WRB) {
                        bulk_g2s(sb, src, C2_WCHUNK, bfull);
                    } else {
#pragma unroll
                        for (int blk = 0; blk < (F16 ? 2 : 4); ++blk)
                            bulk_g2s(sb + blk * (TN * 32), src + blk * 4096 + sub, TN * 32, bfull);
                    }
                    if (++sb_ == C2_NB) { sb_ = 0; pb ^= 1u; }
                }
            }
            if (sb_ & 1) mbar_arrive(OP_FULL(sb_ >> 1));          // odd stage count: complete the last pair (see transform)
        }
    } else if (warp == C2_XW + 10) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C2_REGS_MISC));
        // ===== MAIN issuer (hi*hi): the whole warp runs the (warp-uniform) loop, one elected lane issues =====
        const uint32_t idesc = F16 ? idesc_f16(C2_BM, TN) : idesc_tf32(C2_BM, TN);
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        int so = 0, b = 0;
        uint32_t po = 0, pe0 = 0, pe1 = 0;
        if constexpr (HALO) {
            // halo: the A operand is the transformed halo tile in shared memory; tap (r, s) of a box = the same tile read from
            // the start address of halo pixel (r d, s d): 16 groups of 8 rows (= 8 consecutive pixels of a halo row, 16 B
            // each) one halo row apart (SBO), the two k-groups of a stage one pixel plane apart (LBO)
            const int d = p.dil;
            const int HWd = C2H_TW + 2 * d, P = HWd * (C2H_TH + 2 * d);
            const uint32_t a0 = smem_u32(smem + Cfg::A_OFF);
            int ab = 0;
            uint32_t pf0 = 0, pf1 = 0;
            for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
                int in_chunk = 0, it = 0;
                for (int cb = 0; cb < nrc; ++cb) {
                    const int nh = (2 * cb + 1 < p.ncc) ? 2 : 1;
                    if (ab == 0) { mbar_wait(RAW_FULL(3), pf0); pf0 ^= 1u; }
                    else         { mbar_wait(RAW_FULL(4), pf1); pf1 ^= 1u; }
                    const uint32_t ahi = a0 + ab * C2H_A_BUF;
                    for (int tap = 0; tap < 9; ++tap) {
                        const int tr = tap / 3, ts = tap - 3 * tr;
                        const uint32_t tapoff = (uint32_t)((tr * d * HWd + ts * d) * 16);
                        for (int half = 0; half < nh; ++half, ++it) {
                            if (lane == 0) C2_TRACE(5, it);
                            if ((so & 1) == 0) mbar_wait(OP_FULL(so >> 1), po);      // the weights of both stages of the pair
                            if (lane == 0) C2_TRACE(6, it);
                            const int bb = b, ic = in_chunk;
                            if (in_chunk == 0) {
                                if (b == 0) { mbar_wait(MAIN_EMPTY(0), pe0 ^ 1u); pe0 ^= 1u; }
                                else        { mbar_wait(MAIN_EMPTY(1), pe1 ^ 1u); pe1 ^= 1u; }
                            }
                            const bool ls = (in_chunk + 1 == p.chunk) || (it == nIt - 1);
                            if (ls) { b ^= 1; in_chunk = 0; } else { ++in_chunk; }
                            const bool box_end = tap == 8 && half == nh - 1;
                            tc_fence_after();
                            if (elect_one()) {
                                const uint32_t sb = op0 + so * Cfg::B_BYTES;
                                const uint64_t adesc = smem_desc(ahi + tapoff + (uint32_t)(half * 2 * P * 16), (uint32_t)(P * 16), (uint32_t)(HWd * 16));
                                mma_f16(tb + (uint32_t)(bb * TN), adesc, smem_desc(sb, LBO_BYTES, SBO_BYTES), idesc, ic > 0 ? 1u : 0u);
                                if (so & 1) mma_commit(OP_EMPTY(so >> 1));
                                if (ls) mma_commit(MAIN_FULL(bb));
                                if (box_end) mma_commit(RAW_EMPTY(3 + ab));      // this issuer is done with the operand tile
                            }
                            __syncwarp();
                            if (lane == 0) C2_TRACE(7, it);
                            if (++so == C2_NO) { so = 0; po ^= 1u; }
                        }
                    }
                    ab ^= 1;
                }
            }
        } else
        for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
            const Tile tl = C2_FIRST_TILE(t) ? tile0 : decode(t);
            int in_chunk = 0;
            // Two stages (one operand pair) per iteration where the pair lies inside the tile: one barrier wait, one
            // tcgen05 fence, one elect and one warp sync for both -- this warp's iteration is the period of the kernel
            // (DESIGN section 4), every instruction in front of its tcgen05.mma counts.
            for (int it = tl.it0; it < tl.it1;) {
                const int n = ((so & 1) == 0 && it + 1 < tl.it1) ? 2 : 1;
                if (lane == 0) C2_TRACE(5, it);
                if ((so & 1) == 0) mbar_wait(OP_FULL(so >> 1), po);      // both operands of both stages of the pair
                if (lane == 0) C2_TRACE(6, it);
                // accumulator bookkeeping of the stages: a chain of `chunk` stages per MAIN buffer, a chain's first stage
                // waits for the drain warps to have read the buffer's previous chain
                int bb[2], ic[2];
                bool ls[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (k < n) {
                        bb[k] = b; ic[k] = in_chunk;
                        if (in_chunk == 0) {
                            if (b == 0) { mbar_wait(MAIN_EMPTY(0), pe0 ^ 1u); pe0 ^= 1u; }
                            else        { mbar_wait(MAIN_EMPTY(1), pe1 ^ 1u); pe1 ^= 1u; }
                        }
                        ls[k] = (in_chunk + 1 == p.chunk) || (it + k == tl.it1 - 1);
                        if (ls[k]) { b ^= 1; in_chunk = 0; } else { ++in_chunk; }
                    }
                }
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        if (k < n) {
                            const int slot = so + k;                     // n == 2 only on an even slot: no wrap inside
                            const uint32_t sb = op0 + slot * Cfg::B_BYTES;
                            const uint32_t d_main = tb + (uint32_t)(bb[k] * TN);
                            const uint32_t ta = tb + Cfg::A_TMEM_COL + (uint32_t)(slot * Cfg::A_COLS);
                            if (C2_DBG(16)) {
                            } else if (F16) {
                                mma_f16_ts(d_main, ta, smem_desc(sb, LBO_BYTES, SBO_BYTES), idesc, ic[k] > 0 ? 1u : 0u);
                            } else {
                                mma_tf32_ts(d_main, ta, smem_desc(sb, LBO_BYTES, SBO_BYTES), idesc, ic[k] > 0 ? 1u : 0u);
                                mma_tf32_ts(d_main, ta + 16, smem_desc(sb + TN * 64, LBO_BYTES, SBO_BYTES), idesc, 1u);
                            }
                            if (slot & 1) mma_commit(OP_EMPTY(slot >> 1));
                            if (ls[k]) mma_commit(MAIN_FULL(bb[k]));
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) C2_TRACE(7, it);
                it += n;
                so += n;
                if (so == C2_NO) { so = 0; po ^= 1u; }
            }
        }
    } else if (warp >= C2_XW + 11 && warp < C2_XW + 11 + NCI) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C2_REGS_MISC));
        // ===== CORR issuer(s): lo*hi + hi*lo into CORR (one warp), or one term and one accumulator per warp =====
        const int ci = warp - (C2_XW + 11);
        const uint32_t idesc = F16 ? idesc_f16(C2_BM, TN) : idesc_tf32(C2_BM, TN);
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        int so = 0, cb = 0;
        uint32_t po = 0, pc0 = 0, pc1 = 0;
        if constexpr (HALO) {
            const int d = p.dil;
            const int HWd = C2H_TW + 2 * d, P = HWd * (C2H_TH + 2 * d);
            const uint32_t a0 = smem_u32(smem + Cfg::A_OFF);
            int ab = 0;
            uint32_t pf0 = 0, pf1 = 0;
            for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
                const uint32_t d_corr = tb + (uint32_t)((2 + cb) * TN);
                if (cb == 0) { mbar_wait(CORR_EMPTY(0), pc0 ^ 1u); pc0 ^= 1u; }
                else         { mbar_wait(CORR_EMPTY(1), pc1 ^ 1u); pc1 ^= 1u; }
                int it = 0;
                for (int bx = 0; bx < nrc; ++bx) {
                    const int nh = (2 * bx + 1 < p.ncc) ? 2 : 1;
                    if (ab == 0) { mbar_wait(RAW_FULL(3), pf0); pf0 ^= 1u; }
                    else         { mbar_wait(RAW_FULL(4), pf1); pf1 ^= 1u; }
                    const uint32_t ahi = a0 + ab * C2H_A_BUF, alo = ahi + C2H_A_HALF;
                    for (int tap = 0; tap < 9; ++tap) {
                        const int tr = tap / 3, ts = tap - 3 * tr;
                        const uint32_t tapoff = (uint32_t)((tr * d * HWd + ts * d) * 16);
                        for (int half = 0; half < nh; ++half, ++it) {
                            if ((so & 1) == 0) mbar_wait(OP_FULL(so >> 1), po);
                            if (lane == 0) C2_TRACE(9, it);
                            const bool box_end = tap == 8 && half == nh - 1;
                            tc_fence_after();
                            if (elect_one()) {
                                const uint32_t sb = op0 + so * Cfg::B_BYTES;
                                const uint32_t koff = tapoff + (uint32_t)(half * 2 * P * 16);
                                // lo(A) * hi(B), then hi(A) * lo(B): weight stage = [hi block | lo block] of TN rows x 32 B
                                mma_f16(d_corr, smem_desc(alo + koff, (uint32_t)(P * 16), (uint32_t)(HWd * 16)),
                                        smem_desc(sb, LBO_BYTES, SBO_BYTES), idesc, it > 0 ? 1u : 0u);
                                mma_f16(d_corr, smem_desc(ahi + koff, (uint32_t)(P * 16), (uint32_t)(HWd * 16)),
                                        smem_desc(sb + TN * 32, LBO_BYTES, SBO_BYTES), idesc, 1u);
                                if (so & 1) mma_commit(OP_EMPTY(so >> 1));
                                if (it == nIt - 1) mma_commit(CORR_FULL(cb));
                                if (box_end) mma_commit(RAW_EMPTY(3 + ab));
                            }
                            __syncwarp();
                            if (lane == 0) C2_TRACE(10, it);
                            if (++so == C2_NO) { so = 0; po ^= 1u; }
                        }
                    }
                    ab ^= 1;
                }
                cb ^= 1;
            }
        } else
        for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
            const uint32_t d_corr = tb + (uint32_t)((2 + ci * NCB + cb) * TN);
            // the CORR buffer of this tile must have been read out by the drain warps (two tiles ago when double buffered)
            if (cb == 0) { mbar_wait(CORR_EMPTY(0), pc0 ^ 1u); pc0 ^= 1u; }
            else         { mbar_wait(CORR_EMPTY(1), pc1 ^ 1u); pc1 ^= 1u; }
            const Tile tl = C2_FIRST_TILE(t) ? tile0 : decode(t);
            for (int it = tl.it0; it < tl.it1;) {                      // two stages per iteration, as in the MAIN issuer
                const int n = ((so & 1) == 0 && it + 1 < tl.it1) ? 2 : 1;
                if ((so & 1) == 0) mbar_wait(OP_FULL(so >> 1), po);
                if (lane == 0) C2_TRACE(9, it);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        if (k < n) {
                            const int slot = so + k, itk = it + k;
                            const uint32_t sb = op0 + slot * Cfg::B_BYTES;
                            const uint32_t ta = tb + Cfg::A_TMEM_COL + (uint32_t)(slot * Cfg::A_COLS);
                            if (C2_DBG(4)) {
                            } else if (F16) {
                                // smem stage = [hi block | lo block] of TN rows x 32 B; TMEM stage = [hi: 8 columns | lo: 8 columns]
                                mma_f16_ts(d_corr, ta + 8, smem_desc(sb, LBO_BYTES, SBO_BYTES), idesc, itk > tl.it0 ? 1u : 0u);
                                mma_f16_ts(d_corr, ta, smem_desc(sb + TN * 32, LBO_BYTES, SBO_BYTES), idesc, 1u);
                            } else
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                const uint32_t b_hi = sb + ks * (TN * 64), b_lo = b_hi + TN * 32;
                                const uint32_t ta_hi = ta + (uint32_t)(ks * 16), ta_lo = ta_hi + 8;
                                const uint32_t first = (itk > tl.it0 || ks > 0) ? 1u : 0u;
                                if (NCI == 1) {
                                    mma_tf32_ts(d_corr, ta_lo, smem_desc(b_hi, LBO_BYTES, SBO_BYTES), idesc, first);
                                    mma_tf32_ts(d_corr, ta_hi, smem_desc(b_lo, LBO_BYTES, SBO_BYTES), idesc, 1u);
                                } else if (ci == 0) {
                                    mma_tf32_ts(d_corr, ta_lo, smem_desc(b_hi, LBO_BYTES, SBO_BYTES), idesc, first);
                                } else {
                                    mma_tf32_ts(d_corr, ta_hi, smem_desc(b_lo, LBO_BYTES, SBO_BYTES), idesc, first);
                                }
                            }
                            if (slot & 1) mma_commit(OP_EMPTY(slot >> 1));
                            if (itk == tl.it1 - 1) mma_commit(CORR_FULL(cb));
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) C2_TRACE(10, it);
                it += n;
                so += n;
                if (so == C2_NO) { so = 0; po ^= 1u; }
            }
            if (NCB == 2) cb ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C2_XW + 10) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// w [Cout][taps][Cin] fp32 -> weight image: [row block of 128][stage it = (tap, cc)][ks][hi|lo][128 rows x 8 floats]
__global__ void conv2_pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int ncc,
                                          int rows_padded, uint8_t* __restrict__ out) {
    const int nIt = taps * ncc;
    const long long total = (long long)rows_padded * nIt * 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i % rows_padded);
        const long long rest = i / rows_padded;
        const int g = (int)(rest & 3);                  // 4-float granule inside the 16-channel stage
        const int it = (int)(rest >> 2);
        const int tap = it / ncc, cc = it - tap * ncc;
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ci = cc * C2_KC + g * 4 + e;
            const float v = (r < Cout && ci < Cin) ? __ldg(w + ((size_t)r * taps + tap) * Cin + ci) : 0.f;
            split_tf32(v, hi[e], lo[e]);
        }
        const int rb = r / C2_WRB, rr = r - rb * C2_WRB;
        const int ks = g >> 1, half = g & 1;
        const size_t base = ((size_t)rb * nIt + it) * c2_wchunk(false) + (size_t)ks * 8192 + elem_offset(rr, half * 4);
        *reinterpret_cast<float4*>(out + base) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(out + base + 4096) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// split-fp16 weight image: [row block of 128][stage it = (tap, cc)][hi | lo][128 rows x 16 halves], lo = (w - hi) * 2^11
__global__ void conv2_pack_weights_f16_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int ncc,
                                              int rows_padded, uint8_t* __restrict__ out) {
    const int nIt = taps * ncc;
    const long long total = (long long)rows_padded * nIt * 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i % rows_padded);
        const long long rest = i / rows_padded;
        const int g = (int)(rest & 1);                  // 8-channel granule (one 16-byte core-matrix row) of the stage
        const int it = (int)(rest >> 1);
        const int tap = it / ncc, cc = it - tap * ncc;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ci = cc * C2_KC + g * 8 + e * 2 + h;
                const float v = (r < Cout && ci < Cin) ? __ldg(w + ((size_t)r * taps + tap) * Cin + ci) : 0.f;
                x[h] = fminf(fmaxf(v, -65504.f), 65504.f);
            }
            const __half2 hh = __floats2half2_rn(x[0], x[1]);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn((x[0] - hf.x) * 2048.f, (x[1] - hf.y) * 2048.f);
            hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
            lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const int rb = r / C2_WRB, rr = r - rb * C2_WRB;
        const size_t base = ((size_t)rb * nIt + it) * c2_wchunk(true) + elem_offset16(rr, g * 8);
        *reinterpret_cast<uint4*>(out + base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(out + base + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
        else
            cudaGetLastError();
    }
    return fn;
}

constexpr size_t C2_WS_HEADER = 4096;   // split-K arrival counters (one int per output tile) in front of the partial sums
int g_conv_splitk = 1;   // aoc_set_option("conv_splitk", 0/1)
int g_conv_tail = 1;     // aoc_set_option("conv_tail", 0/1): K-split of the tiles of a partial last wave (tail splitting)
int g_conv_tail_min_stages = 192;   // aoc_set_option("conv_tail_min_stages", n): shortest K loop (16-channel stages) it is used for
int g_conv_pdl = 1;      // aoc_set_option("conv_pdl", 0/1): programmatic dependent launch of the convolution kernels
int g_conv_narrow_nit = 0;    // aoc_set_option("conv_narrow_nit", stages): 64-wide tiles for K loops shorter than this (measured: never better)

template <int TN, bool F16>
static int launch_conv2(const CUtensorMap& map, const Conv2P& p, int tiles, void* workspace, size_t ws_bytes,
                        cudaStream_t stream) {
    static PerDeviceOnce attr;
    if (attr.first()) {
        cudaFuncSetAttribute(conv2_kernel<TN, 0, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2Cfg<TN, F16>::SMEM);
        cudaFuncSetAttribute(conv2_kernel<TN, 1, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2Cfg<TN, F16>::SMEM);
        cudaFuncSetAttribute(conv2_kernel<TN, 2, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2Cfg<TN, F16>::SMEM);
    }
    Conv2P q = p;
    q.tiles_n = cdiv(p.Cout, TN);
    q.total_tiles = tiles * q.tiles_n;
    const int sms = device_sms();
    // split-K: a layer with fewer tiles than half the SMs (the 31x54 maps of the backbone: 14 pixel tiles) would leave
    // the chip idle and run its whole K loop -- up to 1152 stages -- on a few SMs; slices of >= 4 raw stages (8 operand
    // stages) spread it.  Needs the caller's workspace for the partial sums; the epilogue fusions (statistics) do not apply.
    q.ksplit = 1;
    q.ws = nullptr;
    q.ws_cnt = nullptr;
    const long long M = (long long)p.N * p.Ho * p.Wo;
    const int R = p.taps * ((p.ncc + 1) / 2);
    if (g_conv_splitk && q.total_tiles * 2 <= sms && !p.tile_stats && p.Cout % 4 == 0 && p.vec_out && workspace) {
        int S = sms / q.total_tiles;
        if (S > C2_MAX_KSPLIT) S = C2_MAX_KSPLIT;
        if (S > R / 4) S = R / 4;
        const size_t hdr = C2_WS_HEADER;
        if (ws_bytes < hdr || (size_t)q.total_tiles * sizeof(int) > hdr) S = 1;
        while (S > 1 && hdr + (size_t)S * M * p.Cout * sizeof(float) > ws_bytes) --S;
        if (S > 1) { q.ksplit = S; q.ws = (float*)((char*)workspace + hdr); q.ws_cnt = (int*)workspace; }
    }
    // tail splitting: with T tiles on G persistent CTAs the last wave holds T mod G tiles (336 tiles of a 61 x 107 x 6 layer:
    // two full waves and 40 tiles -- a third of the launch spent at 27 % occupancy).  Those tail tiles are cut into S K slices
    // each, S * tail <= G, so that the last wave fills the chip and lasts 1 / S of a tile (+ the hand-over: slices 1.. park their
    // accumulators in the workspace, slice 0 adds them in order and runs the normal epilogue, statistics included).
    q.tail_mode = 0; q.n_plain = 0;
    // (the hand-over costs ~10 us -- park, fence, spin, add -- so it only pays for long K loops: measured +16 us on a 32-stage
    // 1x1 layer, -20 us on the 288-stage dilated ASPP convolutions)
    if (g_conv_tail && q.ksplit == 1 && q.total_tiles > sms && p.nIt >= g_conv_tail_min_stages && workspace &&
        ws_bytes >= C2_WS_HEADER) {
        const int tail = q.total_tiles % sms;
        int S = tail > 0 ? sms / tail : 0;
        if (S > C2_MAX_KSPLIT) S = C2_MAX_KSPLIT;
        if (S > R / 4) S = R / 4;                                     // >= 4 raw stages (8 operand stages) per slice
        while (S > 1 && C2_WS_HEADER + (size_t)tail * (S - 1) * C2_BM * TN * sizeof(float) > ws_bytes) --S;
        if (S > 1 && (size_t)tail * sizeof(int) <= C2_WS_HEADER) {
            q.tail_mode = 1; q.n_plain = q.total_tiles - tail; q.ksplit = S;
            q.ws = (float*)((char*)workspace + C2_WS_HEADER); q.ws_cnt = (int*)workspace;
        }
    }
    c2_set_fastdiv(q);
    const int items = q.tail_mode ? q.n_plain + (q.total_tiles - q.n_plain) * q.ksplit : q.total_tiles * q.ksplit;
    const int grid = items < sms ? items : sms;                       // persistent: one CTA per SM walks the work list
    cudaLaunchAttribute attr_pdl[1];
    attr_pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr_pdl[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(C2Cfg<TN, F16>::THREADS); cfg.dynamicSmemBytes = C2Cfg<TN, F16>::SMEM;
    cfg.stream = stream; cfg.attrs = attr_pdl; cfg.numAttrs = g_conv_pdl ? 1 : 0;
    if (q.tail_mode) {
        cudaLaunchKernelEx(&cfg, conv2_kernel<TN, 2, F16>, map, q);
    } else if (q.ksplit > 1) {
        cudaLaunchKernelEx(&cfg, conv2_kernel<TN, 1, F16>, map, q);
    } else {
        cudaLaunchKernelEx(&cfg, conv2_kernel<TN, 0, F16>, map, q);
    }
    return launch_status("aoc_conv2d_nhwc_tc");
}

// halo variant: no split-K (the launcher only takes it for layers with at least one tile per SM)
template <int TN>
static int launch_conv2_halo(const CUtensorMap& map, const Conv2P& p, int tiles, cudaStream_t stream) {
    using Cfg = C2Cfg<TN, true, true>;
    static_assert(Cfg::SMEM <= 232448, "shared memory of the halo variant");
    static PerDeviceOnce attr;
    if (attr.first())
        cudaFuncSetAttribute(conv2_kernel<TN, 0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    Conv2P q = p;
    q.tiles_n = cdiv(p.Cout, TN);
    q.total_tiles = tiles * q.tiles_n;
    q.ksplit = 1; q.ws = nullptr; q.ws_cnt = nullptr;
    c2_set_fastdiv(q);
    const int sms = device_sms();
    const int grid = q.total_tiles < sms ? q.total_tiles : sms;
    cudaLaunchAttribute attr_pdl[1];
    attr_pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr_pdl[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = stream; cfg.attrs = attr_pdl; cfg.numAttrs = g_conv_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, conv2_kernel<TN, 0, true, true>, map, q);
    return launch_status("aoc_conv2d_nhwc_tc (halo)");
}

int g_conv_halo = 1;    // aoc_set_option("conv_halo", 0/1): halo variant for the 3x3 / stride-1 layers that fill the chip
int g_conv_chunk = 8;   // aoc_set_option("conv_chunk", stages): default accumulation chain length
int g_conv_dbg = 0;     // aoc_set_option("conv_dbg", bits): ablation switches, honoured by the tooling build only (C2_DBG)
unsigned long long* g_conv_trace = nullptr;

}  // namespace aoc

using namespace aoc;

extern "C" size_t aoc_conv_packed_weight_bytes(int Cout, int Cin, int kh, int kw, int operand_mode) {
    const size_t ncc = (size_t)cdiv(Cin, C2_KC);
    return (size_t)cdiv(Cout, C2_WRB) * kh * kw * ncc * c2_wchunk(operand_mode == AOC_CONV_SPLIT_F16);
}

extern "C" int aoc_conv_pack_weights(const float* w, int Cout, int Cin, int kh, int kw, int operand_mode, void* w_packed,
                                     cudaStream_t stream) {
    AOC_CHECK_ARG(w && w_packed && Cout > 0 && Cin > 0 && kh > 0 && kw > 0, "bad args");
    AOC_CHECK_ARG(operand_mode == AOC_CONV_SPLIT_F16 || operand_mode == AOC_CONV_TF32X3, "unknown operand mode");
    const int g_conv_f16 = operand_mode == AOC_CONV_SPLIT_F16;
    const int ncc = cdiv(Cin, C2_KC);
    const int rows_padded = cdiv(Cout, C2_WRB) * C2_WRB;
    const long long total = (long long)rows_padded * kh * kw * ncc * (g_conv_f16 ? 2 : 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (g_conv_f16)
        conv2_pack_weights_f16_kernel<<<blocks, 256, 0, stream>>>(w, Cout, Cin, kh * kw, ncc, rows_padded, (uint8_t*)w_packed);
    else
        conv2_pack_weights_kernel<<<blocks, 256, 0, stream>>>(w, Cout, Cin, kh * kw, ncc, rows_padded, (uint8_t*)w_packed);
    return launch_status("aoc_conv_pack_weights");
}

// does this layer run the halo variant?  3x3, stride 1, pad = dilation <= C2H_D, split-fp16 operands, and at least one
// 16 x 8 tile per SM (smaller layers keep the per-tap kernel and its split-K schedule)
static bool conv_use_halo(int N, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int dil, int operand_mode) {
    (void)Cin;
    if (!g_conv_halo || operand_mode != AOC_CONV_SPLIT_F16) return false;
    if (kh != 3 || kw != 3 || stride != 1 || dil < 1 || dil > C2H_D || pad != dil) return false;
    const long long tiles = (long long)N * cdiv(W, C2H_TW) * cdiv(H, C2H_TH) * cdiv(Cout, Cout <= 64 ? 64 : 128);
    return tiles >= device_sms();
}

// pixel-patch geometry shared by the launcher and the tile-statistics consumers
static void conv_geometry(int H, int W, int kh, int kw, int stride, int pad, int dil, int* gH, int* gW, int* Ho, int* Wo,
                          int* tw_log2) {
    *Ho = (H + 2 * pad - dil * (kh - 1) - 1) / stride + 1;
    *Wo = (W + 2 * pad - dil * (kw - 1) - 1) / stride + 1;
    *gW = W; *gH = H;
    // 1x1/s1/p0 convolutions see each image as one row of H*W pixels
    if (kh == 1 && kw == 1 && stride == 1 && pad == 0) { *gW = H * W; *gH = 1; *Wo = *gW; *Ho = 1; }
    int best_l2 = 7;
    long long best_tiles = -1;
    for (int l2 = 7; l2 >= 3; --l2) {
        int tw = 1 << l2, th = C2_BM >> l2;
        long long t = (long long)cdiv(*Wo, tw) * cdiv(*Ho, th);
        if (best_tiles < 0 || t < best_tiles) { best_tiles = t; best_l2 = l2; }
    }
    *tw_log2 = best_l2;
}

extern "C" size_t aoc_conv_workspace_bytes(int N, int H, int W, int Cout, int kh, int kw, int stride, int pad, int dil) {
    int gH, gW, Ho, Wo, l2;
    conv_geometry(H, W, kh, kw, stride, pad, dil, &gH, &gW, &Ho, &Wo, &l2);
    if (Ho <= 0 || Wo <= 0) return 0;
    // split-K of a small map: up to C2_MAX_KSPLIT partial images; tail splitting of a large one: at most one parked
    // accumulator tile (128 x 128 floats) per SM
    const size_t full = (size_t)C2_MAX_KSPLIT * N * Ho * Wo * Cout * sizeof(float);
    const size_t tail = (size_t)device_sms() * C2_BM * 128 * sizeof(float);
    const long long M = (long long)N * Ho * Wo;
    return C2_WS_HEADER + (M <= 128 * 74 ? (full > tail ? full : tail) : tail);
}

extern "C" int aoc_conv_trace(void* device_buffer_16x256_u64) {
    g_conv_trace = (unsigned long long*)device_buffer_16x256_u64;
    return AOC_OK;
}

extern "C" int aoc_conv_tiles_per_image(int N, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int dil,
                                        int operand_mode) {
    int gH, gW, Ho, Wo, l2;
    conv_geometry(H, W, kh, kw, stride, pad, dil, &gH, &gW, &Ho, &Wo, &l2);
    if (Ho <= 0 || Wo <= 0) return 0;
    if (conv_use_halo(N, H, W, Cin, Cout, kh, kw, stride, pad, dil, operand_mode)) return 4 * cdiv(Wo, C2H_TW) * cdiv(Ho, C2H_TH);
    return 4 * cdiv(Wo, 1 << l2) * cdiv(Ho, C2_BM >> l2);     // one statistics row per 32-pixel quadrant of a 128-pixel tile
}

extern "C" int aoc_conv2d_nhwc_tc(const float* x, const void* w_packed, const float* bias, const float* residual,
                                  const float* in_a, const float* in_b, int in_relu, float* y, float* tile_stats, int N,
                                  int H, int W, int Cin, int ldx, int Cout, int ldy, int ldres, int kh, int kw,
                                  int stride, int pad, int dil, int relu, int chunk_stages, int operand_mode,
                                  int* overflow_flag, void* workspace, size_t ws_bytes, cudaStream_t stream) {
    AOC_CHECK_ARG(x && w_packed && y, "null pointer");
    AOC_CHECK_ARG(operand_mode == AOC_CONV_SPLIT_F16 || operand_mode == AOC_CONV_TF32X3, "unknown operand mode");
    const int g_conv_f16 = operand_mode == AOC_CONV_SPLIT_F16;
    AOC_CHECK_ARG(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && kh > 0 && kw > 0 && dil > 0, "bad dims");
    AOC_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
    AOC_CHECK_ARG(ldx % 4 == 0 && (((uintptr_t)x) & 15) == 0, "ldx must be a multiple of 4 and x 16-byte aligned");
    const bool affine = in_a || in_b || in_relu;
    AOC_CHECK_ARG(!affine || Cin <= C2_MAX_AFFINE_C, "fused input affine supports Cin <= 1024");
    EncodeTiledFn encode = get_encode();
    if (!encode) {
        set_error("aoc_conv2d_nhwc_tc: cuTensorMapEncodeTiled is unavailable (no CUDA driver)");
        return AOC_ELAUNCH;
    }
    int gH, gW, Ho, Wo, best_l2;
    conv_geometry(H, W, kh, kw, stride, pad, dil, &gH, &gW, &Ho, &Wo, &best_l2);
    AOC_CHECK_ARG(Ho > 0 && Wo > 0, "empty output");
    AOC_CHECK_ARG((long long)N * Ho * Wo * (ldy > Cout ? ldy : Cout) < (1ll << 31), "output larger than 2^31 elements");
    Conv2P p;
    p.w = (const uint8_t*)w_packed; p.bias = bias; p.res = residual; p.in_a = in_a; p.in_b = in_b; p.y = y;
    p.tile_stats = tile_stats;
    p.N = N; p.Cin = Cin; p.Cout = Cout; p.ldy = ldy; p.ldres = ldres; p.kw = kw; p.stride = stride; p.pad = pad;
    p.dil = dil; p.relu = relu; p.in_relu = in_relu;
    p.ncc = cdiv(Cin, C2_KC);
    p.nIt = kh * kw * p.ncc;
    p.chunk = chunk_stages > 0 ? chunk_stages : g_conv_chunk;
    p.taps = kh * kw;
    p.trace = g_conv_trace;
    p.dbg = g_conv_dbg;
    p.overflow = overflow_flag;
    p.vec_out = (ldy % 4 == 0) && (((uintptr_t)y & 15) == 0) && (!residual || (ldres % 4 == 0 && ((uintptr_t)residual & 15) == 0));
    p.H = gH; p.W = gW; p.Ho = Ho; p.Wo = Wo;
    const bool halo = conv_use_halo(N, H, W, Cin, Cout, kh, kw, stride, pad, dil, operand_mode);
    if (halo) best_l2 = 3;                                            // 16 rows x 8 columns
    p.tw_log2 = best_l2;
    p.th = C2_BM >> best_l2;
    const int tw = 1 << best_l2;
    p.tiles_x = cdiv(Wo, tw);
    p.tiles_y = cdiv(Ho, p.th);
    CUtensorMap map;
    cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)gW, (cuuint64_t)gH, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)gW * ldx * 4, (cuuint64_t)gH * gW * ldx * 4};
    cuuint32_t box[4] = {(cuuint32_t)C2_RKC, (cuuint32_t)(tw * stride), (cuuint32_t)(p.th * stride), 1u};
    if (halo) { box[1] = (cuuint32_t)(C2H_TW + 2 * dil); box[2] = (cuuint32_t)(C2H_TH + 2 * dil); }
    cuuint32_t estr[4] = {1u, (cuuint32_t)stride, (cuuint32_t)stride, 1u};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("aoc_conv2d_nhwc_tc: cuTensorMapEncodeTiled failed (%d) dims %d x %d x %d x %d ld %d box %d x %d", (int)cr, Cin,
                  gW, gH, N, ldx, tw * stride, p.th * stride);
        return AOC_EINVAL;
    }
    const int tiles = N * p.tiles_x * p.tiles_y;
    // A K = 8 TF32 MMA occupies the tensor pipe ~74 cycles whatever N <= 128, and every output-channel tile repeats the
    // operand transform of its pixels, so the 128-wide tile is used whenever Cout > 64 (its 64 accumulators per drain
    // thread fit since the warpgroups re-partition the register file).  Layers with few pixel tiles are spread over the
    // SMs by split-K, not by narrower tiles.
    const bool narrow = Cout <= 64 || p.nIt < g_conv_narrow_nit;
    if (halo) return Cout <= 64 ? launch_conv2_halo<64>(map, p, tiles, stream) : launch_conv2_halo<128>(map, p, tiles, stream);
    if (g_conv_f16)
        return narrow ? launch_conv2<64, true>(map, p, tiles, workspace, ws_bytes, stream)
                      : launch_conv2<128, true>(map, p, tiles, workspace, ws_bytes, stream);
    return narrow ? launch_conv2<64, false>(map, p, tiles, workspace, ws_bytes, stream)
                  : launch_conv2<128, false>(map, p, tiles, workspace, ws_bytes, stream);
}

// Microbenchmark: what does ONE thread (and two threads of different warps) pay per tcgen05.mma / tcgen05.commit it issues?
// Build here (nvcc cross-compiles), run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I aocb200/csrc -o tools/microbench/umma_issue tools/microbench/umma_issue.cu
// Every variant: `iters` iterations of a warp-uniform loop in which one elected lane issues the instructions, clock64 around
// the loop, then a commit + wait so that the tensor pipe has drained.  Operands are whatever the shared memory holds.
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace aoc::umma;

struct Res { long long cyc[32]; };

template <int V>
__device__ __forceinline__ void body(uint32_t tm, uint32_t a, uint32_t b, uint32_t bar, uint32_t idesc_n[4], int i) {
    // V encodes the instruction mix of one iteration
    const uint64_t ad = smem_desc(a, 128, 256), bd = smem_desc(b, 128, 256);
    if (V == 1) { mma_f16(tm, ad, bd, idesc_n[3], 1); }                                    // N = 128
    if (V == 2) { mma_f16(tm, ad, bd, idesc_n[2], 1); }                                    // N = 64
    if (V == 3) { mma_f16(tm, ad, bd, idesc_n[0], 1); }                                    // N = 16
    if (V == 4) { mma_commit(bar); }
    if (V == 5) { mma_f16(tm, ad, bd, idesc_n[0], 1); mma_commit(bar); }
    if (V == 6) { mma_f16(tm, ad, bd, idesc_n[0], 1); mma_f16(tm + 16, ad, bd, idesc_n[0], 1); mma_commit(bar); }
    if (V == 7) { mma_f16(tm, ad, bd, idesc_n[3], 1); if (i & 1) mma_commit(bar); }        // our MAIN issuer's mix
    if (V == 8) { mma_f16(tm, ad, bd, idesc_n[3], 1); mma_f16(tm + 128, ad, bd, idesc_n[3], 1); mma_f16(tm + 256, ad, bd, idesc_n[3], 1); }
    if (V == 9) { mma_f16(tm, ad, bd, idesc_n[3], 1); mma_f16(tm + 128, ad, bd, idesc_n[3], 1); mma_f16(tm + 256, ad, bd, idesc_n[3], 1);
                  if (i & 1) mma_commit(bar); }
    if (V == 10) { mma_f16(tm, ad, bd, idesc_n[1], 1); }                                   // N = 32
    if (V == 11) { mma_f16(tm, ad, bd, idesc_n[3] , 1); mma_f16(tm, ad, bd, idesc_n[3], 1); }   // two dependent N=128 (same accumulator)
    // A operand in the halo layout of the convolution: 8-row groups 160 B apart (one halo row of 10 pixels), k-groups 2880 B apart
    if (V == 12) { mma_f16(tm, smem_desc(a, 2880, 160), bd, idesc_n[3], 1); }                         // tap (0, 0): 128 B aligned start
    if (V == 13) { mma_f16(tm, smem_desc(a + 16 * ((i % 9) / 3 * 10 + (i % 9) % 3), 2880, 160), bd, idesc_n[3], 1); }   // the nine taps in turn
    if (V == 14) { mma_f16(tm, smem_desc(a, 2048, 128), bd, idesc_n[3], 1); }                         // control: groups 128 B apart (dense, aligned)
    if (V == 15) { mma_f16(tm, smem_desc(a + 16, 2048, 128), bd, idesc_n[3], 1); }                    // dense but shifted by one 16 B row
    if (V == 16) { const uint64_t x = smem_desc(a + 16 * ((i % 9) / 3 * 10 + (i % 9) % 3), 2880, 160), y = smem_desc(a + 8192 + 16 * ((i % 9) / 3 * 10 + (i % 9) % 3), 2880, 160);
                   const uint64_t bl = smem_desc(b + 4096, 128, 256);
                   mma_f16(tm, x, bd, idesc_n[3], 1); mma_f16(tm + 128, y, bd, idesc_n[3], 1); mma_f16(tm + 128, x, bl, idesc_n[3], 1); }   // a halo stage
}

template <int V>
__device__ long long run(uint32_t tm, uint32_t a, uint32_t b, uint32_t bar, uint32_t done_bar, uint32_t& done_phase, int iters, bool elect_each) {
    uint32_t idn[4] = {idesc_f16(128, 16), idesc_f16(128, 32), idesc_f16(128, 64), idesc_f16(128, 128)};
    __syncwarp();
    const long long t0 = clock64();
    if (elect_each) {
        for (int i = 0; i < iters; ++i) {
            tc_fence_after();
            if (elect_one()) body<V>(tm, a, b, bar, idn, i);
            __syncwarp();
        }
    } else {
        if (elect_one()) for (int i = 0; i < iters; ++i) body<V>(tm, a, b, bar, idn, i);
        __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) mma_commit(done_bar);
    __syncwarp();
    mbar_wait(done_bar, done_phase);
    done_phase ^= 1u;
    const long long t2 = clock64();
    (void)t2;
    return t1 - t0;
}

__global__ void __launch_bounds__(128, 1) k(Res* out, int iters, int two_warps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
    if (warp == 1 || (two_warps && warp == 2)) {
        const int w = warp - 1;
        const uint32_t bar = smem_u32(&bars[w]), done = smem_u32(&bars[2 + w]);
        uint32_t ph = 0;
        const uint32_t tmw = tm + (w ? 384u : 0u);
        long long r[24];
        int n = 0;
#define RUN(V, E) r[n++] = run<V>(tmw, a + w * 8192, b + w * 8192, bar, done, ph, iters, E)
        RUN(1, false); RUN(2, false); RUN(10, false); RUN(3, false); RUN(4, false); RUN(5, false); RUN(6, false); RUN(7, false);
        RUN(8, false); RUN(9, false); RUN(11, false);
        RUN(12, false); RUN(13, false); RUN(14, false); RUN(15, false); RUN(16, false);
        RUN(1, true); RUN(3, true); RUN(4, true); RUN(7, true); RUN(9, true);
        if (lane == 0 && blockIdx.x == 0) for (int i = 0; i < n; ++i) out[w].cyc[i] = r[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    const char* names[] = {"1 MMA N=128", "1 MMA N=64", "1 MMA N=32", "1 MMA N=16", "1 commit", "1 MMA N=16 + commit", "2 MMA N=16 + commit",
                           "1 MMA N=128 + commit/2", "3 MMA N=128", "3 MMA N=128 + commit/2", "2 MMA N=128 same acc", "A halo layout (SBO 160), tap (0,0)", "A halo layout, nine taps in turn", "A dense groups (SBO 128)",
                           "A dense, start + 16 B", "halo stage: 3 MMAs (hi*hi, lo*hi, hi*lo)",
                           "[fence+elect+sync each] 1 MMA N=128", "[each] 1 MMA N=16", "[each] 1 commit", "[each] MMA N=128 + commit/2",
                           "[each] 3 MMA N=128 + commit/2"};
    Res* d;
    cudaMalloc(&d, 2 * sizeof(Res));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 2000;
    for (int two = 0; two < 1; ++two) {
        for (int grid : {1, 148}) {
            cudaMemset(d, 0, 2 * sizeof(Res));
            k<<<grid, 128, 65536>>>(d, iters, two);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            Res h[2];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("== %s issuing warp(s), grid %d: cycles per iteration\n", two ? "TWO" : "one", grid);
            for (int i = 0; i < 21; ++i) {
                printf("  %-42s %8.1f", names[i], (double)h[0].cyc[i] / iters);
                if (two) printf("   (second warp %8.1f)", (double)h[1].cyc[i] / iters);
                printf("\n");
            }
        }
    }
    return 0;
}

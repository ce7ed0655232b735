// Shared device/host helpers for libaocb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/aocb200.h"

#define AOC_WRONG_LABEL_PAD 5.0e4f  // reference: networks/layers/matching.py:25

namespace aoc {

void set_error(const char* fmt, ...);

inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return AOC_ELAUNCH;
    }
    return AOC_OK;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// First statement of the small kernels that run between two convolutions (coefficient / statistics / resize glue):
// "my dependents may be scheduled".  The kernel itself is launched normally (it starts after its predecessor has completed);
// the NEXT kernel of the stream -- a convolution launched with the programmatic-stream-serialization attribute -- then
// becomes resident while this one runs and does everything that does not depend on it (barrier init, TMEM allocation,
// tensor-map prefetch, the first ring of weight stages) before its griddepcontrol.wait.  The dependent grid is only
// launched once EVERY block of this grid has executed the trigger (or exited), so it can never hold resources a block of
// this grid still needs.  Without such a dependent the instruction does nothing.
#ifndef AOC_NO_GLUE_TRIGGER      // (tooling build for the A/B measurement: AOCB200_NVCC_FLAGS=-DAOC_NO_GLUE_TRIGGER)
#define AOC_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;")
#else
#define AOC_PDL_TRIGGER() do { } while (0)
#endif
// ... and for a glue kernel that is ITSELF launched as a programmatic dependent (launch_pdl below) of the convolution in
// front of it: everything after this line sees the previous grid's writes.  A no-op under a normal launch.
#define AOC_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")

// <<<grid, block, smem, stream>>> with the programmatic-stream-serialization attribute: the grid may be scheduled as soon as
// every block of the previous kernel of the stream has triggered (the convolution does so at its start) or exited, i.e. its
// blocks start on an SM the moment the previous kernel's CTA there retires instead of after a full launch round trip behind
// the completed grid.  The kernel MUST execute AOC_PDL_WAIT() before it touches global memory.
extern int g_glue_pdl;                       // aoc_set_option("glue_pdl", 0 / 1)
template <typename... P, typename... A>
inline void launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A... args) {
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cfg.attrs = attr; cfg.numAttrs = g_glue_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

// Function attributes (opt-in dynamic shared memory) and the SM count belong to a DEVICE, not to the process: a host
// that drives engines on several GPUs from one process must set / query them once per device ordinal.
constexpr int AOC_MAX_DEVICES = 64;
struct PerDeviceOnce {
    bool done[AOC_MAX_DEVICES] = {};
    // true exactly once per current device (and always for ordinals beyond the table: the attribute call is cheap)
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= AOC_MAX_DEVICES) return true;
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};
inline int device_sms() {
    static int sms[AOC_MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const bool tab = dev >= 0 && dev < AOC_MAX_DEVICES;
    if (tab && sms[dev] > 0) return sms[dev];
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    if (tab) sms[dev] = n;
    return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 2*sigmoid(x)-1 exactly as the reference composes it: (sigmoid(x) - 0.5) * 2   (matching.py:2508)
__device__ __forceinline__ float sig2(float x) {
    float s = 1.0f / (1.0f + expf(-x));
    return (s - 0.5f) * 2.0f;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace aoc

#define AOC_CHECK_ARG(cond, msg)                         \
    do {                                                 \
        if (!(cond)) {                                   \
            aoc::set_error("%s: %s", __func__, msg);     \
            return AOC_EINVAL;                           \
        }                                                \
    } while (0)

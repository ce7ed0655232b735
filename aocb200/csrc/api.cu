// Error reporting, version and device check of libaocb200.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace aoc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int g_glue_pdl = 0;      // measured slower: the coefficient kernels stay normal launches
}  // namespace aoc

namespace aoc { extern int g_conv_chunk; extern int g_conv_dbg; extern int g_conv_splitk; extern int g_conv_narrow_nit; extern int g_conv_pdl; extern int g_match_f16; extern int g_conv_halo; extern int g_match_fast; extern int g_match_collector; extern int g_conv_tail; extern int g_conv_tail_min_stages; }

extern "C" int aoc_version(void) { return 200; }

extern "C" int aoc_set_option(const char* key, int value) {
    if (key && !strcmp(key, "conv_chunk") && value > 0) { aoc::g_conv_chunk = value; return AOC_OK; }
    if (key && !strcmp(key, "conv_dbg") && value >= 0) { aoc::g_conv_dbg = value; return AOC_OK; }
    if (key && !strcmp(key, "conv_splitk")) { aoc::g_conv_splitk = value != 0; return AOC_OK; }
    if (key && !strcmp(key, "match_fast")) { aoc::g_match_fast = value != 0; return AOC_OK; }
    if (key && !strcmp(key, "match_collector")) { aoc::g_match_collector = value != 0; return AOC_OK; }
    if (key && !strcmp(key, "match_f16")) { aoc::g_match_f16 = value != 0; return AOC_OK; }
    if (key && !strcmp(key, "conv_tail_min_stages") && value >= 0) { aoc::g_conv_tail_min_stages = value; return AOC_OK; }
    if (key && !strcmp(key, "conv_tail")) { aoc::g_conv_tail = value != 0; return AOC_OK; }
    if (key && !strcmp(key, "conv_halo")) { aoc::g_conv_halo = value != 0; return AOC_OK; }
    if (key && !strcmp(key, "conv_pdl")) { aoc::g_conv_pdl = value != 0; return AOC_OK; }
    if (key && !strcmp(key, "glue_pdl")) { aoc::g_glue_pdl = value != 0; return AOC_OK; }
    if (key && !strcmp(key, "conv_narrow_nit") && value >= 0) { aoc::g_conv_narrow_nit = value; return AOC_OK; }
    aoc::set_error("aoc_set_option: unknown option '%s'", key ? key : "(null)");
    return AOC_EINVAL;
}

extern "C" const char* aoc_last_error_string(void) { return aoc::g_err; }

extern "C" int aoc_check_device(int dev) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
        aoc::set_error("aoc_check_device: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return AOC_ELAUNCH;
    }
    if (prop.major != 10) {
        aoc::set_error("aoc_check_device: device %d is sm_%d%d, this library is built for sm_100a only", dev,
                       prop.major, prop.minor);
        return AOC_EARCH;
    }
    return AOC_OK;
}

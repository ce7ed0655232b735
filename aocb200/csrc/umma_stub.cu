// temporary stubs (replaced by umma_*.cu)
#include "common.cuh"
extern "C" size_t aoc_conv_packed_weight_bytes(int, int) { return 0; }
extern "C" int aoc_conv_pack_weights_tf32x3(const float*, int, int, void*, cudaStream_t) { aoc::set_error("not built"); return AOC_EINVAL; }
extern "C" int aoc_conv2d_nhwc_tc(const float*, const void*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t) { aoc::set_error("not built"); return AOC_EINVAL; }
extern "C" size_t aoc_bank_tc_bytes(int) { return 0; }
extern "C" int aoc_bank_gather_tc(const float*, const int*, int, void*, float*, cudaStream_t) { aoc::set_error("not built"); return AOC_EINVAL; }
extern "C" int aoc_global_match_tc(const float*, int, const void*, const float*, const int*, const int*, const float*, int, void*, float*, float*, cudaStream_t) { aoc::set_error("not built"); return AOC_EINVAL; }
extern "C" size_t aoc_global_match_tc_workspace_bytes(int) { return 0; }
extern "C" int aoc_gemm_tf32x3_test(const float*, const float*, float*, int, int, int, int, cudaStream_t) { aoc::set_error("not built"); return AOC_EINVAL; }

// tc_conv.cu — tcgen05 implicit-GEMM convolutions (placeholder until the kernels land; returns UNSUPPORTED so
// conv.cu routes to the CUDA-core implicit-GEMM path).
#include "tc_common.cuh"
int agb_tc_conv_fprop(agb_ctx*, int, const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int) { return AGB_ERR_UNSUPPORTED; }
int agb_tc_conv_wgrad(agb_ctx*, int, const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int) { return AGB_ERR_UNSUPPORTED; }

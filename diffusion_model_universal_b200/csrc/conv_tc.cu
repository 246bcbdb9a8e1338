// tcgen05 / TMA implicit-GEMM convolution (bf16 operands, fp32 TMEM accumulators).
#include "common.cuh"

using namespace dmu;

extern "C" {
int dmu_conv2d_tc_supported(const dmu_conv_params*) { return 0; }
int dmu_conv2d_tc(const dmu_conv_params*, dmu_stream_t) { return fail("dmu_conv2d_tc: not built"); }
int dmu_wgrad_tc_supported(const dmu_wgrad_params*) { return 0; }
int dmu_wgrad_tc(const dmu_wgrad_params*, dmu_stream_t) { return fail("dmu_wgrad_tc: not built"); }
}

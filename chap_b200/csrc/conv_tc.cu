// tcgen05 / TMEM / TMA implicit-GEMM convolution path (placeholder until the kernel lands).
#include "common.cuh"
#include "conv_plan.cuh"
namespace chap {
bool tc_supports(const Geom&, bool) { return false; }
int tc_conv(const Geom&, bool, const float*, const float*, const float*, float*, double*, cudaStream_t) { return 0; }
}

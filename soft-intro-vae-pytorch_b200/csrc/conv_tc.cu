// tcgen05 TF32 implicit-GEMM convolution -- placeholder until the kernels land (everything routes to SIMT).
#include "kernels.h"
namespace sivae {
bool conv_tc_supported_fwd(const ConvShape&) { return false; }
int launch_conv_fwd_tc(const float*, const float*, const float*, const float*, float*, const ConvShape&, cudaStream_t) { return -100; }
bool conv_tc_supported_wgrad(const ConvShape&) { return false; }
size_t conv_wgrad_tc_scratch_bytes(const ConvShape&) { return 0; }
int launch_conv_wgrad_tc(const float*, const float*, float*, const ConvShape&, bool, void*, size_t, cudaStream_t) { return -100; }
}

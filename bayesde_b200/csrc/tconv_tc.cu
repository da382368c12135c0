// K2 (tensor-core mode): tcgen05 implicit-GEMM tiles.  Placeholder until the UMMA kernel lands:
// reports "unsupported" so callers use the fp32 CUDA-core kernels.
#include "common.cuh"

namespace bsde {
bool tconv_tc_supported(const bsde_conv_layer*, int) { return false; }
int tconv_fwd_tc(const bsde_conv_layer*, int, int, const float*, const float*, long long, float, int, int,
                 float, const float*, float*, cudaStream_t) {
  set_error("tensor-core conv not built");
  return BSDE_ERR_UNSUPPORTED;
}
}  // namespace bsde

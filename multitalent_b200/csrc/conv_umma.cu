// tcgen05 (UMMA) + TMA implementation of the tap-table convolution -- placeholder until the kernel lands.
#include "common.cuh"

namespace mtb {

int umma_available() { return 0; }

int conv_taps_umma(const mtb200_conv_params&, cudaStream_t) {
  set_error("conv_taps(umma): not built");
  return MTB200_ERR_UNSUPPORTED;
}

int wgrad_taps_umma(const mtb200_wgrad_params&, cudaStream_t) {
  set_error("wgrad_taps(umma): not built");
  return MTB200_ERR_UNSUPPORTED;
}

}  // namespace mtb

// tcgen05 engine placeholder (filled in by the 3xTF32 UMMA kernel).
#include "kernels.cuh"
namespace vpb {
bool conv_tc_supported(const ConvParams&) { return false; }
cudaError_t launch_conv_tc(const ConvParams&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace vpb

// tcgen05 tensor-core implicit-GEMM engine (B200LIC_ENGINE_TC).  Placeholder until the kernel lands:
// every shape is reported as unsupported so AUTO falls through to the SIMT engine.
#include "common.cuh"

namespace b200lic {
int tc_conv_fwd(const b200lic_conv_desc*, const float*, const float*, const float*, const float*, float*, float*,
                cudaStream_t) {
  set_error("conv_fwd: tensor-core engine not built");
  return B200LIC_ERR_UNSUPPORTED;
}
}  // namespace b200lic

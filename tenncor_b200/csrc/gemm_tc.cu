// gemm_tc.cu — tcgen05 (5th-gen tensor core) fp32 GEMM: TF32 and 3xTF32, TMEM accumulators.
// Placeholder dispatch until the UMMA kernel lands: reports "not handled" so tcr_gemm
// runs the exact SIMT kernel (still on the device).
#include "common.cuh"

namespace tcr {
int gemm_tc_dispatch(const void*, const void*, void*, const tcr_gemm_desc*, bool* handled) {
  *handled = false;
  return TCR_OK;
}
}  // namespace tcr

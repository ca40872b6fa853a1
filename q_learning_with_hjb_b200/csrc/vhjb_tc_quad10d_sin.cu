// tensor-core vhjb kernel instantiations (sin value nets) for one system; see vhjb_tc.cuh.
#include "vhjb_tc.cuh"

namespace hjb {
cudaError_t vhjb_tc_launch_quad10d_sin(const VhjbArgs& a, const VhjbLaunch& l, int uform, int rform, cudaStream_t st) {
  using S = Quad10DSys<false>;
  if (uform == HJB_U_CLIPPED && rform == HJB_RES_NORMALIZED)
    return tc::launch_vhjb_tc_variant<S, HJB_ACT_SIN, HJB_U_CLIPPED, HJB_RES_NORMALIZED>(a, l, st);
  return cudaErrorNotSupported;
}
}  // namespace hjb

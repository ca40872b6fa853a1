// tensor-core vhjb kernel instantiations (relu value nets) for one system; see vhjb_tc.cuh.
#include "vhjb_tc.cuh"

namespace hjb {
cudaError_t vhjb_tc_launch_quad2d_relu(const VhjbArgs& a, const VhjbLaunch& l, int uform, int rform, cudaStream_t st) {
  using S = Quad2DSys<false>;
  if (uform == HJB_U_CLIPPED && rform == HJB_RES_NORMALIZED)
    return tc::launch_vhjb_tc_variant<S, HJB_ACT_RELU, HJB_U_CLIPPED, HJB_RES_NORMALIZED>(a, l, st);
  return cudaErrorNotSupported;
}
cudaError_t vhjb_tc_launch_quad2d_tanh(const VhjbArgs& a, const VhjbLaunch& l, int uform, int rform, cudaStream_t st);
cudaError_t vhjb_tc_launch_quad2d_sin(const VhjbArgs& a, const VhjbLaunch& l, int uform, int rform, cudaStream_t st);
cudaError_t vhjb_tc_launch_quad2d(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st) {
  switch (act) {
    case HJB_ACT_RELU: return vhjb_tc_launch_quad2d_relu(a, l, uform, rform, st);
    case HJB_ACT_TANH: return vhjb_tc_launch_quad2d_tanh(a, l, uform, rform, st);
    case HJB_ACT_SIN: return vhjb_tc_launch_quad2d_sin(a, l, uform, rform, st);
    default: return cudaErrorNotSupported;
  }
}
}  // namespace hjb

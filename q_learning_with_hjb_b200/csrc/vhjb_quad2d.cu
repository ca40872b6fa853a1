// vhjb kernel instantiations for one system; see vhjb_simt.cuh.
#include "vhjb_simt.cuh"

namespace hjb {
cudaError_t vhjb_launch_quad2d(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st) {
  return launch_vhjb_system<Quad2DSys<false>, false>(a, l, act, uform, rform, st);
}
}  // namespace hjb

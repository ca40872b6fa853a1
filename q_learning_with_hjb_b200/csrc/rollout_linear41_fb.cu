// Rollout kernel instantiations for one (system, controller) pair; see rollout_kernel.cuh.
#include "rollout_kernel.cuh"

namespace hjb {
#define SYS_LINEAR21(F) LinearSys<2, 1, F>
#define SYS_LINEAR41(F) LinearSys<4, 1, F>
#define SYS_LINEAR42(F) LinearSys<4, 2, F>
#define SYS_CARTPOLE(F) CartpoleSys<F>
#define SYS_ACROBOT(F) AcrobotSys<F>
#define SYS_QUAD2D(F) Quad2DSys<F>
#define SYS_QUAD10D(F) Quad10DSys<F>
HJB_DEFINE_FB_PROBLEM(linear41_fb, SYS_LINEAR41, true)
}  // namespace hjb

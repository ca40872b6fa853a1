// Rollout kernel instantiations for the double integrator under the bang-bang controllers of the minimum-time comparison
// (HJB_CTL_SWITCH_CURVE, HJB_CTL_GRID_SIGN); see mintime_ctl.cuh and rollout_kernel.cuh.
#include "mintime_ctl.cuh"

namespace hjb {
#define SYS_LINEAR21(F) LinearSys<2, 1, F>
HJB_DEFINE_PROBLEM(linear21_switch, SYS_LINEAR21, SwitchCurveCtl, true)
HJB_DEFINE_PROBLEM(linear21_grid, SYS_LINEAR21, GridSignCtl, true)
}  // namespace hjb

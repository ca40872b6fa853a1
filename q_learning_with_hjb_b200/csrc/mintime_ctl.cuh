// Bang-bang controllers of the double integrator's minimum-time comparison ("the level-set / analytic
// comparison", examples/double_integrator_optimal_time.ipynb cells 18-21), as controllers of the rollout kernel — the
// notebook steps ten trajectories through them in Python; here every environment of a launch runs them in the step loop.
//   SwitchCurveCtl  get_analytical_control (cell 18): 0 inside |x|^2 <= metric, else +a below / -a above the switching
//                   curve p = -v |v| / 2 (with the notebook's <= / < asymmetry on the two branches)
//   GridSignCtl     get_level_set_control (cell 18): u = -a sign(T[iv][ip]), T = dV/dvel on a regular (vel, pos) grid
//                   (central differences of the level-set solver's value function), looked up at the NEAREST node with
//                   scipy's RegularGridInterpolator(method="nearest", bounds_error=False, fill_value=None) convention:
//                   node = ceil(t - 1/2) of the fractional index t (ties go down), clamped to the grid (extrapolation)
// n = 2, m = 1 (x = [pos, vel]); the system's own clip (Dynamics.simulate, dynamics_basic.py:118) stays in the loop.
#pragma once
#include "rollout_kernel.cuh"

namespace hjb {

// aux = {metric, amplitude}
struct SwitchCurveCtl {
  static constexpr int KIND = HJB_CTL_SWITCH_CURVE;
  static constexpr bool kClips = false;
  template <class S>
  static __device__ __forceinline__ void control(const DevSys&, const DevCtl& pc, const float* x, const typename S::Trig&,
                                                 float* u, int = 0) {
    static_assert(S::N == 2 && S::M == 1, "the switching curve is the double integrator's");
    const float p = x[0], v = x[1];
    const float h = 0.5f * v * v;
    const bool plus = (v < 0.f && p <= h) || (v >= 0.f && p < -h);
    const float r = fmaf(p, p, v * v);
    // a NaN state fails every comparison: u = -a, the state stays NaN (the reference's branch structure does the same)
    u[0] = r <= pc.aux[0] ? 0.f : (plus ? pc.aux[1] : -pc.aux[1]);
  }
};

// ref = table [ref_steps = nv][ref_offset = np]; aux = {pos_min, 1 / dpos, vel_min, 1 / dvel, amplitude}
struct GridSignCtl {
  static constexpr int KIND = HJB_CTL_GRID_SIGN;
  static constexpr bool kClips = false;
  static __device__ __forceinline__ int node(float x, float lo, float inv_h, int n) {
    const float t = ceilf(fmaf(x - lo, inv_h, -0.5f));
    // clamp in float first: a far-away state must not overflow the conversion; NaN -> node 0 (u from a valid entry; the
    // state itself stays NaN through the dynamics)
    return (int)fminf(fmaxf(t, 0.f), (float)(n - 1));
  }
  template <class S>
  static __device__ __forceinline__ void control(const DevSys&, const DevCtl& pc, const float* x, const typename S::Trig&,
                                                 float* u, int = 0) {
    static_assert(S::N == 2 && S::M == 1, "the grid policy is defined on (pos, vel)");
    const int ip = node(x[0], pc.aux[0], pc.aux[1], pc.ref_offset);
    const int iv = node(x[1], pc.aux[2], pc.aux[3], pc.ref_steps);
    const float g = __ldg(pc.ref + (int64_t)iv * pc.ref_offset + ip);
    u[0] = g > 0.f ? -pc.aux[4] : (g < 0.f ? pc.aux[4] : 0.f);
  }
};

HJB_DECLARE_PROBLEM(linear21_switch);
HJB_DECLARE_PROBLEM(linear21_grid);

// first time a recorded trajectory is inside the goal ball (cell 20's loop: `if x_{k+1}^T x_{k+1} <= metric:
// t = min(k dt, t)`); defined in mintime.cu
cudaError_t first_hit(const float* xs, int64_t N, int32_t n, int32_t rows, float metric, float dt, float t_max, float* t_hit,
                      cudaStream_t st);

}  // namespace hjb

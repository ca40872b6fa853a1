// Batched single-step pieces of the rollout path (hjb_dynamics / hjb_control_efforts / hjb_states_wrap):
// the same device functions the fused rollout kernel uses, exposed per state for the class methods
// (Dynamics.get_control_affine_matrix / dynamics_step / simulate, Controller.get_control_efforts) and for
// the per-step parity tests.
#include "rollout_kernel.cuh"
#include "mintime_ctl.cuh"
#include "step_kernels.cuh"

namespace hjb {

// One explicit Euler step in the reference's own order, x + xdot dt, then states_wrap (dynamics_basic.py:120-121).  The per-step
// entry points (Dynamics.simulate, the learned-policy step) keep this order; the whole-horizon rollout kernel folds dt into the
// constants of the two quadrotor systems (systems.cuh::euler: fewer instructions, the same step to a few ulps).  Data sets
// grown by learned-policy rollouts therefore do not depend on that optimisation — a 200-epoch on-policy training is chaotic in
// the last bit of its samples.
template <class S>
__device__ __forceinline__ void euler_step_plain(const DevSys& ps, float* x, const typename S::Trig& tr, const float* u) {
  float d[S::N];
  S::xdot(ps, x, tr, u, d);
#pragma unroll
  for (int i = 0; i < S::N; ++i) x[i] = fmaf(d[i], ps.dt, x[i]);
  wrap_state<S>(x);
}

template <class S, int INTEG>
__global__ void __launch_bounds__(256) dynamics_kernel(const __grid_constant__ DynArgs a) {
  constexpr int N = S::N, M = S::M;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B) return;
  float x[N], u[M];
  {
    float xr[N];
    load_row<N>(a.x, i, xr);
    to_internal<S>(a.sys, xr, x);   // aoff = 0 here: only wraps the angles (f, g are periodic in them)
  }
#pragma unroll
  for (int k = 0; k < M; ++k) u[k] = 0.f;
  if (a.u) load_row<M>(a.u, i, u);
  typename S::Trig tr;
  S::trig(a.sys, x, tr);
  if (a.f || a.g) {
    float f[N], g[N * M];
    S::fg(a.sys, x, tr, f, g);
    if (a.f) store_row<N>(a.f, i, f);
    if (a.g) store_row<N * M>(a.g, i, g);
  }
  if (a.xdot) {
    float d[N];
    S::xdot(a.sys, x, tr, u, d);   // dynamics_step does NOT clip (dynamics_basic.py:96-105)
    store_row<N>(a.xdot, i, d);
  }
  if (a.x_next) {
    clip_u<S>(a.sys, u);           // simulate clips (dynamics_basic.py:118)
    if constexpr (INTEG == HJB_INT_EULER) euler_step_plain<S>(a.sys, x, tr, u);
    else integrate<S, INTEG>(a.sys, x, tr, u, DirectTrig<S::kFast>{});
    float xo[N];
    to_external<S>(a.sys, x, xo);
    store_row<N>(a.x_next, i, xo);
  }
}

template <class S>
static cudaError_t launch_dyn(const DynArgs& a, int integ, cudaStream_t st) {
  const int block = 256;
  const unsigned grid = (unsigned)((a.B + block - 1) / block);
  switch (integ) {
    case HJB_INT_EULER: dynamics_kernel<S, HJB_INT_EULER><<<grid, block, 0, st>>>(a); break;
    case HJB_INT_RK4: dynamics_kernel<S, HJB_INT_RK4><<<grid, block, 0, st>>>(a); break;
    case HJB_INT_DISCRETE:
      if constexpr (S::KIND == HJB_SYS_LINEAR) {
        dynamics_kernel<S, HJB_INT_DISCRETE><<<grid, block, 0, st>>>(a);
        break;
      } else {
        return cudaErrorNotSupported;
      }
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <template <bool> class SysT>
static cudaError_t launch_dyn_fast(const DynArgs& a, int integ, bool fast, cudaStream_t st) {
  return fast ? launch_dyn<SysT<true>>(a, integ, st) : launch_dyn<SysT<false>>(a, integ, st);
}

template <bool F> using Lin21 = LinearSys<2, 1, F>;
template <bool F> using Lin41 = LinearSys<4, 1, F>;
template <bool F> using Lin42 = LinearSys<4, 2, F>;

cudaError_t step_dynamics(int kind, const DynArgs& a, int integ, bool fast, cudaStream_t st) {
  switch (kind) {
    case HJB_SYS_LINEAR:
      if (a.sys.n == 2 && a.sys.m == 1) return launch_dyn_fast<Lin21>(a, integ, fast, st);
      if (a.sys.n == 4 && a.sys.m == 1) return launch_dyn_fast<Lin41>(a, integ, fast, st);
      if (a.sys.n == 4 && a.sys.m == 2) return launch_dyn_fast<Lin42>(a, integ, fast, st);
      return cudaErrorNotSupported;
    case HJB_SYS_CARTPOLE: return launch_dyn_fast<CartpoleSys>(a, integ, fast, st);
    case HJB_SYS_ACROBOT: return launch_dyn_fast<AcrobotSys>(a, integ, fast, st);
    case HJB_SYS_QUAD2D: return launch_dyn_fast<Quad2DSys>(a, integ, fast, st);
    case HJB_SYS_QUAD10D: return launch_dyn_fast<Quad10DSys>(a, integ, fast, st);
    default: return cudaErrorNotSupported;
  }
}

// ---------------------------------------------------------------------------------------------
template <class S, class C>
__global__ void __launch_bounds__(256) control_kernel(const __grid_constant__ CtlArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B) return;
  float x[S::N], u[S::M];
  {
    float xr[S::N];
    load_row<S::N>(a.x, i, xr);
    to_internal<S>(a.sys, xr, x);
  }
  typename S::Trig tr;
  S::trig(a.sys, x, tr);
  C::template control<S>(a.sys, a.ctl, x, tr, u);
  store_row<S::M>(a.u, i, u);
}

template <class S, class C>
static cudaError_t launch_ctl(const CtlArgs& a, cudaStream_t st) {
  const int block = 256;
  const unsigned grid = (unsigned)((a.B + block - 1) / block);
  control_kernel<S, C><<<grid, block, 0, st>>>(a);
  return cudaGetLastError();
}
template <template <bool> class SysT, class C>
static cudaError_t launch_ctl_fast(const CtlArgs& a, bool fast, cudaStream_t st) {
  return fast ? launch_ctl<SysT<true>, C>(a, st) : launch_ctl<SysT<false>, C>(a, st);
}

template <class FB>
static cudaError_t step_control_fb(int sys_kind, const CtlArgs& a, bool fast, cudaStream_t st);

cudaError_t step_control(int sys_kind, int ctl_kind, const CtlArgs& a, bool fast, cudaStream_t st) {
  if (ctl_kind == HJB_CTL_FEEDBACK)
    return a.ctl.clip ? step_control_fb<FeedbackCtl<true>>(sys_kind, a, fast, st)
                      : step_control_fb<FeedbackCtl<false>>(sys_kind, a, fast, st);
  if (ctl_kind == HJB_CTL_CARTPOLE_ES && sys_kind == HJB_SYS_CARTPOLE)
    return launch_ctl_fast<CartpoleSys, CartpoleESCtl>(a, fast, st);
  if (ctl_kind == HJB_CTL_ACROBOT_ES && sys_kind == HJB_SYS_ACROBOT)
    return launch_ctl_fast<AcrobotSys, AcrobotESCtl>(a, fast, st);
  if (ctl_kind == HJB_CTL_TRACK && sys_kind == HJB_SYS_QUAD2D) return launch_ctl_fast<Quad2DSys, TrackCtl>(a, fast, st);
  if (sys_kind == HJB_SYS_LINEAR && a.sys.n == 2 && a.sys.m == 1) {
    if (ctl_kind == HJB_CTL_SWITCH_CURVE) return launch_ctl_fast<Lin21, SwitchCurveCtl>(a, fast, st);
    if (ctl_kind == HJB_CTL_GRID_SIGN) return launch_ctl_fast<Lin21, GridSignCtl>(a, fast, st);
  }
  return cudaErrorNotSupported;
}

template <class FeedbackCtl>
static cudaError_t step_control_fb(int sys_kind, const CtlArgs& a, bool fast, cudaStream_t st) {
  {
    switch (sys_kind) {
      case HJB_SYS_LINEAR:
        if (a.sys.n == 2 && a.sys.m == 1) return launch_ctl_fast<Lin21, FeedbackCtl>(a, fast, st);
        if (a.sys.n == 4 && a.sys.m == 1) return launch_ctl_fast<Lin41, FeedbackCtl>(a, fast, st);
        if (a.sys.n == 4 && a.sys.m == 2) return launch_ctl_fast<Lin42, FeedbackCtl>(a, fast, st);
        return cudaErrorNotSupported;
      case HJB_SYS_CARTPOLE: return launch_ctl_fast<CartpoleSys, FeedbackCtl>(a, fast, st);
      case HJB_SYS_ACROBOT: return launch_ctl_fast<AcrobotSys, FeedbackCtl>(a, fast, st);
      case HJB_SYS_QUAD2D: return launch_ctl_fast<Quad2DSys, FeedbackCtl>(a, fast, st);
      case HJB_SYS_QUAD10D: return launch_ctl_fast<Quad10DSys, FeedbackCtl>(a, fast, st);
      default: return cudaErrorNotSupported;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Learned-policy rollout, one step of every trajectory (VHJBController.rollout_trajectory, controller/vhjb.py:171-193):
//   dx = wrap(x - xf); outside the observation box (:176-177) or `terminal` (:188-191): the trajectory ends with the
//   sample (x, dx^T P dx, done = 1) (:167-169, :178-181); otherwise the sample is (x, l(x, u) dt, done = 0) with
//   l = dx^T Q dx + (u - uf)^T R (u - uf) (:162-165, :184) and x <- simulate(x, u) (dynamics_basic.py:107-122).
// u is the policy's control at x (hjb_vhjb_residual); trajectories that already ended produce no sample.
template <class S>
__global__ void __launch_bounds__(256) policy_step_kernel(const __grid_constant__ PolicyStepArgs a) {
  constexpr int N = S::N, M = S::M;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  if (a.alive[i] == 0.f) {
    if (a.rec_done) a.rec_done[i] = -1.f;
    return;
  }
  float xr[N], z[N];
  load_row<N>(a.x, i, xr);
#pragma unroll
  for (int k = 0; k < N; ++k) z[k] = xr[k] - a.xf[k];
  wrap_state<S>(z);
  bool outside = a.terminal != 0;
#pragma unroll
  for (int k = 0; k < N; ++k) outside = outside || (z[k] > a.hi[k]) || (z[k] < a.lo[k]);
  float cost = 0.f;
  if (outside) {
#pragma unroll
    for (int r = 0; r < N; ++r) {
      float row = 0.f;
#pragma unroll
      for (int c = 0; c < N; ++c) row = fmaf(a.P[r * N + c], z[c], row);
      cost = fmaf(z[r], row, cost);
    }
    a.alive[i] = 0.f;
  } else {
    float u[M], du[M];
    load_row<M>(a.u, i, u);
#pragma unroll
    for (int r = 0; r < N; ++r) {
      float row = 0.f;
#pragma unroll
      for (int c = 0; c < N; ++c) row = fmaf(a.Q[r * N + c], z[c], row);
      cost = fmaf(z[r], row, cost);
    }
#pragma unroll
    for (int k = 0; k < M; ++k) du[k] = u[k] - a.uf[k];
#pragma unroll
    for (int r = 0; r < M; ++r) {
      float row = 0.f;
#pragma unroll
      for (int c = 0; c < M; ++c) row = fmaf(a.R[r * M + c], du[c], row);
      cost = fmaf(du[r], row, cost);
    }
    cost *= a.sys.dt;
    float x[N];
    to_internal<S>(a.sys, xr, x);
    typename S::Trig tr;
    S::trig(a.sys, x, tr);
    clip_u<S>(a.sys, u);
    euler_step_plain<S>(a.sys, x, tr, u);
    float xo[N];
    to_external<S>(a.sys, x, xo);
    store_row<N>(a.x, i, xo);
  }
  a.total_cost[i] += cost;
  if (a.rec_x) store_row<N>(a.rec_x, i, xr);
  if (a.rec_cost) a.rec_cost[i] = cost;
  if (a.rec_done) a.rec_done[i] = outside ? 1.f : 0.f;
}

template <class S>
static cudaError_t launch_policy(const PolicyStepArgs& a, cudaStream_t st) {
  policy_step_kernel<S><<<(unsigned)((a.N + 255) / 256), 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t step_policy(int kind, const PolicyStepArgs& a, cudaStream_t st) {
  switch (kind) {
    case HJB_SYS_LINEAR:
      if (a.sys.n == 2 && a.sys.m == 1) return launch_policy<LinearSys<2, 1, false>>(a, st);
      return cudaErrorNotSupported;
    case HJB_SYS_CARTPOLE: return launch_policy<CartpoleSys<false>>(a, st);
    case HJB_SYS_QUAD2D: return launch_policy<Quad2DSys<false>>(a, st);
    case HJB_SYS_QUAD10D: return launch_policy<Quad10DSys<false>>(a, st);
    default: return cudaErrorNotSupported;
  }
}

// ---------------------------------------------------------------------------------------------
struct WrapArgs {
  float* x;
  int64_t B;
};
template <class S>
__global__ void __launch_bounds__(256) wrap_kernel(WrapArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B) return;
  float x[S::N];
  load_row<S::N>(a.x, i, x);
  wrap_state<S>(x);
  store_row<S::N>(a.x, i, x);
}

cudaError_t step_wrap(int kind, int n, float* x, int64_t B, cudaStream_t st) {
  const int block = 256;
  const unsigned grid = (unsigned)((B + block - 1) / block);
  WrapArgs a{x, B};
  switch (kind) {
    case HJB_SYS_LINEAR: return cudaSuccess;  // identity (dynamics/linear.py:17-18)
    case HJB_SYS_CARTPOLE: wrap_kernel<CartpoleSys<false>><<<grid, block, 0, st>>>(a); break;
    case HJB_SYS_ACROBOT: wrap_kernel<AcrobotSys<false>><<<grid, block, 0, st>>>(a); break;
    case HJB_SYS_QUAD2D: wrap_kernel<Quad2DSys<false>><<<grid, block, 0, st>>>(a); break;
    case HJB_SYS_QUAD10D: wrap_kernel<Quad10DSys<false>><<<grid, block, 0, st>>>(a); break;
    default: return cudaErrorNotSupported;
  }
  (void)n;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// FP32 FMA peak probe: 8 independent FMA chains per thread, `iters` rounds -> 16*iters flops/thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fma_probe_kernel(float* sink, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
  }
  sink[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

cudaError_t fma_probe(float* sink, int64_t sink_len, int iters, double* flops, cudaStream_t st) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  int64_t blocks = (int64_t)sms * 8;  // 8 x 256 threads = 64 warps per SM
  if (blocks * 256 > sink_len) blocks = sink_len / 256;
  if (blocks <= 0) return cudaErrorInvalidValue;
  fma_probe_kernel<<<(unsigned)blocks, 256, 0, st>>>(sink, iters);
  if (flops) *flops = (double)blocks * 256.0 * (double)iters * 8.0 * 8.0 * 2.0;
  return cudaGetLastError();
}

}  // namespace hjb

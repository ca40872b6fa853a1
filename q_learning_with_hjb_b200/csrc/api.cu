// extern "C" entry points of libhjb_b200.so (see include/hjb_b200.h): argument checking, folding of the
// public parameter structs into the device parameter blocks (derived constants in double precision), and
// dispatch to the kernel instantiations.
#include <cmath>
#include <cstring>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "rollout_kernel.cuh"
#include "mintime_ctl.cuh"
#include "step_kernels.cuh"

namespace hjb {

static bool dims_ok(const hjb_system* s) {
  switch (s->kind) {
    case HJB_SYS_LINEAR: return (s->n == 2 && s->m == 1) || (s->n == 4 && (s->m == 1 || s->m == 2));
    case HJB_SYS_CARTPOLE: return s->n == 4 && s->m == 1;
    case HJB_SYS_ACROBOT: return s->n == 4 && s->m == 1;
    case HJB_SYS_QUAD2D: return s->n == 6 && s->m == 2;
    case HJB_SYS_QUAD10D: return s->n == 10 && s->m == 3;
    default: return false;
  }
}

// indices of the angle components per system kind (states_wrap)
static int angle_indices(int kind, int* idx) {
  switch (kind) {
    case HJB_SYS_CARTPOLE: idx[0] = 1; return 1;
    case HJB_SYS_ACROBOT: idx[0] = 0; idx[1] = 1; return 2;
    case HJB_SYS_QUAD2D: idx[0] = 2; return 1;
    case HJB_SYS_QUAD10D: idx[0] = 3; idx[1] = 4; return 2;
    default: return 0;
  }
}
static bool is_angle(int kind, int i) {
  int idx[2];
  const int na = angle_indices(kind, idx);
  for (int k = 0; k < na; ++k)
    if (idx[k] == i) return true;
  return false;
}
static double wrap_pi_host(double a) {
  const double two_pi = 6.283185307179586476925286766559;
  return a - two_pi * std::floor((a + 0.5 * two_pi) / two_pi);
}

void make_dev_sys(const hjb_system* s, DevSys& d) {
  std::memset(&d, 0, sizeof(d));
  d.n = s->n;
  d.m = s->m;
  d.dt = s->dt;
  for (int k = 0; k < HJB_MAX_M; ++k) { d.umin[k] = s->umin[k]; d.umax[k] = s->umax[k]; }
  const float* p = s->par;
  switch (s->kind) {
    case HJB_SYS_LINEAR:
      std::memcpy(d.A, s->A, sizeof(d.A));
      std::memcpy(d.B, s->B, sizeof(d.B));
      break;
    case HJB_SYS_CARTPOLE: {  // par = {mc, mp, l, g}
      const double mc = p[0], mp = p[1], l = p[2], g = p[3];
      d.c[0] = (float)(mc + mp);
      d.c[1] = (float)(mp * l);
      d.c[2] = (float)(mp * l * l);
      d.c[3] = (float)(mp * g * l);
      d.c[4] = (float)((mc + mp) * (mp * l * l));
      d.c[5] = (float)(1.0 / l);
      d.c[6] = (float)(g / l);
      break;
    }
    case HJB_SYS_ACROBOT: {  // par = {l1, l2, m1, m2, I1, I2, g}
      const double l1 = p[0], l2 = p[1], m1 = p[2], m2 = p[3], I1 = p[4], I2 = p[5], g = p[6];
      d.c[0] = (float)(I1 + I2 + m2 * l1 * l1);
      d.c[1] = (float)(m2 * l1 * l2 / 2);
      d.c[2] = (float)I2;
      d.c[3] = (float)((m1 * l1 / 2 + m2 * l1) * g);
      d.c[4] = (float)(m2 * g * l2 / 2);
      break;
    }
    case HJB_SYS_QUAD2D: {  // par = {g, m, r, I}
      const double dt = s->dt;
      d.c[0] = p[0];
      d.c[1] = (float)(1.0 / (double)p[1]);
      d.c[2] = (float)((double)p[2] / (double)p[3]);
      d.c[3] = (float)(dt / (double)p[1]);                       // the explicit Euler step with dt folded in (systems.cuh)
      d.c[4] = (float)((double)p[0] * dt);
      d.c[5] = (float)((double)p[2] * dt / (double)p[3]);
      break;
    }
    case HJB_SYS_QUAD10D: {  // par = {g, m, kT, n0}
      const double dt = s->dt;
      d.c[0] = p[0];
      d.c[1] = (float)((double)p[2] / (double)p[1]);
      d.c[2] = p[3];
      d.c[3] = (float)((double)p[0] * dt);
      d.c[4] = (float)((double)p[2] * dt / (double)p[1]);
      d.c[5] = (float)((double)p[3] * dt);
      break;
    }
  }
}

// Fills the controller block AND the internal-coordinate angle offsets of `ds` (FEEDBACK: the goal angles).
static void make_dev_ctl(const hjb_system* s, const hjb_control* c, DevSys& ds, DevCtl& d) {
  std::memset(&d, 0, sizeof(d));
  d.clip = c->clip;
  ds.aoff[0] = ds.aoff[1] = 0.f;
  if (c->kind == HJB_CTL_FEEDBACK) {
    int idx[2];
    const int na = angle_indices(s->kind, idx);
    for (int k = 0; k < na; ++k) ds.aoff[k] = c->xf[idx[k]];
    for (int k = 0; k < s->m; ++k) {
      double acc = c->uf[k];
      for (int i = 0; i < s->n; ++i)
        if (!is_angle(s->kind, i)) acc += (double)c->K[k * s->n + i] * (double)c->xf[i];
      d.u0[k] = (float)acc;
    }
  }
  std::memcpy(d.K, c->K, sizeof(d.K));
  std::memcpy(d.P, c->P, sizeof(d.P));
  if (c->kind == HJB_CTL_ACROBOT_ES) {
    // dx^T P dx is evaluated over the upper triangle: P_ii dx_i^2 + (P_ij + P_ji) dx_i dx_j, j > i (14 instead of 20
    // instructions; the same value for any P, symmetric or not)
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
        d.P[i * 4 + j] = j > i ? (float)((double)c->P[i * 4 + j] + (double)c->P[j * 4 + i]) : (j == i ? c->P[i * 4 + i] : 0.f);
  }
  std::memcpy(d.xf, c->xf, sizeof(d.xf));
  std::memcpy(d.uf, c->uf, sizeof(d.uf));
  std::memcpy(d.aux, c->aux, sizeof(d.aux));
  d.ref = (c->kind == HJB_CTL_TRACK || c->kind == HJB_CTL_GRID_SIGN) ? c->ref : nullptr;
  d.ref_steps = c->ref_steps;
  d.ref_offset = c->ref_offset;
  if (c->kind == HJB_CTL_CARTPOLE_ES) {
    // aux in = {Ke0, Ke1, Ke2, eps_energy, eps_state}; E(xf) = 0.5 dth_f^2 - cos(th_f)
    d.aux[4] = c->aux[4] * c->aux[4];
    d.aux[5] = (float)(0.5 * (double)c->xf[3] * (double)c->xf[3] - std::cos((double)c->xf[1]));
  } else if (c->kind == HJB_CTL_ACROBOT_ES) {
    // aux in = {Ks0, Ks1, Ks2, eps}; E(xf) per dynamics/acrobot.py:60-70 in double
    const float* p = s->par;
    const double l1 = p[0], l2 = p[1], m1 = p[2], m2 = p[3], I1 = p[4], I2 = p[5], g = p[6];
    const double q1 = c->xf[0], q2 = c->xf[1], dq1 = c->xf[2], dq2 = c->xf[3];
    const double a = m2 * l1 * l2 / 2, c1 = std::cos(q1), c2 = std::cos(q2);
    const double T1 = 0.5 * I1 * dq1 * dq1;
    const double T2 = 0.5 * (m2 * l1 * l1 + I2 + 2 * a * c2) * dq1 * dq1 + 0.5 * I2 * dq2 * dq2 + (I2 + a * c2) * dq1 * dq2;
    const double U = -m1 * g * l1 / 2 * c1 - m2 * g * (l1 * c1 + l2 / 2 * std::cos(q1 + q2));
    d.aux[4] = (float)(T1 + T2 + U);
  }
}

static void make_dev_cost(const hjb_system* s, const hjb_cost* c, const DevSys& ds, DevCost& d) {
  std::memset(&d, 0, sizeof(d));
  std::memcpy(d.Q, c->Q, sizeof(d.Q));
  std::memcpy(d.R, c->R, sizeof(d.R));
  std::memcpy(d.xf, c->xf, sizeof(d.xf));
  std::memcpy(d.uf, c->uf, sizeof(d.uf));
  int idx[2];
  const int na = angle_indices(s->kind, idx);
  for (int k = 0; k < na; ++k) d.dang[k] = (float)wrap_pi_host((double)ds.aoff[k] - (double)c->xf[idx[k]]);
  for (int i = 0; i < s->n; ++i) {
    const double q = c->Q[i * s->n + i];
    d.sq[i] = (float)std::sqrt(q > 0 ? q : 0.0);
    d.c0[i] = is_angle(s->kind, i) ? 0.f : (float)(-std::sqrt(q > 0 ? q : 0.0) * (double)c->xf[i]);
  }
  for (int k = 0; k < s->m; ++k) {
    const double r = c->R[k * s->m + k];
    d.sr[k] = (float)std::sqrt(r > 0 ? r : 0.0);
    d.r0[k] = (float)(-std::sqrt(r > 0 ? r : 0.0) * (double)c->uf[k]);
  }
}

static void make_dev_box(const hjb_system* s, const hjb_rollout_opts* o, const DevSys& ds, DevBox& d) {
  std::memset(&d, 0, sizeof(d));
  std::memcpy(d.xf, o->box_xf, sizeof(d.xf));
  std::memcpy(d.lo, o->box_lo, sizeof(d.lo));
  std::memcpy(d.hi, o->box_hi, sizeof(d.hi));
  int idx[2];
  const int na = angle_indices(s->kind, idx);
  for (int k = 0; k < na; ++k) d.dang[k] = (float)wrap_pi_host((double)ds.aoff[k] - (double)o->box_xf[idx[k]]);
}

// COST_DIAG needs diagonal Q, R with non-negative entries (it uses their square roots)
static int cost_mode(const hjb_system* s, const hjb_cost* c) {
  if (!c) return COST_NONE;
  for (int i = 0; i < s->n; ++i)
    if (c->Q[i * s->n + i] < 0.f) return COST_DENSE;
  for (int i = 0; i < s->m; ++i)
    if (c->R[i * s->m + i] < 0.f) return COST_DENSE;
  for (int i = 0; i < s->n; ++i)
    for (int j = 0; j < s->n; ++j)
      if (i != j && c->Q[i * s->n + j] != 0.f) return COST_DENSE;
  for (int i = 0; i < s->m; ++i)
    for (int j = 0; j < s->m; ++j)
      if (i != j && c->R[i * s->m + j] != 0.f) return COST_DENSE;
  bool unit = true;   // Q = I, R = I, goal with zero non-angle components: the cheaper COST_UNIT form
  for (int i = 0; i < s->n; ++i) unit = unit && c->Q[i * s->n + i] == 1.f && (is_angle(s->kind, i) || c->xf[i] == 0.f);
  for (int i = 0; i < s->m; ++i) unit = unit && c->R[i * s->m + i] == 1.f;
  return unit ? COST_UNIT : COST_DIAG;
}

static int to_status(cudaError_t e) { return e == cudaSuccess ? HJB_OK : (e == cudaErrorNotSupported ? HJB_ERR_UNSUPPORTED : (int)e); }

}  // namespace hjb

using namespace hjb;

extern "C" {

int hjb_abi_version(void) { return HJB_ABI_VERSION; }

const char* hjb_status_string(int status) {
  switch (status) {
    case HJB_OK: return "ok";
    case HJB_ERR_BAD_ARG: return "hjb: bad argument";
    case HJB_ERR_UNSUPPORTED: return "hjb: unsupported (system, controller, integrator, n, m) combination";
    case HJB_ERR_NO_DEVICE: return "hjb: no CUDA device";
    default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "hjb: unknown status";
  }
}

int hjb_rollout(const hjb_system* sys, const hjb_control* ctl, const hjb_cost* cost_spec, const hjb_rollout_opts* opts,
                const float* x0, int64_t N, int32_t T, float* xs, float* us, float* x_final, float* cost,
                int32_t* steps, void* stream) {
  if (!sys || !ctl || !opts || N < 0 || T < 0) return HJB_ERR_BAD_ARG;
  if (!dims_ok(sys)) return HJB_ERR_UNSUPPORTED;
  if (ctl->kind == HJB_CTL_TRACK && (!ctl->ref || ctl->ref_steps <= 0 || ctl->ref_offset < 0)) return HJB_ERR_BAD_ARG;
  if (ctl->kind == HJB_CTL_GRID_SIGN && (!ctl->ref || ctl->ref_steps <= 0 || ctl->ref_offset <= 0)) return HJB_ERR_BAD_ARG;
  if (N == 0) return HJB_OK;
  if (!x0) return HJB_ERR_BAD_ARG;
  if (cost && !cost_spec) return HJB_ERR_BAD_ARG;
  if ((xs || us) && opts->record_stride <= 0) return HJB_ERR_BAD_ARG;

  RolloutArgs a;
  std::memset(&a, 0, sizeof(a));
  make_dev_sys(sys, a.sys);
  make_dev_ctl(sys, ctl, a.sys, a.ctl);
  if (cost_spec) make_dev_cost(sys, cost_spec, a.sys, a.cost);
  if (opts->box_enabled) make_dev_box(sys, opts, a.sys, a.box);
  a.x0 = x0; a.xs = xs; a.us = us; a.x_final = x_final; a.cost_out = cost; a.steps_out = steps;
  a.N = N; a.T = T;
  a.stride = opts->record_stride > 0 ? opts->record_stride : 1;
  a.n_rec = T / a.stride;
  {  // TMA bulk stores of whole 32-row blocks need 16-byte aligned time slices (HJB_ROLLOUT_STORES=direct: A/B switch)
    const char* e = std::getenv("HJB_ROLLOUT_STORES");
    const bool direct = e && std::strcmp(e, "direct") == 0;
    const bool ok_x = !xs || (((uintptr_t)xs % 16 == 0) && ((N * sys->n) % 4 == 0));
    const bool ok_u = !us || (((uintptr_t)us % 16 == 0) && ((N * sys->m) % 4 == 0));
    a.staged = (!direct && ok_x && ok_u) ? 1 : 0;
  }

  RolloutVariant v;
  v.integrator = opts->integrator;
  v.rec = (xs || us);
  v.cost = cost ? cost_mode(sys, cost_spec) : COST_NONE;
  v.box = opts->box_enabled != 0;
  const bool fast = opts->fast_trig != 0;
  cudaStream_t st = (cudaStream_t)stream;

  cudaError_t e = cudaErrorNotSupported;
  switch (sys->kind) {
    case HJB_SYS_LINEAR:
      if (ctl->kind == HJB_CTL_SWITCH_CURVE || ctl->kind == HJB_CTL_GRID_SIGN) {   // the double integrator's bang-bang laws
        if (sys->n == 2 && sys->m == 1)
          e = ctl->kind == HJB_CTL_SWITCH_CURVE ? rollout_linear21_switch(a, v, fast, st) : rollout_linear21_grid(a, v, fast, st);
        break;
      }
      if (ctl->kind != HJB_CTL_FEEDBACK) break;
      if (sys->n == 2 && sys->m == 1) e = rollout_linear21_fb(a, v, fast, st);
      else if (sys->n == 4 && sys->m == 1) e = rollout_linear41_fb(a, v, fast, st);
      else if (sys->n == 4 && sys->m == 2) e = rollout_linear42_fb(a, v, fast, st);
      break;
    case HJB_SYS_CARTPOLE:
      if (ctl->kind == HJB_CTL_FEEDBACK) e = rollout_cartpole_fb(a, v, fast, st);
      else if (ctl->kind == HJB_CTL_CARTPOLE_ES) e = rollout_cartpole_es(a, v, fast, st);
      break;
    case HJB_SYS_ACROBOT:
      if (ctl->kind == HJB_CTL_FEEDBACK) e = rollout_acrobot_fb(a, v, fast, st);
      else if (ctl->kind == HJB_CTL_ACROBOT_ES) e = rollout_acrobot_es(a, v, fast, st);
      break;
    case HJB_SYS_QUAD2D:
      if (ctl->kind == HJB_CTL_FEEDBACK) e = rollout_quad2d_fb(a, v, fast, st);
      else if (ctl->kind == HJB_CTL_TRACK) e = rollout_quad2d_track(a, v, fast, st);
      break;
    case HJB_SYS_QUAD10D:
      if (ctl->kind == HJB_CTL_FEEDBACK) e = rollout_quad10d_fb(a, v, fast, st);
      break;
  }
  return to_status(e);
}

int hjb_rollout_variant(const hjb_system* sys, const hjb_control* ctl, const hjb_cost* cost_spec, const hjb_rollout_opts* opts,
                        int32_t recorded, int32_t out[6]) {
  if (!sys || !ctl || !opts || !out) return HJB_ERR_BAD_ARG;
  if (!dims_ok(sys)) return HJB_ERR_UNSUPPORTED;
  out[0] = opts->integrator;
  out[1] = recorded != 0;
  out[2] = cost_mode(sys, cost_spec);
  // the compiled combinations (rollout_kernel.cuh::launch_cost): a box runs the dense form, recorded runs have no DIAG form
  if (opts->box_enabled) out[2] = COST_DENSE;
  else if (recorded && out[2] == COST_DIAG) out[2] = COST_DENSE;
  out[3] = opts->box_enabled != 0;
  out[4] = opts->fast_trig != 0;
  out[5] = ctl->kind == HJB_CTL_FEEDBACK ? (ctl->clip != 0) : 1;
  return HJB_OK;
}

int hjb_dynamics(const hjb_system* sys, int32_t integrator, int32_t fast_trig, const float* x, const float* u, int64_t B,
                 float* f, float* g, float* xdot, float* x_next, void* stream) {
  if (!sys || B < 0) return HJB_ERR_BAD_ARG;
  if (!dims_ok(sys)) return HJB_ERR_UNSUPPORTED;
  if (B == 0) return HJB_OK;
  if (!x || ((xdot || x_next) && !u)) return HJB_ERR_BAD_ARG;
  DynArgs a;
  make_dev_sys(sys, a.sys);
  a.x = x; a.u = u; a.f = f; a.g = g; a.xdot = xdot; a.x_next = x_next; a.B = B;
  return to_status(step_dynamics(sys->kind, a, integrator, fast_trig != 0, (cudaStream_t)stream));
}

int hjb_control_efforts(const hjb_system* sys, const hjb_control* ctl, int32_t fast_trig, const float* x, int64_t B,
                        float* u, void* stream) {
  if (!sys || !ctl || B < 0) return HJB_ERR_BAD_ARG;
  if (!dims_ok(sys)) return HJB_ERR_UNSUPPORTED;
  if (B == 0) return HJB_OK;
  if (!x || !u) return HJB_ERR_BAD_ARG;
  if (ctl->kind == HJB_CTL_TRACK && (!ctl->ref || ctl->ref_steps <= 0 || ctl->ref_offset < 0)) return HJB_ERR_BAD_ARG;
  if (ctl->kind == HJB_CTL_GRID_SIGN && (!ctl->ref || ctl->ref_steps <= 0 || ctl->ref_offset <= 0)) return HJB_ERR_BAD_ARG;
  CtlArgs a;
  make_dev_sys(sys, a.sys);
  make_dev_ctl(sys, ctl, a.sys, a.ctl);
  a.x = x; a.u = u; a.B = B;
  return to_status(step_control(sys->kind, ctl->kind, a, fast_trig != 0, (cudaStream_t)stream));
}

int hjb_time_to_goal(const float* xs, int64_t N, int32_t n, int32_t rows, float metric, float dt, float t_max, float* t_hit,
                     void* stream) {
  if (N < 0 || n <= 0 || n > HJB_MAX_N || rows < 1) return HJB_ERR_BAD_ARG;
  if (N == 0) return HJB_OK;
  if (!xs || !t_hit) return HJB_ERR_BAD_ARG;
  return to_status(first_hit(xs, N, n, rows, metric, dt, t_max, t_hit, (cudaStream_t)stream));
}

int hjb_states_wrap(const hjb_system* sys, float* x, int64_t B, void* stream) {
  if (!sys || B < 0) return HJB_ERR_BAD_ARG;
  if (!dims_ok(sys)) return HJB_ERR_UNSUPPORTED;
  if (B == 0) return HJB_OK;
  if (!x) return HJB_ERR_BAD_ARG;
  return to_status(step_wrap(sys->kind, sys->n, x, B, (cudaStream_t)stream));
}

int hjb_policy_step(const hjb_system* sys, const hjb_task* task, const float* xf, const float* obs_lo, const float* obs_hi,
                    const float* P, int32_t terminal, float* x, const float* u, float* alive, float* total_cost, float* rec_x,
                    float* rec_cost, float* rec_done, int64_t N, void* stream) {
  if (!sys || !task || !xf || !obs_lo || !obs_hi || !P || N < 0) return HJB_ERR_BAD_ARG;
  if (!dims_ok(sys)) return HJB_ERR_UNSUPPORTED;
  if (N == 0) return HJB_OK;
  if (!x || !u || !alive || !total_cost) return HJB_ERR_BAD_ARG;
  PolicyStepArgs a;
  std::memset(&a, 0, sizeof(a));
  make_dev_sys(sys, a.sys);
  const int n = sys->n, m = sys->m;
  for (int i = 0; i < n; ++i) { a.xf[i] = xf[i]; a.lo[i] = obs_lo[i]; a.hi[i] = obs_hi[i]; }
  for (int i = 0; i < n * n; ++i) { a.Q[i] = task->Q[i]; a.P[i] = P[i]; }
  for (int i = 0; i < m * m; ++i) a.R[i] = task->R[i];
  for (int i = 0; i < m; ++i) a.uf[i] = task->uf[i];
  a.terminal = terminal;
  a.x = x; a.u = u; a.alive = alive; a.total_cost = total_cost;
  a.rec_x = rec_x; a.rec_cost = rec_cost; a.rec_done = rec_done; a.N = N;
  return to_status(step_policy(sys->kind, a, (cudaStream_t)stream));
}

int hjb_policy_rollout(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xf, const float* obs_lo,
                       const float* obs_hi, const float* P, int32_t T, float* x, float* u, const float* zeros, const float* ones,
                       float* alive, float* total_cost, float* rec_x, float* rec_cost, float* rec_done, int64_t N,
                       void* workspace, void* stream) {
  if (!sys || !net || !task || T < 0 || N < 0) return HJB_ERR_BAD_ARG;
  if (N == 0) return HJB_OK;
  if (!x || !u || !zeros || !ones || !alive || !total_cost || !rec_x || !rec_cost || !rec_done || !workspace) return HJB_ERR_BAD_ARG;
  const int64_t n = sys->n;
  for (int32_t i = 0; i <= T; ++i) {
    const int32_t terminal = i == T;   // controller/vhjb.py:188-191: what is still running ends with a boundary sample
    if (!terminal) {
      const int rc = hjb_vhjb_residual(sys, net, task, x, zeros, ones, N, nullptr, nullptr, u, nullptr, nullptr, workspace, stream);
      if (rc != HJB_OK) return rc;
    }
    const int rc = hjb_policy_step(sys, task, xf, obs_lo, obs_hi, P, terminal, x, u, alive, total_cost, rec_x + (int64_t)i * N * n,
                                   rec_cost + (int64_t)i * N, rec_done + (int64_t)i * N, N, stream);
    if (rc != HJB_OK) return rc;
  }
  return HJB_OK;
}

int hjb_fma_peak_probe(float* sink, int64_t sink_len, int32_t iters, double* flops, void* stream) {
  if (!sink || sink_len <= 0 || iters <= 0) return HJB_ERR_BAD_ARG;
  return to_status(fma_probe(sink, sink_len, iters, flops, (cudaStream_t)stream));
}

}  // extern "C"

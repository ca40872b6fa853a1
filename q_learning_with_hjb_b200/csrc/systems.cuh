// Control-affine systems x' = f(x) + g(x) u and the model-based controllers, as device functions.
//
// INTERNAL COORDINATES.  Kernels keep a state z in registers: z_i = x_i for the non-angle components and
// z_th = wrap(x_th - aoff_th) for the angle components, where aoff (DevSys::aoff) is the goal angle of a
// feedback controller (0 otherwise).  Everything downstream is periodic in the angles, so this is the same
// dynamical system, but (i) the controller's and the cost's error coordinate wrap(x - xf) IS z_th — no
// subtraction, no second wrap per step — and (ii) near the goal z_th is small, so fp32 keeps ~1e-10 absolute
// resolution where the raw angle (e.g. th ~ pi for the cart-pole) would be quantised at 2.4e-7.
// x is materialised (to_external) only when a state is written out.
//
// Each system S<FAST> provides
//   N, M, NANG, ang(k)   dimensions; indices of the angle components
//   Trig / trig(p, z, t) sin/cos of the state's angles, computed once per evaluation point and shared between
//                        the control law and the dynamics of the same state
//   xdot(p, z, tr, u, d) d = f + g u with g's structural zeros skipped (the fused fast path)
//   fg(p, z, tr, f, g)   explicit f [N] and g [N*M] (row-major) for the API / the vhjb pass
// Reference formulas: dynamics/{linear,cartpole,acrobot,quadrotors}.py through the manipulator form of
// dynamics/dynamics_basic.py:64-94, with the 2x2 M^-1 written in closed form.
#pragma once
#include <type_traits>

#include "hjb_common.cuh"

namespace hjb {

// How a kernel evaluates the trigonometry of a state.  angle(k, z, aoff, s, c) returns sin / cos of (z + aoff) for the
// system's k-th angle argument (the acrobot has a third: q1 + q2 with offset aoff0 + aoff1).
//   DirectTrig  in line: polynomial (FAST) or libdevice — the per-step kernels and the vhjb pass
//   TableTrig   the rollout kernels' shared-memory tables (hjb_common.cuh::sincos_tab; aoff baked into the table);
//               GUARD: arguments may leave the table's range (RK4 stage states are not wrapped) -> in-line fallback
template <bool FAST>
struct DirectTrig {
  static constexpr bool kSumByAddition = false;
  template <bool GUARD = false>
  __device__ __forceinline__ void angle(int, float z, float aoff, float& s, float& c) const { sincos_<FAST>(z + aoff, s, c); }
  template <bool GUARD = false>
  __device__ __forceinline__ float tangent(int, float z, float aoff) const { return tan_<FAST>(z + aoff); }
};
// WIDE: entries (S, C, -S/2, -C/2) — one instruction less per evaluation, twice the shared memory (rollout_kernel.cuh)
template <bool WIDE>
struct TableTrigT {
  // sin / cos of the acrobot's q1 + q2 from those of q1 and q2 by the addition theorems (4 FMA-pipe instructions instead of
  // a third 14-instruction table evaluation; the error of the sum of two ~7e-8 table values stays below 2e-7)
  static constexpr bool kSumByAddition = true;
  using Entry = std::conditional_t<WIDE, float4, float2>;
  const Entry* tab;   // [tables][kTrigSize], shared memory
  template <bool GUARD = false>
  __device__ __forceinline__ void angle(int k, float z, float aoff, float& s, float& c) const {
    if (GUARD && !(fabsf(z) <= kTrigRange)) sincos_poly(z + aoff, s, c);
    else if constexpr (WIDE) sincos_tab4(tab + k * kTrigSize, z, s, c);
    else sincos_tab(tab + k * kTrigSize, z, s, c);
  }
  template <bool GUARD = false>
  __device__ __forceinline__ float tangent(int k, float z, float aoff) const {
    float s, c;
    angle<GUARD>(k, z, aoff, s, c);
    return s * rcp_approx(c);
  }
};
// number of tables a system needs
template <class S>
constexpr int trig_tables() { return S::NANG; }

// states_wrap on the angle components (cartpole.py:52-64, acrobot.py:72-81, quadrotors.py:48-70,151-170)
template <class S>
__device__ __forceinline__ void wrap_state(float* z) {
#pragma unroll
  for (int k = 0; k < S::NANG; ++k) z[S::ang(k)] = wrap_pi_<S::kFast>(z[S::ang(k)]);
}
template <class S>
__device__ __forceinline__ void to_internal(const DevSys& p, const float* x, float* z) {
#pragma unroll
  for (int i = 0; i < S::N; ++i) z[i] = x[i];
#pragma unroll
  for (int k = 0; k < S::NANG; ++k) z[S::ang(k)] = wrap_pi_<S::kFast>(x[S::ang(k)] - p.aoff[k]);
}
// x_th = wrap(z_th + aoff); with aoff == 0 the state is returned bit-for-bit
template <class S>
__device__ __forceinline__ void to_external(const DevSys& p, const float* z, float* x) {
#pragma unroll
  for (int i = 0; i < S::N; ++i) x[i] = z[i];
#pragma unroll
  for (int k = 0; k < S::NANG; ++k)
    if (p.aoff[k] != 0.f) x[S::ang(k)] = wrap_pi_<S::kFast>(z[S::ang(k)] + p.aoff[k]);
}

// ------------------------------------------------------------------------------------------------
// LINEAR  (dynamics/linear.py:20-22)   f = A x, g = B, wrap = identity
// ------------------------------------------------------------------------------------------------
template <int N_, int M_, bool FAST>
struct LinearSys {
  static constexpr int N = N_, M = M_, NANG = 0;
  static constexpr bool kFast = FAST;
  static constexpr bool kFusedEuler = false;
  static constexpr int KIND = HJB_SYS_LINEAR;
  static __device__ __forceinline__ constexpr int ang(int) { return 0; }
  struct Trig {};
  template <bool GUARD = false, class TC = DirectTrig<FAST>>
  static __device__ __forceinline__ void trig(const DevSys&, const float*, Trig&, const TC& = TC{}) {}
  // A x + B u   (also the exact-ZOH update when A, B are the discretised matrices)
  static __device__ __forceinline__ void xdot(const DevSys& p, const float* x, const Trig&, const float* u, float* d) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) acc = fmaf(p.A[i * N + j], x[j], acc);
#pragma unroll
      for (int k = 0; k < M; ++k) acc = fmaf(p.B[i * M + k], u[k], acc);
      d[i] = acc;
    }
  }
  static __device__ __forceinline__ void fg(const DevSys& p, const float* x, const Trig&, float* f, float* g) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) acc = fmaf(p.A[i * N + j], x[j], acc);
      f[i] = acc;
#pragma unroll
      for (int k = 0; k < M; ++k) g[i * M + k] = p.B[i * M + k];
    }
  }
};

// ------------------------------------------------------------------------------------------------
// CARTPOLE  (dynamics/cartpole.py:19-64)   x = [p, th, dp, dth]
//   M = [[M11, k c],[k c, M22]], C dq = [-k dth^2 s, 0], G = [0, gamma s], B = [1, 0]
//   c = {M11, kappa, M22, gamma, M11*M22, 1/l, g/l}
// ------------------------------------------------------------------------------------------------
template <bool FAST>
struct CartpoleSys {
  static constexpr int N = 4, M = 1, NANG = 1;
  static constexpr bool kFast = FAST;
  static constexpr bool kFusedEuler = false;
  static constexpr int KIND = HJB_SYS_CARTPOLE;
  static __device__ __forceinline__ constexpr int ang(int) { return 1; }
  struct Trig { float s, c; };
  template <bool GUARD = false, class TC = DirectTrig<FAST>>
  static __device__ __forceinline__ void trig(const DevSys& p, const float* z, Trig& t, const TC& tc = TC{}) {
    tc.template angle<GUARD>(0, z[1], p.aoff[0], t.s, t.c);
  }
  static __device__ __forceinline__ void xdot(const DevSys& p, const float* x, const Trig& t, const float* u, float* d) {
    const float m12 = p.c[1] * t.c;
    const float inv = rcp_<FAST>(fmaf(-m12, m12, p.c[4]));
    const float h1 = fmaf(p.c[1] * t.s, x[3] * x[3], u[0]);   // kappa dth^2 s + u
    const float h2 = -p.c[3] * t.s;                            // -gamma s
    d[0] = x[2];
    d[1] = x[3];
    d[2] = fmaf(p.c[2], h1, -m12 * h2) * inv;
    d[3] = fmaf(p.c[0], h2, -m12 * h1) * inv;
  }
  static __device__ __forceinline__ void fg(const DevSys& p, const float* x, const Trig& t, float* f, float* g) {
    const float m12 = p.c[1] * t.c;
    const float inv = rcp_<FAST>(fmaf(-m12, m12, p.c[4]));
    const float h1 = p.c[1] * t.s * x[3] * x[3];
    const float h2 = -p.c[3] * t.s;
    f[0] = x[2];
    f[1] = x[3];
    f[2] = fmaf(p.c[2], h1, -m12 * h2) * inv;
    f[3] = fmaf(p.c[0], h2, -m12 * h1) * inv;
    g[0] = 0.f;
    g[1] = 0.f;
    g[2] = p.c[2] * inv;
    g[3] = -m12 * inv;
  }
};

// ------------------------------------------------------------------------------------------------
// ACROBOT  (dynamics/acrobot.py:39-81)   x = [q1, q2, dq1, dq2]
//   c = {M11_0, a, I2, G1c, G12c}
//   M = [[M11_0 + 2 a c2, I2 + a c2],[I2 + a c2, I2]]
//   C dq = [-a s2 dq2 (2 dq1 + dq2), a s2 dq1^2],  G = [G1c s1 + G12c s12, G12c s12],  B = [0, 1]
// ------------------------------------------------------------------------------------------------
template <bool FAST>
struct AcrobotSys {
  static constexpr int N = 4, M = 1, NANG = 2;
  static constexpr bool kFast = FAST;
  static constexpr bool kFusedEuler = false;
  static constexpr int KIND = HJB_SYS_ACROBOT;
  static __device__ __forceinline__ constexpr int ang(int k) { return k; }
  struct Trig { float s1, c1, s2, c2, s12, c12; };
  template <bool GUARD = false, class TC = DirectTrig<FAST>>
  static __device__ __forceinline__ void trig(const DevSys& p, const float* z, Trig& t, const TC& tc = TC{}) {
    tc.template angle<GUARD>(0, z[0], p.aoff[0], t.s1, t.c1);
    tc.template angle<GUARD>(1, z[1], p.aoff[1], t.s2, t.c2);
    if constexpr (TC::kSumByAddition) {
      t.s12 = fmaf(t.s1, t.c2, t.c1 * t.s2);
      t.c12 = fmaf(t.c1, t.c2, -t.s1 * t.s2);
    } else {
      tc.template angle<GUARD>(2, z[0] + z[1], p.aoff[0] + p.aoff[1], t.s12, t.c12);
    }
  }
  struct Terms { float m11, m12, m22, h1, h2; };  // h = C dq + G
  static __device__ __forceinline__ void terms(const DevSys& p, const float* x, const Trig& t, Terms& r) {
    const float ac2 = p.c[1] * t.c2;
    r.m11 = fmaf(2.f, ac2, p.c[0]);
    r.m12 = p.c[2] + ac2;
    r.m22 = p.c[2];
    const float as2 = p.c[1] * t.s2;
    const float g2 = p.c[4] * t.s12;
    r.h1 = fmaf(p.c[3], t.s1, g2) - as2 * x[3] * fmaf(2.f, x[2], x[3]);
    r.h2 = fmaf(as2 * x[2], x[2], g2);
  }
  static __device__ __forceinline__ void xdot(const DevSys& p, const float* x, const Trig& t, const float* u, float* d) {
    Terms r;
    terms(p, x, t, r);
    const float inv = rcp_<FAST>(fmaf(r.m11, r.m22, -r.m12 * r.m12));
    const float b1 = -r.h1;          // B u - h, B = [0, 1]
    const float b2 = u[0] - r.h2;
    d[0] = x[2];
    d[1] = x[3];
    d[2] = fmaf(r.m22, b1, -r.m12 * b2) * inv;
    d[3] = fmaf(r.m11, b2, -r.m12 * b1) * inv;
  }
  static __device__ __forceinline__ void fg(const DevSys& p, const float* x, const Trig& t, float* f, float* g) {
    Terms r;
    terms(p, x, t, r);
    const float inv = rcp_<FAST>(fmaf(r.m11, r.m22, -r.m12 * r.m12));
    f[0] = x[2];
    f[1] = x[3];
    f[2] = fmaf(-r.m22, r.h1, r.m12 * r.h2) * inv;
    f[3] = fmaf(-r.m11, r.h2, r.m12 * r.h1) * inv;
    g[0] = 0.f;
    g[1] = 0.f;
    g[2] = -r.m12 * inv;
    g[3] = r.m11 * inv;
  }
  // dynamics/acrobot.py:60-70
  static __device__ __forceinline__ float energy(const DevSys& p, const float* x, const Trig& t, const Terms& r) {
    float e = 0.5f * r.m11 * x[2] * x[2];
    e = fmaf(0.5f * r.m22 * x[3], x[3], e);
    e = fmaf(r.m12 * x[2], x[3], e);
    e = fmaf(-p.c[3], t.c1, e);
    e = fmaf(-p.c[4], t.c12, e);
    return e;
  }
};

// ------------------------------------------------------------------------------------------------
// QUAD2D  (dynamics/quadrotors.py:17-70)   x = [x, y, th, dx, dy, dth],  c = {g, 1/m, r/I}
// ------------------------------------------------------------------------------------------------
template <bool FAST>
struct Quad2DSys {
  static constexpr int N = 6, M = 2, NANG = 1;
  static constexpr bool kFast = FAST;
  static constexpr bool kFusedEuler = true;
  static constexpr int KIND = HJB_SYS_QUAD2D;
  static __device__ __forceinline__ constexpr int ang(int) { return 2; }
  struct Trig { float s, c; };
  template <bool GUARD = false, class TC = DirectTrig<FAST>>
  static __device__ __forceinline__ void trig(const DevSys& p, const float* z, Trig& t, const TC& tc = TC{}) {
    tc.template angle<GUARD>(0, z[2], p.aoff[0], t.s, t.c);
  }
  static __device__ __forceinline__ void xdot(const DevSys& p, const float* x, const Trig& t, const float* u, float* d) {
    const float sm = (u[0] + u[1]) * p.c[1];
    d[0] = x[3];
    d[1] = x[4];
    d[2] = x[5];
    d[3] = -t.s * sm;
    d[4] = fmaf(t.c, sm, -p.c[0]);
    d[5] = (u[0] - u[1]) * p.c[2];
  }
  // x <- x + dt xdot with dt folded into the constants (c[3] = dt/m, c[4] = g dt, c[5] = r dt/I): 10 FMA-pipe instructions
  // instead of the 12 of xdot + axpy (no separate S/m, a_x, a_z products)
  static __device__ __forceinline__ void euler(const DevSys& p, float* x, const Trig& t, const float* u) {
    const float sd = (u[0] + u[1]) * p.c[3];
    x[0] = fmaf(x[3], p.dt, x[0]);
    x[1] = fmaf(x[4], p.dt, x[1]);
    x[2] = fmaf(x[5], p.dt, x[2]);
    x[3] = fmaf(-t.s, sd, x[3]);
    x[4] = fmaf(t.c, sd, x[4] - p.c[4]);
    x[5] = fmaf(u[0] - u[1], p.c[5], x[5]);
  }
  static __device__ __forceinline__ void fg(const DevSys& p, const float* x, const Trig& t, float* f, float* g) {
    f[0] = x[3]; f[1] = x[4]; f[2] = x[5];
    f[3] = 0.f; f[4] = -p.c[0]; f[5] = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = 0.f;
    g[6] = g[7] = -t.s * p.c[1];
    g[8] = g[9] = t.c * p.c[1];
    g[10] = p.c[2];
    g[11] = -p.c[2];
  }
};

// ------------------------------------------------------------------------------------------------
// QUAD10D  (dynamics/quadrotors.py:118-170)  x = [p(3), thx, thy, v(3), wx, wy],  c = {g, kT/m, n0}
// ------------------------------------------------------------------------------------------------
template <bool FAST>
struct Quad10DSys {
  static constexpr int N = 10, M = 3, NANG = 2;
  static constexpr bool kFast = FAST;
  static constexpr bool kFusedEuler = true;
  static constexpr int KIND = HJB_SYS_QUAD10D;
  static __device__ __forceinline__ constexpr int ang(int k) { return 3 + k; }
  struct Trig { float tx, ty; };
  template <bool GUARD = false, class TC = DirectTrig<FAST>>
  static __device__ __forceinline__ void trig(const DevSys& p, const float* z, Trig& t, const TC& tc = TC{}) {
    t.tx = tc.template tangent<GUARD>(0, z[3], p.aoff[0]);
    t.ty = tc.template tangent<GUARD>(1, z[4], p.aoff[1]);
  }
  static __device__ __forceinline__ void xdot(const DevSys& p, const float* x, const Trig& t, const float* u, float* d) {
#pragma unroll
    for (int i = 0; i < 5; ++i) d[i] = x[5 + i];
    d[5] = p.c[0] * t.tx;
    d[6] = p.c[0] * t.ty;
    d[7] = fmaf(p.c[1], u[0], -p.c[0]);
    d[8] = p.c[2] * u[1];
    d[9] = p.c[2] * u[2];
  }
  // explicit Euler step with dt folded in (c[3] = g dt, c[4] = kT dt/m, c[5] = n0 dt): 11 instead of 15 instructions
  static __device__ __forceinline__ void euler(const DevSys& p, float* x, const Trig& t, const float* u) {
#pragma unroll
    for (int i = 0; i < 5; ++i) x[i] = fmaf(x[5 + i], p.dt, x[i]);
    x[5] = fmaf(t.tx, p.c[3], x[5]);
    x[6] = fmaf(t.ty, p.c[3], x[6]);
    x[7] = fmaf(u[0], p.c[4], x[7] - p.c[3]);
    x[8] = fmaf(u[1], p.c[5], x[8]);
    x[9] = fmaf(u[2], p.c[5], x[9]);
  }
  static __device__ __forceinline__ void fg(const DevSys& p, const float* x, const Trig& t, float* f, float* g) {
#pragma unroll
    for (int i = 0; i < 5; ++i) f[i] = x[5 + i];
    f[5] = p.c[0] * t.tx;
    f[6] = p.c[0] * t.ty;
    f[7] = -p.c[0];
    f[8] = 0.f;
    f[9] = 0.f;
#pragma unroll
    for (int i = 0; i < 30; ++i) g[i] = 0.f;
    g[7 * 3 + 0] = p.c[1];
    g[8 * 3 + 1] = p.c[2];
    g[9 * 3 + 2] = p.c[2];
  }
};

// ------------------------------------------------------------------------------------------------
// controllers: control(ps, pc, z, tr, u) writes the controller OUTPUT (get_control_efforts), i.e. before
// Dynamics.simulate's own clip.  z is the internal state.
// ------------------------------------------------------------------------------------------------
template <class S>
__device__ __forceinline__ void clip_u(const DevSys& ps, float* u) {
#pragma unroll
  for (int k = 0; k < S::M; ++k) u[k] = clampf(u[k], ps.umin[k], ps.umax[k]);
}

// u = -K wrap(x - xf) + uf [clipped]   (lqr.py:29-30; cartpole_balancing.ipynb cell 4:24-25;
// quadrotors_model_based_controller.py:36-38, 73-75).
// With aoff = xf's angles the wrapped error IS z on the angle components; on the others the subtraction is
// folded on the host: u = u0 - K z, u0 = uf + sum_{i not an angle} K_i xf_i  (n FMAs per output).
template <bool CLIP>
struct FeedbackCtl {
  static constexpr int KIND = HJB_CTL_FEEDBACK;
  static constexpr bool kClips = CLIP;  // output already inside [umin, umax]: simulate's clip is then a no-op
  template <class S>
  static __device__ __forceinline__ void control(const DevSys& ps, const DevCtl& pc, const float* z,
                                                 const typename S::Trig&, float* u, int = 0) {
#pragma unroll
    for (int k = 0; k < S::M; ++k) {
      float acc = pc.u0[k];
#pragma unroll
      for (int i = 0; i < S::N; ++i) acc = fmaf(-pc.K[k * S::N + i], z[i], acc);
      u[k] = acc;
    }
    if constexpr (CLIP) clip_u<S>(ps, u);
  }
};

// Tracking of a time-varying reference (HJB_CTL_TRACK): u_t = clip(u_ref[t] - K wrap(z - x_ref[t])) with (x_ref, u_ref)(t)
// what Quadrotors2DWaypointsPlanner.update(t dt) returns (controller/quadrotors_model_based_controller.py:77-233) and K
// the hover gain of :36-38.  aoff = 0: z is the raw (wrapped) state.  The reference row is the same for every thread of
// the launch: two uniform 16-byte loads per step, served by L1.
struct TrackCtl {
  static constexpr int KIND = HJB_CTL_TRACK;
  static constexpr bool kClips = true;
  template <class S>
  static __device__ __forceinline__ void control(const DevSys& ps, const DevCtl& pc, const float* z,
                                                 const typename S::Trig&, float* u, int t = 0) {
    constexpr int W = S::N + S::M;
    const int row = min(t + pc.ref_offset, pc.ref_steps - 1);
    float r[W];
    load_row<W>(pc.ref, row, r);
    float d[S::N];
#pragma unroll
    for (int i = 0; i < S::N; ++i) d[i] = z[i] - r[i];
#pragma unroll
    for (int k = 0; k < S::NANG; ++k) d[S::ang(k)] = wrap_pi_<S::kFast>(d[S::ang(k)]);
#pragma unroll
    for (int k = 0; k < S::M; ++k) {
      float acc = r[S::N + k];
#pragma unroll
      for (int i = 0; i < S::N; ++i) acc = fmaf(-pc.K[k * S::N + i], d[i], acc);
      u[k] = acc;
    }
    clip_u<S>(ps, u);
  }
};

// controller/cartpole_energy_shaping.py:65-110 (aoff = 0: z is the raw state).  Both branches are evaluated
// and selected (no divergence).  aux = {Ke0, Ke1, Ke2, eps_energy, eps_state^2, E(xf)}; ps.c[5] = 1/l, ps.c[6] = g/l
struct CartpoleESCtl {
  static constexpr int KIND = HJB_CTL_CARTPOLE_ES;
  static constexpr bool kClips = true;
  template <class S>
  static __device__ __forceinline__ void control(const DevSys& ps, const DevCtl& pc, const float* x,
                                                 const typename S::Trig& t, float* u, int = 0) {
    static_assert(S::KIND == HJB_SYS_CARTPOLE, "cartpole energy shaping needs the cartpole");
    float dx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) dx[i] = x[i] - pc.xf[i];
    dx[1] = wrap_pi_<S::kFast>(dx[1]);                               // :75
    const float de = fmaf(0.5f * x[3], x[3], -t.c) - pc.aux[5];      // :78, :90-95
    const bool near = (fabsf(de) < pc.aux[3]) && (fmaf(dx[1], dx[1], dx[3] * dx[3]) < pc.aux[4]);  // :79
    float ulqr = 0.f;                                                // :80
#pragma unroll
    for (int i = 0; i < 4; ++i) ulqr = fmaf(-pc.K[i], dx[i], ulqr);
    const float ubar = de * x[3] * t.c;                              // :99
    const float a1 = fmaf(pc.aux[2], ubar, -fmaf(pc.aux[0], x[0], pc.aux[1] * x[2]));          // :100
    const float a2 = -fmaf(t.c * ps.c[5], a1, ps.c[6] * t.s);                                    // :101
    float ues = fmaf(ps.c[0], a1, ps.c[1] * t.c * a2);                                           // :102
    ues = fmaf(-ps.c[1] * t.s, x[3] * x[3], ues);                                                // :103
    u[0] = clampf(near ? ulqr : ues, ps.umin[0], ps.umax[0]);        // :86
  }
};

// controller/acrobot_energy_shaping.py:74-121 (Spong collocated swing-up + LQR catch; aoff = 0).
// aux = {Ks0, Ks1, Ks2, eps, E(xf)}
struct AcrobotESCtl {
  static constexpr int KIND = HJB_CTL_ACROBOT_ES;
  static constexpr bool kClips = true;
  template <class S>
  static __device__ __forceinline__ void control(const DevSys& ps, const DevCtl& pc, const float* x,
                                                 const typename S::Trig& t, float* u, int = 0) {
    static_assert(S::KIND == HJB_SYS_ACROBOT, "acrobot energy shaping needs the acrobot");
    float dx[4];
    // (the internal state is wrapped: z in [-pi, pi); a goal angle of 0 — q2 in the reference, :131 — needs no second wrap)
    dx[0] = wrap_pi_<S::kFast>(x[0] - pc.xf[0]);                     // :109
    dx[1] = pc.xf[1] != 0.f ? wrap_pi_<S::kFast>(x[1] - pc.xf[1]) : x[1];
    dx[2] = x[2] - pc.xf[2];
    dx[3] = x[3] - pc.xf[3];
    float quad = 0.f;                                                // :114  dx^T P dx (upper triangle, folded on the host)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float row = pc.P[i * 4 + i] * dx[i];
#pragma unroll
      for (int j = i + 1; j < 4; ++j) row = fmaf(pc.P[i * 4 + j], dx[j], row);
      quad = fmaf(dx[i], row, quad);
    }
    float ulqr = 0.f;                                                // :115
#pragma unroll
    for (int i = 0; i < 4; ++i) ulqr = fmaf(-pc.K[i], dx[i], ulqr);
    typename S::Terms r;
    S::terms(ps, x, t, r);                                           // :83-86
    const float ubar = (S::energy(ps, x, t, r) - pc.aux[4]) * x[2];  // :88
    const float a2 = fmaf(pc.aux[2], ubar, -fmaf(pc.aux[0], x[1], pc.aux[1] * x[3]));  // :90 (wrap(q2) = q2: already wrapped)
    const float i11 = rcp_<S::kFast>(r.m11);
    const float usw = fmaf(fmaf(-r.m12 * r.m12, i11, r.m22), a2, fmaf(-r.m12 * i11, r.h1, r.h2));  // :92
    u[0] = clampf(quad < pc.aux[3] ? ulqr : usw, ps.umin[0], ps.umax[0]);  // :119
  }
};

}  // namespace hjb

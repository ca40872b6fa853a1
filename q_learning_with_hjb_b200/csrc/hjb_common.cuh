// Shared device-side definitions for libhjb_b200 (sm_100a).
//
// Everything a kernel needs about a (system, controller, cost) triple travels in the kernel's parameter
// space (constant bank): the structs below are built on the host by api.cu from the public C structs of
// include/hjb_b200.h, with derived constants folded in double precision before rounding to fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hjb_b200.h"

namespace hjb {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvTwoPi = 0.15915494309189533577f;
// 2*pi split so that k * kTwoPiHi is exact for small |k| (Cody-Waite)
constexpr float kTwoPiHi = 6.28318548202514648438f;   // fp32(2*pi)
constexpr float kTwoPiLo = -1.74845553146951715e-07f; // 2*pi - fp32(2*pi)

// ---------------------------------------------------------------------------------------------
// device parameter blocks (POD, passed by value as __grid_constant__)
// ---------------------------------------------------------------------------------------------
struct DevSys {
  int n, m;
  float dt;
  float umin[HJB_MAX_M], umax[HJB_MAX_M];
  // derived constants, per kind (filled by api.cu::make_dev_sys):
  //  CARTPOLE c = {M11=mc+mp, kappa=mp*l, M22=mp*l^2, gamma=mp*g*l, M11*M22, 1/l, g/l}
  //  ACROBOT  c = {M11_0=I1+I2+m2*l1^2, a=m2*l1*l2/2, I2, G1=(m1*l1/2+m2*l1)*g, G12=m2*g*l2/2}
  //  QUAD2D   c = {g, 1/m, r/I,   dt/m, g dt, r dt/I}      (c[3..5]: the explicit Euler step with dt folded in)
  //  QUAD10D  c = {g, kT/m, n0,   g dt, kT dt/m, n0 dt}
  float c[8];
  float A[16], B[8];  // LINEAR
  // internal-coordinate angle offsets (see systems.cuh): the goal angles of a FEEDBACK controller, else 0
  float aoff[2];
};

struct DevCtl {
  int clip;
  float u0[HJB_MAX_M];  // FEEDBACK: uf + sum_{i not an angle} K_i xf_i (host-folded in double)
  float K[HJB_MAX_M * HJB_MAX_N];
  float P[16];
  float xf[HJB_MAX_N], uf[HJB_MAX_M];
  // CARTPOLE_ES aux = {Ke0, Ke1, Ke2, eps_energy, eps_state^2, E(xf)}
  // ACROBOT_ES  aux = {Ks0, Ks1, Ks2, eps, E(xf)}
  float aux[8];
  // TRACK: the time-varying reference, ref[t] = {x_ref (n), u_ref (m)} for step t (rows past the end repeat the last)
  const float* ref;
  int ref_steps, ref_offset;
};

struct DevCost {
  float Q[HJB_MAX_N * HJB_MAX_N];
  float R[HJB_MAX_M * HJB_MAX_M];
  float xf[HJB_MAX_N], uf[HJB_MAX_M];
  // diagonal fast path: l = sum_i (sq_i z_i + c0_i)^2 + sum_k (sr_k u_k + r0_k)^2 with sq = sqrt(Q_ii),
  // c0 = -sq xf (0 on angle components), sr = sqrt(R_kk), r0 = -sr uf
  float sq[HJB_MAX_N], c0[HJB_MAX_N], sr[HJB_MAX_M], r0[HJB_MAX_M];
  float dang[2];  // aoff - xf on the angle components (0 when the cost and the controller share the goal)
};

// public struct -> device block (derived constants folded in double); defined in api.cu
void make_dev_sys(const hjb_system* s, DevSys& d);

struct DevBox {
  float xf[HJB_MAX_N], lo[HJB_MAX_N], hi[HJB_MAX_N];
  float dang[2];  // aoff - xf on the angle components
};

// ---------------------------------------------------------------------------------------------
// scalar math
// ---------------------------------------------------------------------------------------------
// states_wrap: np.remainder(a + pi, 2 pi) - pi  (floor-mod into [-pi, pi))
__device__ __forceinline__ float wrap_pi(float a) {
  float k = floorf(fmaf(a, kInvTwoPi, 0.5f));
  float r = fmaf(-k, kTwoPiHi, a);
  return fmaf(-k, kTwoPiLo, r);
}
// The fast-math kernels round a / 2 pi to the nearest integer with the 1.5 * 2^23 trick (two FMA-pipe instructions)
// instead of floor (FRND runs on the quarter-rate XU pipe next to sin / cos / rcp: with 6 wraps per acrobot step the XU
// pipe was ~80 % busy).  Same value as wrap_pi except at an exact tie, a = pi (mod 2 pi), which maps to +pi instead of
// -pi: the same angle.
template <bool FAST>
__device__ __forceinline__ float wrap_pi_(float a) {
  if constexpr (FAST) {
    const float k = __fadd_rn(__fmaf_rn(a, kInvTwoPi, 12582912.f), -12582912.f);
    return fmaf(-k, kTwoPiLo, fmaf(-k, kTwoPiHi, a));
  } else {
    return wrap_pi(a);
  }
}

// --- trigonometry of the fast path ---------------------------------------------------------------------------------
// Round 1 used MUFU.SIN / MUFU.COS here.  Measured on B200 over 2^28 points of [-pi, pi) (tests/cuda/trig_probe.cu):
// max |err| 3.5e-7 / 4.0e-7 — and that is the interpolator's own error, not the rounding of x / 2 pi (a first-order
// correction of the argument's rounding error leaves 3.6e-7).  Behind gains of 40-80 (thrust sum, gravity terms) that is
// 2e-5 on a state derivative: outside the 1e-5 per-step bound.  The fast path therefore evaluates sin and cos with a
// quadrant reduction and the classic degree-7 / degree-8 minimax polynomials on [-pi/4, pi/4] (Cephes sinf / cosf
// coefficients): max |err| 6.3e-8 / 7.1e-8 (libdevice: 8.0e-8 / 6.9e-8), 21 instructions for the pair (13 on the FMA
// pipe, 8 on the ALU pipe, none on the quarter-rate XU pipe), no slow-path branch, no local memory.
//   k = rint(2 x / pi) by the 1.5 * 2^23 trick (the quadrant sits in the low mantissa bits of t);
//   r = x - k pi/2 in two FMAs (Cody-Waite; k * fp32(pi/2) is exact inside the FMA for |x| up to thousands).
__device__ __forceinline__ void sincos_poly(float x, float& s, float& c) {
  const float t = __fmaf_rn(x, 0.636619772367581343f, 12582912.f);
  const int q = __float_as_int(t);
  const float k = t - 12582912.f;
  float r = __fmaf_rn(k, -1.5707963705062866f, x);
  r = __fmaf_rn(k, 4.371139006309477e-08f, r);
  const float r2 = r * r;
  float ps = __fmaf_rn(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = __fmaf_rn(ps, r2, -1.6666654611e-1f);
  const float sr = __fmaf_rn(ps, r2 * r, r);
  float pc = __fmaf_rn(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = __fmaf_rn(pc, r2, 4.166664568298827e-2f);
  pc = __fmaf_rn(pc, r2, -0.5f);
  const float cr = __fmaf_rn(pc, r2, 1.0f);
  const float sv = (q & 1) ? cr : sr, cv = (q & 1) ? sr : cr;
  s = __int_as_float(__float_as_int(sv) ^ ((q & 2) << 30));
  c = __int_as_float(__float_as_int(cv) ^ (((q + 1) & 2) << 30));
}

// Rollout kernels (whole horizon in one launch): sin / cos by table + second-order Taylor step.  Each CTA builds, once, a
// shared-memory table of (sin, cos)(k 2^-7 + aoff) for |k 2^-7| <= 3.5 rad (wrapped angles live in [-pi, pi]; 897 entries,
// 7 KB), evaluated in DOUBLE and rounded to fp32 (the internal-coordinate offset aoff is baked in, so the per-step
// "z + aoff" add disappears).  Per evaluation:
//   k = rint(128 x) (1.5 * 2^23 trick), r = x - k / 128 (exact, one FMA, |r| <= 2^-8), one LDS.64, and
//   s = S + r (C - S r / 2),  c = C - r (S + C r / 2)        (next term r^3 / 6 <= 1e-8)
// — 8 FMA-pipe instructions + 3 for index and load (the 2^-5 table of the first version needed the third-order term:
// 11 + 3; an in-line quadrant reduction + minimax polynomials 21).  The FMA pipe is what bounds the step loop
// (tests/cuda/rollout_x2_probe.cu), so these three instructions are 5 % of a quad-2D Euler step.
// The index is clamped (unsigned min): a NaN state reads a valid entry and still yields NaN through r.
constexpr int kTrigLog2 = 7;
constexpr float kTrigRange = 3.5f;
constexpr int kTrigHalf = 7 << (kTrigLog2 - 1);      // 3.5 * 2^7
constexpr int kTrigSize = 2 * kTrigHalf + 1;
__device__ __forceinline__ void sincos_tab(const float2* __restrict__ tab, float x, float& s, float& c) {
  const float t = __fmaf_rn(x, (float)(1 << kTrigLog2), 12582912.f);
  const float k = t - 12582912.f;
  const float r = __fmaf_rn(k, -1.0f / (float)(1 << kTrigLog2), x);
  const unsigned idx = min((unsigned)(__float_as_int(t) - (0x4B400000 - kTrigHalf)), (unsigned)(kTrigSize - 1));
  const float2 e = tab[idx];
  const float hr = 0.5f * r;
  s = __fmaf_rn(r, __fmaf_rn(-e.x, hr, e.y), e.x);
  c = __fmaf_rn(-r, __fmaf_rn(e.y, hr, e.x), e.y);
}

// The same evaluation from a WIDE table entry (S, C, -S/2, -C/2): the halved values come with the load (one LDS.128), so the
// multiply by r / 2 disappears — four FFMAs, bit-identical to sincos_tab (scaling by 1/2 is exact: fma(r, -S/2, C) and
// fma(-S, r/2, C) round the same number).  Used where the tables of all resident CTAs fit in shared memory.
__device__ __forceinline__ void sincos_tab4(const float4* __restrict__ tab, float x, float& s, float& c) {
  const float t = __fmaf_rn(x, (float)(1 << kTrigLog2), 12582912.f);
  const float k = t - 12582912.f;
  const float r = __fmaf_rn(k, -1.0f / (float)(1 << kTrigLog2), x);
  const unsigned idx = min((unsigned)(__float_as_int(t) - (0x4B400000 - kTrigHalf)), (unsigned)(kTrigSize - 1));
  const float4 e = tab[idx];
  s = __fmaf_rn(r, __fmaf_rn(r, e.z, e.y), e.x);
  c = __fmaf_rn(r, __fmaf_rn(r, e.w, -e.x), e.y);
}

template <bool FAST>
__device__ __forceinline__ void sincos_(float a, float& s, float& c) {
  if constexpr (FAST) sincos_poly(a, s, c);
  else sincosf(a, &s, &c);
}
template <bool FAST>
__device__ __forceinline__ float sin_(float a) {
  if constexpr (FAST) { float s, c; sincos_poly(a, s, c); return s; }
  else return sinf(a);
}
template <bool FAST>
__device__ __forceinline__ float cos_(float a) {
  if constexpr (FAST) { float s, c; sincos_poly(a, s, c); return c; }
  else return cosf(a);
}
// one MUFU.RCP (1 ulp): the operands of the fast paths (the determinant of a 2x2 mass matrix, M11, cos of a tilt angle
// inside (-pi/2, pi/2)) are normal numbers far from the ends of the exponent range, so the denormal / overflow fix-up
// sequence that __fdividef adds (compare, select, two multiplies) has nothing to do
__device__ __forceinline__ float rcp_approx(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
template <bool FAST>
__device__ __forceinline__ float tan_(float a) {
  if constexpr (FAST) { float s, c; sincos_poly(a, s, c); return s * rcp_approx(c); }
  else return tanf(a);
}
template <bool FAST>
__device__ __forceinline__ float rcp_(float a) {
  if constexpr (FAST) return rcp_approx(a);
  else return 1.0f / a;
}
// np.clip / jnp.clip propagate NaN (dynamics_basic.py:118, vhjb.py:220); fminf(fmaxf(NaN, lo), hi) would return lo and a
// diverged environment would keep integrating with u = umin and report finite costs.  max.NaN / min.NaN (FMNMX.NAN, the
// same two instructions) keep the NaN visible.
__device__ __forceinline__ float clampf(float v, float lo, float hi) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;\n\tmin.NaN.f32 %0, %0, %3;" : "=&f"(r) : "f"(v), "f"(lo), "f"(hi));
  return r;
}

// ---------------------------------------------------------------------------------------------
// row load / store: W contiguous floats per row, vectorised to the widest aligned access.
// Rows start at multiples of 4*W bytes from a >=16 B aligned base.
// ---------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void load_row(const float* __restrict__ base, int64_t row, float* v) {
  const float* p = base + row * W;
  if constexpr (W % 4 == 0) {
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else if constexpr (W % 2 == 0) {
#pragma unroll
    for (int i = 0; i < W / 2; ++i) {
      float2 t = __ldg(reinterpret_cast<const float2*>(p) + i);
      v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = __ldg(p + i);
  }
}

// streaming (evict-first) store: trajectories are written once and never re-read by the kernel
template <int W>
__device__ __forceinline__ void store_row(float* __restrict__ base, int64_t row, const float* v) {
  float* p = base + row * W;
  if constexpr (W % 4 == 0) {
#pragma unroll
    for (int i = 0; i < W / 4; ++i)
      __stcs(reinterpret_cast<float4*>(p) + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
  } else if constexpr (W % 2 == 0) {
#pragma unroll
    for (int i = 0; i < W / 2; ++i) __stcs(reinterpret_cast<float2*>(p) + i, make_float2(v[2 * i], v[2 * i + 1]));
  } else {
#pragma unroll
    for (int i = 0; i < W; ++i) __stcs(p + i, v[i]);
  }
}

}  // namespace hjb

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
  //  QUAD2D   c = {g, 1/m, r/I}
  //  QUAD10D  c = {g, kT/m, n0}
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

template <bool FAST>
__device__ __forceinline__ void sincos_(float a, float& s, float& c) {
  if constexpr (FAST) {
    s = __sinf(a);
    c = __cosf(a);
  } else {
    sincosf(a, &s, &c);
  }
}
template <bool FAST>
__device__ __forceinline__ float sin_(float a) {
  if constexpr (FAST) return __sinf(a);
  else return sinf(a);
}
template <bool FAST>
__device__ __forceinline__ float cos_(float a) {
  if constexpr (FAST) return __cosf(a);
  else return cosf(a);
}
// one MUFU.RCP: the operands of the fast paths (the determinant of a 2x2 mass matrix, M11, cos of a tilt angle inside
// (-pi/2, pi/2)) are normal numbers far from the ends of the exponent range, so the denormal / overflow fix-up sequence
// that __fdividef adds (compare, select, two multiplies) has nothing to do
__device__ __forceinline__ float rcp_approx(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
template <bool FAST>
__device__ __forceinline__ float tan_(float a) {
  if constexpr (FAST) return __sinf(a) * rcp_approx(__cosf(a));
  else return tanf(a);
}
template <bool FAST>
__device__ __forceinline__ float rcp_(float a) {
  if constexpr (FAST) return rcp_approx(a);
  else return 1.0f / a;
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

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

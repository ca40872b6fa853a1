// Per-state epilogue of the vhjb pass, shared by the kernels: from the value-net outputs of ONE state to its control,
// Hamiltonian residual, loss contributions and (GRAD) adjoint seeds.
//   reference: controller/vhjb.py:204-221 (get_control_efforts_with_additional_term), :162-165 (running_cost),
//   :227-253 (hjb_loss, termination_loss); reverse pass hand-derived, verified against autograd by
//   oracle/vhjb_oracle.py::closed_form_grads.
#pragma once
#include "vhjb_simt.cuh"

namespace hjb {

// g0 = dV/dh0 through the net (before the 1/sd and eps_s terms), Vy = |y|^2, z = wrap(x - xf), zz = |z|^2,
// lz = z^T Q z, f/G = dynamics at x.  Adds this state's terms to hjb_sum / term_sum, writes the requested per-state
// outputs, and returns pbar = dL/dp (N values) and Vbar = dL/dV.
template <class S, int UFORM, int RFORM, bool GRAD>
__device__ __forceinline__ void state_epilogue(const VhjbArgs& a, const float* g0, float Vy, const float* z, float zz, float lz,
                                               const float* f, const float* G, float done, float cost, bool valid, int64_t idx,
                                               float inv_norm0, float inv_norm1, float& hjb_sum, float& term_sum, float* pbar,
                                               float& Vbar, float icost_pre = 0.f) {
  constexpr int N = S::N, M = S::M;
  float p[N];
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fmaf(g0[i], a.inv_std[i], 2.f * a.eps_s * z[i]);
  const float V = fmaf(a.eps_s, zz, Vy);
  float c[M], u[M], du[M];
  bool inside[M];
#pragma unroll
  for (int k = 0; k < M; ++k) {
    // (gradient kernel: two partial sums, half the dependent chain — its epilogue is one warp per scheduler on the
    // critical path of the tile; the residual kernels run at their register cap and overlap their epilogues)
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (GRAD && (i & 1)) s1 = fmaf(p[i], G[i * M + k], s1);
      else s0 = fmaf(p[i], G[i * M + k], s0);
    }
    c[k] = GRAD ? s0 + s1 : s0;
  }
#pragma unroll
  for (int k = 0; k < M; ++k) {
    if constexpr (UFORM == HJB_U_CLIPPED) {
      float ur = a.uf[k];
#pragma unroll
      for (int jj = 0; jj < M; ++jj) ur = fmaf(-0.5f * a.Rinv[k * M + jj], c[jj], ur);
      inside[k] = (ur > a.sys.umin[k]) && (ur < a.sys.umax[k]);
      u[k] = clampf(ur, a.sys.umin[k], a.sys.umax[k]);
    } else {
      inside[k] = false;
      u[k] = -sign0(c[k]);
    }
    du[k] = u[k] - a.uf[k];
  }
  float xdot[N], vdot0 = 0.f, vdot1 = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = f[i];
#pragma unroll
    for (int k = 0; k < M; ++k) s = fmaf(G[i * M + k], u[k], s);
    xdot[i] = s;
    if (GRAD && (i & 1)) vdot1 = fmaf(p[i], s, vdot1);
    else vdot0 = fmaf(p[i], s, vdot0);
  }
  const float vdot = GRAD ? vdot0 + vdot1 : vdot0;
  float r;
#pragma unroll
  for (int i = 0; i < N; ++i) pbar[i] = 0.f;
  Vbar = 0.f;
  if constexpr (RFORM == HJB_RES_NORMALIZED) {
    float l = lz;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      float row = 0.f;
#pragma unroll
      for (int jj = 0; jj < M; ++jj) row = fmaf(a.R[k * M + jj], du[jj], row);
      l = fmaf(du[k], row, l);
    }
    const float den = l + a.eps;
    const float iden = 1.0f / den;
    r = fmaf(vdot, iden, 1.f);
    // 1 / (cost + eps) depends on the input alone: the gradient kernel computes it a phase ahead (icost_pre), off the
    // critical path of the epilogue (three IEEE divisions were a third of its samples)
    const float icost = icost_pre > 0.f ? icost_pre : 1.0f / (cost + a.eps);
    const float tq = fmaf(V, icost, -1.f);
    if (valid) {
      hjb_sum += fabsf(r) * (1.f - done);
      term_sum += fabsf(tq) * done;
    }
    if constexpr (GRAD) {
      const float rbar = valid ? (1.f - done) * inv_norm0 * sign0(r) : 0.f;
      const float vbar = rbar * iden;
      const float lbar = -rbar * vdot * iden * iden;
      float t[M];
#pragma unroll
      for (int k = 0; k < M; ++k) {
        float ub = vbar * c[k];
#pragma unroll
        for (int jj = 0; jj < M; ++jj) ub = fmaf(lbar * a.Rsym[k * M + jj], du[jj], ub);
        t[k] = inside[k] ? ub : 0.f;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) pbar[i] = vbar * xdot[i];
#pragma unroll
      for (int jj = 0; jj < M; ++jj) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < M; ++k) s = fmaf(t[k], a.Rinv[k * M + jj], s);
        s *= -0.5f;
#pragma unroll
        for (int i = 0; i < N; ++i) pbar[i] = fmaf(G[i * M + jj], s, pbar[i]);
      }
      Vbar = valid ? a.reg * done * inv_norm1 * sign0(tq) * icost : 0.f;
    }
  } else {
    r = vdot + cost;
    if (valid) hjb_sum += fabsf(r);
    if constexpr (GRAD) {
      const float rbar = valid ? inv_norm0 * sign0(r) : 0.f;
#pragma unroll
      for (int i = 0; i < N; ++i) pbar[i] = rbar * xdot[i];
    }
  }
  if (valid) {
    if (a.V) a.V[idx] = V;
    if (a.r) a.r[idx] = r;
    if (a.p) store_row<N>(a.p, idx, p);
    if (a.u) store_row<M>(a.u, idx, u);
  }
}

}  // namespace hjb

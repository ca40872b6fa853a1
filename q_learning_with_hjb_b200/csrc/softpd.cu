// The notebooks' soft-PD baseline — an unconstrained value net with biases and a Dense(1) head,
//     z = wrap(x - xf);  V = w4 . s(a3) + b4,  a3 = s(a2) W3 + b3,  a2 = s(a1) W2 + b2,  a1 = z W1 + b1,   s = tanh | relu
// (SoftPDValueApproximator: examples/cartpole_balancing.ipynb cell 6, examples/drone_hovering.ipynb cell 6), trained on
//     mean_i [ res_i + reg max(0, V(xf) - V(x_i)) ],   res = |vdot + l|  (cart-pole nb cell 11)  |  |vdot / (l + eps) + 1|  (drone nb)
// with u = clip(-R^-1 g^T dV/dx / 2 + uf), and warmed up on  |V - z^T P z|  (cart-pole nb) or on the same residual under the
// LQR's control clip(-K z + uf) (drone nb).  One fused fp32 CUDA-core kernel per batch: forward, input gradient, control,
// residual, and the full parameter gradient (weights AND biases; sigma'' terms for tanh) — the structure of vhjb_simt.cuh
// (feature-major activations of a 32-state tile in shared memory, weights resident, weight-gradient accumulators in
// registers, per-CTA partials reduced in a fixed order), plus
//   * biases in the three hidden layers and the head, bias gradients as warp-shuffle column sums of the adjoints;
//   * the head  V = w4 . s(a3) + b4  instead of |y|^2: gy = w4 s'(a3), a3-bar = gy-bar w4 s''(a3) + V-bar w4 s'(a3);
//   * the hinge on V(xf): V(xf) is evaluated first (a one-state launch), the main launch counts the states with V < V(xf),
//     and a one-state backward launch at xf with the seed reg count / B adds d V(xf) / d theta.
// This is a baseline of the notebooks, not the headline path: CUDA cores only.
#include <cmath>
#include <cstring>

#include "vhjb_simt.cuh"

namespace hjb {

constexpr int kSoftSlots = 160;

__host__ __device__ constexpr int soft_param_count(int n) { return n * VH1 + VH1 + VH1 * VH2 + VH2 + VH2 * VH3 + VH3 + VH3 + 1; }
static int64_t soft_pstride(int n) { return ((int64_t)soft_param_count(n) + 4 + 3) / 4 * 4; }

enum SoftLoss { SOFT_HJB = 0, SOFT_VALUE_MATCH = 1, SOFT_HJB_LQR = 2, SOFT_XF_BACKWARD = 3 };

struct SoftArgs {
  DevSys sys;
  const float* params;   // [W1 | b1 | W2 | b2 | W3 | b3 | w4 | b4]
  float xf[HJB_MAX_N];
  float Q[HJB_MAX_N * HJB_MAX_N], R[HJB_MAX_M * HJB_MAX_M], Rsym[HJB_MAX_M * HJB_MAX_M], Rinv[HJB_MAX_M * HJB_MAX_M];
  float uf[HJB_MAX_M];
  float K[HJB_MAX_M * HJB_MAX_N];   // SOFT_HJB_LQR
  float Pm[HJB_MAX_N * HJB_MAX_N];  // SOFT_VALUE_MATCH
  float eps, reg, inv_B;
  int loss, normalized;
  const float* xs;
  int64_t B, n_tiles;
  const float* v0;      // device scalar V(xf) (hinge)
  const float* sums;    // SOFT_XF_BACKWARD: reduced sums of the main launch: [res, hinge, count]
  float* V;
  float* p;
  float* u;
  float* partial;
  int64_t pstride;
  int slot0;
  int at_xf;            // the batch is the single state x = xf (the two one-state launches of the hinge term)
};

__host__ __device__ constexpr int soft_smem_floats(int n) {
  return soft_param_count(n) + 6 * VH1 * VLD + 2 * VH3 * VLD + 3 * n * VLD + VBM;
}

template <class S, int ACT, bool GRAD>
__global__ void __launch_bounds__(VTHREADS, 1) softpd_kernel(const __grid_constant__ SoftArgs a) {
  constexpr int N = S::N, M = S::M;
  constexpr int NA1 = (N + 1) / 2;
  extern __shared__ __align__(16) float smem[];
  float* sW1 = smem;
  float* sb1 = sW1 + N * VH1;
  float* sW2 = sb1 + VH1;
  float* sb2 = sW2 + VH1 * VH2;
  float* sW3 = sb2 + VH2;
  float* sb3 = sW3 + VH2 * VH3;
  float* sw4 = sb3 + VH3;
  float* sb4 = sw4 + VH3;
  float* sA1 = smem + ((soft_param_count(N) + 3) / 4) * 4;
  float* sA2 = sA1 + VH1 * VLD;
  float* sB1 = sA2 + VH2 * VLD;
  float* sB2 = sB1 + VH1 * VLD;
  float* sT1 = sB2 + VH2 * VLD;
  float* sT2 = sT1 + VH1 * VLD;
  float* sY = sT2 + VH2 * VLD;     // a3
  float* sYb = sY + VH3 * VLD;     // a3-bar
  float* sH0 = sYb + VH3 * VLD;    // z (the net's input)
  float* sP = sH0 + N * VLD;       // dV/dz, then p-bar
  float* sVb = sP + N * VLD;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ti = tid >> 4, to = tid & 15;
  const int o1 = tid & 127, ih1 = tid >> 7;
  {
    const int P = soft_param_count(N);
    for (int i = tid; i < P; i += VTHREADS) smem[i] = __ldg(a.params + i);
  }
  float acc1[NA1];
  float acc2[8][8];
  float acc3[8][4];
  float ab1[16], ab2[16], ab3[8], aw4[8];   // bias / head-weight gradients of this warp's features (lane 0 holds the sums)
  float ab4 = 0.f;
  if constexpr (GRAD) {
#pragma unroll
    for (int i = 0; i < NA1; ++i) acc1[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc2[i][j] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc3[i][j] = 0.f;
      ab3[i] = 0.f;
      aw4[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { ab1[i] = 0.f; ab2[i] = 0.f; }
  }
  float res_sum = 0.f, hinge_sum = 0.f, hinge_cnt = 0.f;   // warp 0
  const float v0 = (a.v0 != nullptr) ? __ldg(a.v0) : 0.f;
  __syncthreads();
  auto wsum = [](float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
  };

  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t idx = tile * VBM + lane;
    const bool valid = a.at_xf ? lane == 0 : idx < a.B;
    float xraw[N];
    if (warp == 0) {
      if (valid && !a.at_xf) load_row<N>(a.xs, idx, xraw);
      else {
#pragma unroll
        for (int i = 0; i < N; ++i) xraw[i] = a.xf[i];
      }
      float z[N];
#pragma unroll
      for (int i = 0; i < N; ++i) z[i] = xraw[i] - a.xf[i];
      wrap_state<S>(z);
#pragma unroll
      for (int i = 0; i < N; ++i) sH0[i * VLD + lane] = z[i];
    }
    __syncthreads();
    // ---- forward ----
    gemm_fwd<N, VH1>(sW1, [&](int k, int r) { return sH0[k * VLD + r]; },
                     [&](int j, int r, float v) { sA1[j * VLD + r] = v + sb1[j]; }, warp, lane);
    __syncthreads();
    gemm_fwd<VH1, VH2>(sW2, [&](int k, int r) { return act_f<ACT>(sA1[k * VLD + r]); },
                       [&](int j, int r, float v) { sA2[j * VLD + r] = v + sb2[j]; }, warp, lane);
    __syncthreads();
    gemm_fwd<VH2, VH3>(sW3, [&](int k, int r) { return act_f<ACT>(sA2[k * VLD + r]); },
                       [&](int j, int r, float v) { sY[j * VLD + r] = v + sb3[j]; }, warp, lane);
    __syncthreads();
    // ---- input gradient: gy = w4 s'(a3), then back through W3, W2, W1 ----
    gemm_bwd<VH3, VH2>(sW3, [&](int o, int r) { return sw4[o] * act_d1<ACT>(sY[o * VLD + r]); },
                       [&](int i, int r, float v) { sB2[i * VLD + r] = v; }, warp, lane);
    __syncthreads();
    gemm_bwd<VH2, VH1>(sW2, [&](int o, int r) { return sB2[o * VLD + r] * act_d1<ACT>(sA2[o * VLD + r]); },
                       [&](int i, int r, float v) { sB1[i * VLD + r] = v; }, warp, lane);
    __syncthreads();
    gemm_bwd<VH1, N>(sW1, [&](int o, int r) { return sB1[o * VLD + r] * act_d1<ACT>(sA1[o * VLD + r]); },
                     [&](int i, int r, float v) { sP[i * VLD + r] = v; }, warp, lane);
    __syncthreads();
    // ---- per-state epilogue ----
    if (warp == 0) {
      float z[N], p[N];
      float V = sb4[0];
#pragma unroll 8
      for (int j = 0; j < VH3; ++j) V = fmaf(sw4[j], act_f<ACT>(sY[j * VLD + lane]), V);
#pragma unroll
      for (int i = 0; i < N; ++i) { z[i] = sH0[i * VLD + lane]; p[i] = sP[i * VLD + lane]; }
      float zi[N], f[N], G[N * M];
      to_internal<S>(a.sys, xraw, zi);
      typename S::Trig tr;
      S::trig(a.sys, zi, tr);
      S::fg(a.sys, zi, tr, f, G);
      float c[M], u[M], du[M];
      bool inside[M];
#pragma unroll
      for (int k = 0; k < M; ++k) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < N; ++i) s = fmaf(p[i], G[i * M + k], s);
        c[k] = s;
      }
#pragma unroll
      for (int k = 0; k < M; ++k) {
        float ur = a.uf[k];
        if (a.loss == SOFT_HJB_LQR) {
#pragma unroll
          for (int i = 0; i < N; ++i) ur = fmaf(-a.K[k * N + i], z[i], ur);
          inside[k] = false;                       // u does not depend on the net: no gradient through it
        } else {
#pragma unroll
          for (int j = 0; j < M; ++j) ur = fmaf(-0.5f * a.Rinv[k * M + j], c[j], ur);
          inside[k] = (ur > a.sys.umin[k]) && (ur < a.sys.umax[k]);
        }
        u[k] = clampf(ur, a.sys.umin[k], a.sys.umax[k]);
        du[k] = u[k] - a.uf[k];
      }
      if (valid) {
        if (a.V) a.V[idx] = V;
        if (a.p) store_row<N>(a.p, idx, p);
        if (a.u) store_row<M>(a.u, idx, u);
      }
      if constexpr (GRAD) {
        float pbar[N], Vbar = 0.f;
#pragma unroll
        for (int i = 0; i < N; ++i) pbar[i] = 0.f;
        if (a.loss == SOFT_XF_BACKWARD) {
          Vbar = valid ? a.reg * __ldg(a.sums + 2) * a.inv_B : 0.f;     // d/dV(xf) of reg mean_i max(0, V(xf) - V_i)
        } else if (a.loss == SOFT_VALUE_MATCH) {
          float tgt = 0.f;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            float row = 0.f;
#pragma unroll
            for (int j = 0; j < N; ++j) row = fmaf(a.Pm[i * N + j], z[j], row);
            tgt = fmaf(z[i], row, tgt);
          }
          const float d = V - tgt;
          if (valid) { res_sum += fabsf(d); Vbar = sign0(d) * a.inv_B; }
        } else {
          float xdot[N], vdot = 0.f;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            float s = f[i];
#pragma unroll
            for (int k = 0; k < M; ++k) s = fmaf(G[i * M + k], u[k], s);
            xdot[i] = s;
            vdot = fmaf(p[i], s, vdot);
          }
          float l = 0.f;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            float row = 0.f;
#pragma unroll
            for (int j = 0; j < N; ++j) row = fmaf(a.Q[i * N + j], z[j], row);
            l = fmaf(z[i], row, l);
          }
#pragma unroll
          for (int k = 0; k < M; ++k) {
            float row = 0.f;
#pragma unroll
            for (int j = 0; j < M; ++j) row = fmaf(a.R[k * M + j], du[j], row);
            l = fmaf(du[k], row, l);
          }
          float r, vbar, lbar;
          if (a.normalized) {
            const float iden = 1.0f / (l + a.eps);
            r = fmaf(vdot, iden, 1.f);
            const float rbar = valid ? sign0(r) * a.inv_B : 0.f;
            vbar = rbar * iden;
            lbar = -rbar * vdot * iden * iden;
          } else {
            r = vdot + l;
            const float rbar = valid ? sign0(r) * a.inv_B : 0.f;
            vbar = rbar;
            lbar = rbar;
          }
          const bool below = V < v0;                              // hinge max(0, V(xf) - V) active
          if (valid) {
            res_sum += fabsf(r);
            hinge_sum += below ? v0 - V : 0.f;
            hinge_cnt += below ? 1.f : 0.f;
            Vbar = below ? -a.reg * a.inv_B : 0.f;
          }
#pragma unroll
          for (int i = 0; i < N; ++i) pbar[i] = vbar * xdot[i];
          float t[M];
#pragma unroll
          for (int k = 0; k < M; ++k) {
            float ub = vbar * c[k];
#pragma unroll
            for (int j = 0; j < M; ++j) ub = fmaf(lbar * a.Rsym[k * M + j], du[j], ub);
            t[k] = inside[k] ? ub : 0.f;
          }
#pragma unroll
          for (int j = 0; j < M; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < M; ++k) s = fmaf(t[k], a.Rinv[k * M + j], s);
            s *= -0.5f;
#pragma unroll
            for (int i = 0; i < N; ++i) pbar[i] = fmaf(G[i * M + j], s, pbar[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) sP[i * VLD + lane] = pbar[i];   // g0-bar (no input normalisation in the notebooks' nets)
        sVb[lane] = Vbar;
        ab4 += wsum(Vbar);
      }
    }
    __syncthreads();
    if constexpr (GRAD) {
      // ---- reverse pass ----
      gemm_fwd<N, VH1>(sW1, [&](int k, int r) { return sP[k * VLD + r]; },
                       [&](int j, int r, float v) { sT1[j * VLD + r] = v; }, warp, lane);          // g1-bar
      __syncthreads();
      gemm_fwd<VH1, VH2>(sW2, [&](int k, int r) { return sT1[k * VLD + r] * act_d1<ACT>(sA1[k * VLD + r]); },
                         [&](int j, int r, float v) { sT2[j * VLD + r] = v; }, warp, lane);        // g2-bar
      __syncthreads();
      gemm_fwd<VH2, VH3>(sW3, [&](int k, int r) { return sT2[k * VLD + r] * act_d1<ACT>(sA2[k * VLD + r]); },
                         [&](int j, int r, float v) {                                               // v = gy-bar
                           const float a3 = sY[j * VLD + r], d1 = act_d1<ACT>(a3), vb = sVb[r];
                           const float a3b = sw4[j] * fmaf(v, act_d2<ACT>(a3), vb * d1);            // a3-bar
                           sYb[j * VLD + r] = a3b;
                           const float cw = wsum(fmaf(v, d1, vb * act_f<ACT>(a3)));                 // w4-bar
                           const float cb = wsum(a3b);                                              // b3-bar
                           aw4[j & 7] += cw;
                           ab3[j & 7] += cb;
                         }, warp, lane);
      __syncthreads();
      wgrad<8, 8>(acc2, [&](int i, int r) { return sT1[i * VLD + r] * act_d1<ACT>(sA1[i * VLD + r]); },
                  [&](int o, int r) { return sB2[o * VLD + r] * act_d1<ACT>(sA2[o * VLD + r]); }, ti, to);
      wgrad<8, 4>(acc3, [&](int i, int r) { return sT2[i * VLD + r] * act_d1<ACT>(sA2[i * VLD + r]); },
                  [&](int o, int r) { return sw4[o] * act_d1<ACT>(sY[o * VLD + r]); }, ti, to);
      wgrad<8, 4>(acc3, [&](int i, int r) { return act_f<ACT>(sA2[i * VLD + r]); },
                  [&](int o, int r) { return sYb[o * VLD + r]; }, ti, to);
      {
        const float* g1b = sB1 + o1 * VLD;
        const float* g1a = sA1 + o1 * VLD;
#pragma unroll 4
        for (int r = 0; r < VBM; ++r) {
          const float g1 = g1b[r] * act_d1<ACT>(g1a[r]);
#pragma unroll
          for (int q = 0; q < NA1; ++q) {
            const int i = ih1 + 2 * q;
            if (i < N) acc1[q] = fmaf(sP[i * VLD + r], g1, acc1[q]);
          }
        }
      }
      __syncthreads();
      gemm_bwd<VH3, VH2>(sW3, [&](int o, int r) { return sYb[o * VLD + r]; },
                         [&](int i, int r, float v) {
                           const float a2 = sA2[i * VLD + r];
                           float out = v * act_d1<ACT>(a2);
                           if constexpr (ACT != HJB_ACT_RELU) out = fmaf(sT2[i * VLD + r] * sB2[i * VLD + r], act_d2<ACT>(a2), out);
                           sT2[i * VLD + r] = out;                                                  // a2-bar
                           ab2[i & 15] += wsum(out);
                         }, warp, lane);
      __syncthreads();
      wgrad<8, 8>(acc2, [&](int i, int r) { return act_f<ACT>(sA1[i * VLD + r]); },
                  [&](int o, int r) { return sT2[o * VLD + r]; }, ti, to);
      gemm_bwd<VH2, VH1>(sW2, [&](int o, int r) { return sT2[o * VLD + r]; },
                         [&](int i, int r, float v) {
                           const float a1 = sA1[i * VLD + r];
                           float out = v * act_d1<ACT>(a1);
                           if constexpr (ACT != HJB_ACT_RELU) out = fmaf(sT1[i * VLD + r] * sB1[i * VLD + r], act_d2<ACT>(a1), out);
                           sT1[i * VLD + r] = out;                                                  // a1-bar
                           ab1[i & 15] += wsum(out);
                         }, warp, lane);
      __syncthreads();
      {
        const float* a1b = sT1 + o1 * VLD;
#pragma unroll 4
        for (int r = 0; r < VBM; ++r) {
          const float v = a1b[r];
#pragma unroll
          for (int q = 0; q < NA1; ++q) {
            const int i = ih1 + 2 * q;
            if (i < N) acc1[q] = fmaf(sH0[i * VLD + r], v, acc1[q]);
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- per-CTA partials ----
  float* part = a.partial + (int64_t)(blockIdx.x + a.slot0) * a.pstride;
  if constexpr (GRAD) {
    float* pW1 = part;
    float* pb1 = pW1 + N * VH1;
    float* pW2 = pb1 + VH1;
    float* pb2 = pW2 + VH1 * VH2;
    float* pW3 = pb2 + VH2;
    float* pb3 = pW3 + VH2 * VH3;
    float* pw4 = pb3 + VH3;
    float* pb4 = pw4 + VH3;
#pragma unroll
    for (int q = 0; q < NA1; ++q) {
      const int i = ih1 + 2 * q;
      if (i < N) pW1[i * VH1 + o1] = acc1[q];
    }
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
      for (int y = 0; y < 8; ++y) pW2[(ti + 16 * x) * VH2 + (to + 16 * y)] = acc2[x][y];
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) pW3[(ti + 16 * x) * VH3 + (to + 16 * y)] = acc3[x][y];
    if (lane == 0) {   // this warp's features: 16 of the 128-wide layers (gemm_bwd rows warp * 16 + q), 8 of the head's 64
#pragma unroll
      for (int q = 0; q < 16; ++q) { pb1[warp * 16 + q] = ab1[q]; pb2[warp * 16 + q] = ab2[q]; }
#pragma unroll
      for (int q = 0; q < 8; ++q) { pb3[warp * 8 + q] = ab3[q]; pw4[warp * 8 + q] = aw4[q]; }
      if (warp == 0) pb4[0] = ab4;
    }
  }
  if (warp == 0) {
    res_sum = wsum(res_sum);
    hinge_sum = wsum(hinge_sum);
    hinge_cnt = wsum(hinge_cnt);
    if (lane == 0) {
      const int P = soft_param_count(N);
      part[P] = res_sum;
      part[P + 1] = hinge_sum;
      part[P + 2] = hinge_cnt;
    }
  }
}

// out[j] = sum over the CTAs' partial slots [0, ncta) and, if nx, the xf slot
__global__ void __launch_bounds__(256) soft_reduce_kernel(const float* __restrict__ partial, int64_t pstride, int ncta, int xslot,
                                                          int first, int count, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  float s = 0.f;
  for (int c = 0; c < ncta; ++c) s += partial[(int64_t)c * pstride + first + j];
  if (xslot >= 0) s += partial[(int64_t)xslot * pstride + first + j];
  out[j] = s;
}

static int soft_sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
      cached = sms;
  }
  return cached > 0 ? (cached < kSoftSlots - 1 ? cached : kSoftSlots - 1) : 148;
}

template <class S, int ACT>
static cudaError_t soft_launch(const SoftArgs& a, int grid, bool grad, cudaStream_t st) {
  const size_t smem = sizeof(float) * (soft_smem_floats(S::N) + 4);
  cudaError_t e;
  if (grad) {
    auto k = softpd_kernel<S, ACT, true>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, VTHREADS, smem, st>>>(a);
  } else {
    auto k = softpd_kernel<S, ACT, false>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, VTHREADS, smem, st>>>(a);
  }
  return cudaGetLastError();
}

static cudaError_t soft_dispatch(int kind, int act, const SoftArgs& a, int grid, bool grad, cudaStream_t st) {
  if (kind == HJB_SYS_CARTPOLE) {
    if (act == HJB_ACT_TANH) return soft_launch<CartpoleSys<false>, HJB_ACT_TANH>(a, grid, grad, st);
    if (act == HJB_ACT_RELU) return soft_launch<CartpoleSys<false>, HJB_ACT_RELU>(a, grid, grad, st);
  } else if (kind == HJB_SYS_QUAD2D) {
    if (act == HJB_ACT_TANH) return soft_launch<Quad2DSys<false>, HJB_ACT_TANH>(a, grid, grad, st);
    if (act == HJB_ACT_RELU) return soft_launch<Quad2DSys<false>, HJB_ACT_RELU>(a, grid, grad, st);
  }
  return cudaErrorNotSupported;
}

}  // namespace hjb

using namespace hjb;

extern "C" {

int64_t hjb_softpd_param_count(int32_t n) { return n > 0 && n <= HJB_MAX_N ? soft_param_count(n) : -1; }

int64_t hjb_softpd_workspace_bytes(int32_t n) {
  if (n <= 0 || n > HJB_MAX_N) return -1;
  return ((int64_t)kSoftSlots * soft_pstride(n) + 8) * (int64_t)sizeof(float);
}

static int soft_fill(const hjb_system* sys, const hjb_softpd* net, SoftArgs& a, void* workspace) {
  if (!sys || !net || !net->params || !workspace) return HJB_ERR_BAD_ARG;
  if (net->n != sys->n) return HJB_ERR_UNSUPPORTED;
  std::memset(&a, 0, sizeof(a));
  make_dev_sys(sys, a.sys);
  const int n = sys->n, m = sys->m;
  a.params = net->params;
  for (int i = 0; i < n; ++i) a.xf[i] = net->xf[i];
  for (int i = 0; i < n * n; ++i) { a.Q[i] = net->Q[i]; a.Pm[i] = net->P[i]; }
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      a.R[i * m + j] = net->R[i * m + j];
      a.Rsym[i * m + j] = net->R[i * m + j] + net->R[j * m + i];
      a.Rinv[i * m + j] = net->Rinv[i * m + j];
    }
  for (int i = 0; i < m; ++i) a.uf[i] = net->uf[i];
  for (int i = 0; i < m * n; ++i) a.K[i] = net->K[i];
  a.eps = net->eps;
  a.normalized = net->normalized_residual;
  a.partial = static_cast<float*>(workspace);
  a.pstride = soft_pstride(n);
  return HJB_OK;
}

int hjb_softpd_policy(const hjb_system* sys, const hjb_softpd* net, const float* xs, int64_t B, float* V, float* p, float* u,
                      void* workspace, void* stream) {
  SoftArgs a;
  const int rc = soft_fill(sys, net, a, workspace);
  if (rc != HJB_OK) return rc;
  if (B < 0) return HJB_ERR_BAD_ARG;
  if (B == 0) return HJB_OK;
  if (!xs) return HJB_ERR_BAD_ARG;
  a.xs = xs; a.B = B; a.n_tiles = (B + VBM - 1) / VBM;
  a.V = V; a.p = p; a.u = u;
  a.loss = SOFT_HJB;
  const int grid = (int)(a.n_tiles < soft_sm_count() ? a.n_tiles : soft_sm_count());
  const cudaError_t e = soft_dispatch(sys->kind, net->act, a, grid, false, (cudaStream_t)stream);
  return e == cudaSuccess ? HJB_OK : (e == cudaErrorNotSupported ? HJB_ERR_UNSUPPORTED : (int)e);
}

int hjb_softpd_loss_grad(const hjb_system* sys, const hjb_softpd* net, const float* xs, int64_t B, int32_t loss_form, float reg,
                         float* grad, float* sums, void* workspace, void* stream) {
  SoftArgs a;
  const int rc = soft_fill(sys, net, a, workspace);
  if (rc != HJB_OK) return rc;
  if (B <= 0 || !xs || !grad || !sums || loss_form < SOFT_HJB || loss_form > SOFT_HJB_LQR) return HJB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = sys->n, P = soft_param_count(n);
  float* tail = a.partial + (int64_t)kSoftSlots * a.pstride;   // [0]: V(xf)
  const bool hinge = loss_form != SOFT_VALUE_MATCH;
  cudaError_t e;
  if (hinge) {   // V(xf): one state, value only
    SoftArgs v = a;
    v.at_xf = 1; v.B = 1; v.n_tiles = 1; v.V = tail; v.loss = SOFT_HJB;
    e = soft_dispatch(sys->kind, net->act, v, 1, false, st);
    if (e != cudaSuccess) return e == cudaErrorNotSupported ? HJB_ERR_UNSUPPORTED : (int)e;
  }
  a.xs = xs; a.B = B; a.n_tiles = (B + VBM - 1) / VBM;
  a.loss = loss_form; a.reg = reg; a.inv_B = (float)(1.0 / (double)B);
  a.v0 = hinge ? tail : nullptr;
  const int grid = (int)(a.n_tiles < soft_sm_count() ? a.n_tiles : soft_sm_count());
  e = soft_dispatch(sys->kind, net->act, a, grid, true, st);
  if (e != cudaSuccess) return e == cudaErrorNotSupported ? HJB_ERR_UNSUPPORTED : (int)e;
  soft_reduce_kernel<<<1, 256, 0, st>>>(a.partial, a.pstride, grid, -1, P, 3, sums);     // [res sum, hinge sum, count]
  int xslot = -1;
  if (hinge) {   // d V(xf) / d theta with the seed reg count / B
    SoftArgs x = a;
    x.at_xf = 1; x.B = 1; x.n_tiles = 1; x.loss = SOFT_XF_BACKWARD; x.sums = sums; x.v0 = nullptr;
    x.slot0 = kSoftSlots - 1;
    xslot = kSoftSlots - 1;
    e = soft_dispatch(sys->kind, net->act, x, 1, true, st);
    if (e != cudaSuccess) return (int)e;
  }
  soft_reduce_kernel<<<(P + 255) / 256, 256, 0, st>>>(a.partial, a.pstride, grid, xslot, 0, P, grad);
  e = cudaGetLastError();
  return e == cudaSuccess ? HJB_OK : (int)e;
}

}  // extern "C"

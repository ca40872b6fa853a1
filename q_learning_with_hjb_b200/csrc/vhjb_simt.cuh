// K2/K3 (CUDA-core version) — fused vhjb pass: value-MLP forward, input gradient dV/dx, optimal control,
// Hamiltonian residual and (GRAD) the full parameter gradient, for one tile of 32 sampled states at a time.
//
// One persistent CTA per SM (256 threads).  The three weight matrices live in shared memory (fp32, ~101 KB) for
// the whole launch; activations of the current tile live in shared memory FEATURE-MAJOR ([feature][state], leading
// dimension 33) so that lane <-> state makes every activation access conflict-free and every weight access a
// broadcast; the 17 GEMMs of SURVEY.md 8a-V6 run back to back on chip — nothing but x, done, cost is read from
// HBM and nothing but the requested per-state outputs is written.  Weight-gradient accumulators stay in registers
// across all tiles of the CTA; per-CTA partials are reduced in a fixed order by vhjb_reduce_kernel (deterministic).
//
// Math (reference: controller/vhjb.py:29-60, 201-253, 282-285; reverse pass hand-derived, verified against
// autograd by oracle/vhjb_oracle.py::closed_form_grads):
//   z = wrap(x - xf), h0 = (z - mu)/sd, a1 = h0 W1, a2 = s(a1) W2, y = s(a2) W3, V = |y|^2 + eps_s |z|^2
//   b2 = 2y W3^T, b1 = (b2 s'(a2)) W2^T, g0 = (b1 s'(a1)) W1^T, p = g0/sd + 2 eps_s z
//   c = G^T p, u = clip(uf - R^-1 c / 2), xdot = f + G u, vdot = p.xdot, l = z^T Q z + du^T R du, r = vdot/(l+eps) + 1
#pragma once
#include "systems.cuh"

namespace hjb {

constexpr int VH1 = 128, VH2 = 128, VH3 = 64;
constexpr int VBM = 32;   // states per tile: lane <-> state
constexpr int VLD = 33;   // leading dimension of the feature-major activation arrays (conflict-free both ways)
constexpr int VTHREADS = 256;

struct VhjbArgs {
  DevSys sys;  // aoff = 0
  const float* params;
  float mean[HJB_MAX_N], inv_std[HJB_MAX_N], xf[HJB_MAX_N];
  float eps_s;
  float Q[HJB_MAX_N * HJB_MAX_N], R[HJB_MAX_M * HJB_MAX_M], Rsym[HJB_MAX_M * HJB_MAX_M], Rinv[HJB_MAX_M * HJB_MAX_M];
  float uf[HJB_MAX_M];
  float eps;
  const float* xs;
  const float* dones;
  const float* costs;
  int64_t B;
  const float* norm;  // [2] device: sum(1-done)+eps, sum(done)+eps (GRAD only)
  float reg;
  float* V;
  float* p;
  float* u;
  float* r;
  float* partial;     // [gridDim.x][pstride]: per-CTA gradient partials followed by the two loss sums
  int64_t pstride;
  int64_t n_tiles;
  long long* dbg;     // developer timing probe (HJB_TC_DEBUG_TIMING=1), else null
  // streamed batches (tensor-core gradient kernel only): the states / costs of tile t may be read once
  // ready[t / piece_tiles] != 0 — the flags are written by the copy stream behind each piece of the batch
  const int* ready;
  int64_t piece_tiles;
  unsigned poll_limit;  // polls (2 us apart) before a wait for a piece gives up: 2^22 (~8 s); HJB_STREAM_POLL_LIMIT overrides
  // Deferred states (tensor-core gradient kernel -> fp32 pass): a state whose adjoint seeds lie beyond the fp16 range
  // management of vhjb_tc.cuh contributes NOTHING on the tensor path; its index is appended to the list of the (CTA,
  // epilogue warp) that met it — defer_index[list][0 .. defer_count[list]) in tile order, so the lists are the same in
  // every run — and the CUDA-core kernel, launched behind the tensor kernel with defer_gather = 1, runs exactly those
  // states in fp32 and adds their weight gradients through its own per-CTA partial slots (part_slot0 onwards).
  int* defer_count;     // [defer_lists]
  int* defer_index;     // [defer_lists][kDeferCap]
  int defer_lists;      // lists in use = 4 x (CTAs of the tensor launch)
  int defer_gather;     // CUDA-core kernel: 1 = the batch is the concatenation of the deferred lists
  int part_slot0;       // first per-CTA partial slot of this launch
  float* tail;          // workspace tail words: [0] saturation (launch), [1] saturation (total), [2] stream failures,
                        // [3] CTAs of the deferred pass that wrote partials, [4] deferred states of the last launch,
                        // [5] 1 = the deferred pass ran the whole batch (the tensor launch's partials are void)
};
constexpr int kDeferCap = 2048;   // entries per list; a list that is full counts further states as saturated (not deferred)

__host__ __device__ constexpr int vhjb_param_count(int n) { return n * VH1 + VH1 * VH2 + VH2 * VH3; }
__host__ __device__ constexpr int vhjb_smem_floats(int n) {
  return vhjb_param_count(n) + 6 * VH1 * VLD + 2 * VH3 * VLD + 3 * n * VLD + VBM + 640;   // + prefix sums of the deferred lists
}

template <int ACT>
__device__ __forceinline__ float act_f(float a) {
  if constexpr (ACT == HJB_ACT_RELU) return fmaxf(a, 0.f);
  else if constexpr (ACT == HJB_ACT_TANH) return tanhf(a);
  else return sinf(a);
}
template <int ACT>
__device__ __forceinline__ float act_d1(float a) {
  if constexpr (ACT == HJB_ACT_RELU) return a > 0.f ? 1.f : 0.f;
  else if constexpr (ACT == HJB_ACT_TANH) { const float t = tanhf(a); return fmaf(-t, t, 1.f); }
  else return cosf(a);
}
template <int ACT>
__device__ __forceinline__ float act_d2(float a) {
  if constexpr (ACT == HJB_ACT_RELU) return 0.f;
  else if constexpr (ACT == HJB_ACT_TANH) { const float t = tanhf(a); return -2.f * t * fmaf(-t, t, 1.f); }
  else return -sinf(a);
}
__device__ __forceinline__ float sign0(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// out(j, r) = sum_k A(k, r) W[k][j], W row-major [K][NOUT]; warp w owns outputs j in [w NOUT/8, (w+1) NOUT/8)
template <int K, int NOUT, class AF, class EF>
__device__ __forceinline__ void gemm_fwd(const float* __restrict__ W, AF a_of, EF epi, int warp, int lane) {
  constexpr int JW = NOUT / 8;
  static_assert(JW % 4 == 0, "vector weight loads");
  float acc[JW];
#pragma unroll
  for (int q = 0; q < JW; ++q) acc[q] = 0.f;
  const float* wrow = W + warp * JW;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float av = a_of(k, lane);
#pragma unroll
    for (int q = 0; q < JW; q += 4) {
      const float4 w = *reinterpret_cast<const float4*>(wrow + k * NOUT + q);
      acc[q] = fmaf(av, w.x, acc[q]);
      acc[q + 1] = fmaf(av, w.y, acc[q + 1]);
      acc[q + 2] = fmaf(av, w.z, acc[q + 2]);
      acc[q + 3] = fmaf(av, w.w, acc[q + 3]);
    }
  }
#pragma unroll
  for (int q = 0; q < JW; ++q) epi(warp * JW + q, lane, acc[q]);
}

// out(i, r) = sum_o A(o, r) W[i][o], W row-major [NOUT][KO] (i.e. the transposed use of a weight matrix)
template <int KO, int NOUT, class AF, class EF>
__device__ __forceinline__ void gemm_bwd(const float* __restrict__ W, AF a_of, EF epi, int warp, int lane) {
  constexpr bool kWide = (NOUT % 8 == 0);
  constexpr int JW = kWide ? NOUT / 8 : (NOUT + 7) / 8;
  float acc[JW];
  int row[JW];
#pragma unroll
  for (int q = 0; q < JW; ++q) {
    acc[q] = 0.f;
    row[q] = kWide ? warp * JW + q : warp + 8 * q;
  }
#pragma unroll 2
  for (int o = 0; o < KO; o += 4) {
    const float a0 = a_of(o, lane), a1 = a_of(o + 1, lane), a2 = a_of(o + 2, lane), a3 = a_of(o + 3, lane);
#pragma unroll
    for (int q = 0; q < JW; ++q) {
      const int i = row[q] < NOUT ? row[q] : NOUT - 1;
      const float4 w = *reinterpret_cast<const float4*>(W + i * KO + o);
      acc[q] = fmaf(a0, w.x, acc[q]);
      acc[q] = fmaf(a1, w.y, acc[q]);
      acc[q] = fmaf(a2, w.z, acc[q]);
      acc[q] = fmaf(a3, w.w, acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < JW; ++q)
    if (row[q] < NOUT) epi(row[q], lane, acc[q]);
}

// acc[a][b] += sum_r X(ti + 16 a, r) Y(to + 16 b, r)
template <int NA, int NB, class XF, class YF>
__device__ __forceinline__ void wgrad(float (&acc)[NA][NB], XF x_of, YF y_of, int ti, int to) {
#pragma unroll 2
  for (int r = 0; r < VBM; ++r) {
    float x[NA], y[NB];
#pragma unroll
    for (int a = 0; a < NA; ++a) x[a] = x_of(ti + 16 * a, r);
#pragma unroll
    for (int b = 0; b < NB; ++b) y[b] = y_of(to + 16 * b, r);
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[a][b] = fmaf(x[a], y[b], acc[a][b]);
  }
}

template <class S, int ACT, int UFORM, int RFORM, bool GRAD>
__global__ void __launch_bounds__(VTHREADS, 1) vhjb_kernel(const __grid_constant__ VhjbArgs a) {
  constexpr int N = S::N, M = S::M;
  constexpr int NA1 = (N + 1) / 2;  // rows of W1-bar per thread
  extern __shared__ __align__(16) float smem[];
  float* sW1 = smem;
  float* sW2 = sW1 + N * VH1;
  float* sW3 = sW2 + VH1 * VH2;
  float* sA1 = sW3 + VH2 * VH3;
  float* sA2 = sA1 + VH1 * VLD;
  float* sB1 = sA2 + VH2 * VLD;
  float* sB2 = sB1 + VH1 * VLD;
  float* sT1 = sB2 + VH2 * VLD;
  float* sT2 = sT1 + VH1 * VLD;
  float* sY = sT2 + VH2 * VLD;
  float* sYb = sY + VH3 * VLD;
  float* sH0 = sYb + VH3 * VLD;
  float* sZ = sH0 + N * VLD;
  float* sP = sZ + N * VLD;
  float* sVb = sP + N * VLD;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ti = tid >> 4, to = tid & 15;        // 16 x 16 thread grid of the weight-gradient blocks
  const int o1 = tid & 127, ih1 = tid >> 7;      // W1-bar: column o1, rows ih1 + 2a

  // ---- deferred-state mode: the batch is the concatenation of the tensor kernel's deferred lists ----
  int* sPre = reinterpret_cast<int*>(sVb + VBM);     // [defer_lists + 1] exclusive prefix sums of the list lengths
  int64_t n_tiles = a.n_tiles;
  int64_t n_states = a.B;
  bool redo_all = false;
  if constexpr (GRAD) {
    if (a.defer_gather) {
      // The tensor launch counts (partial[cta][P + 2]) what its fp16 range management could not hold even so: an adjoint
      // chain that reached the fp16 ceiling behind in-range seeds, a deferred list that was full.  Any such event and this
      // pass runs the WHOLE batch in fp32; tail[5] tells the reductions to leave the tensor launch's partials out.  Rare
      // (a trained net with a large backward gain and states next to the goal), exact, decided on the device.
      __shared__ int sRedo;
      if (tid == 0) sRedo = 0;
      __syncthreads();
      for (int c = tid; c < a.defer_lists / 4; c += VTHREADS)
        if (__ldcg(a.partial + (int64_t)c * a.pstride + vhjb_param_count(N) + 2) != 0.f) sRedo = 1;
      if (warp == 0) {
        int run = 0;
        for (int base = 0; base < a.defer_lists; base += 32) {
          const int l = base + lane;
          const int c = l < a.defer_lists ? min(__ldcg(a.defer_count + l), kDeferCap) : 0;
          int inc = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
          }
          if (l < a.defer_lists) sPre[l] = run + inc - c;
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) sPre[a.defer_lists] = run;
      }
      __syncthreads();
      redo_all = sRedo != 0;
      n_states = redo_all ? a.B : (int64_t)sPre[a.defer_lists];
      n_tiles = (n_states + VBM - 1) / VBM;
      if (blockIdx.x == 0 && tid == 0) {
        a.tail[3] = (float)min((int64_t)gridDim.x, n_tiles);
        a.tail[4] = (float)n_states;                      // hjb_vhjb_deferred: states of this launch that took the fp32 pass
        a.tail[5] = redo_all ? 1.f : 0.f;
      }
      if ((int64_t)blockIdx.x >= n_tiles) return;      // (whole CTA: nothing to do, no partial written, none read)
    }
  }
  {  // weights -> shared memory, once per CTA
    constexpr int P4 = vhjb_param_count(N) / 4;
    const float4* src = reinterpret_cast<const float4*>(a.params);
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < P4; i += VTHREADS) dst[i] = __ldg(src + i);
  }

  float acc1[NA1];
  float acc2[8][8];
  float acc3[8][4];
  if constexpr (GRAD) {
#pragma unroll
    for (int i = 0; i < NA1; ++i) acc1[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc2[i][j] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc3[i][j] = 0.f;
    }
  }
  float hjb_sum = 0.f, term_sum = 0.f;  // warp 0 only
  float inv_norm0 = 0.f, inv_norm1 = 0.f;
  if constexpr (GRAD) {
    // (eps = 0 with an all-done or all-interior shard: norm = 0 -> weight 0, not 0 * inf = NaN on the masked terms)
    const float nm0 = __ldg(a.norm), nm1 = __ldg(a.norm + 1);
    inv_norm0 = nm0 > 0.f ? 1.0f / nm0 : 0.f;
    inv_norm1 = nm1 > 0.f ? 1.0f / nm1 : 0.f;
  }
  __syncthreads();

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int64_t idx = tile * VBM + lane;   // the state this lane owns
    const bool valid = idx < n_states;
    if constexpr (GRAD) {
      if (a.defer_gather && !redo_all && warp == 0) {   // entry idx of the concatenated lists: binary search over the prefix sums
        int lo = 0, hi = a.defer_lists;
        const int j = (int)idx;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (sPre[mid] <= j) lo = mid; else hi = mid;
        }
        idx = valid ? (int64_t)__ldcg(a.defer_index + (int64_t)lo * kDeferCap + (j - sPre[lo])) : 0;
      }
    }
    // ---- phase 0: load states, error coordinates, normalised input (vhjb.py:39, :45) ----
    float xraw[N];
    if (warp == 0) {
      if (valid) load_row<N>(a.xs, idx, xraw);
      else {
#pragma unroll
        for (int i = 0; i < N; ++i) xraw[i] = a.xf[i];
      }
      float z[N];
#pragma unroll
      for (int i = 0; i < N; ++i) z[i] = xraw[i] - a.xf[i];
      wrap_state<S>(z);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        sZ[i * VLD + lane] = z[i];
        sH0[i * VLD + lane] = (z[i] - a.mean[i]) * a.inv_std[i];
      }
    }
    __syncthreads();
    // ---- forward (vhjb.py:47-58) ----
    gemm_fwd<N, VH1>(sW1, [&](int k, int r) { return sH0[k * VLD + r]; },
                     [&](int j, int r, float v) { sA1[j * VLD + r] = v; }, warp, lane);
    __syncthreads();
    gemm_fwd<VH1, VH2>(sW2, [&](int k, int r) { return act_f<ACT>(sA1[k * VLD + r]); },
                       [&](int j, int r, float v) { sA2[j * VLD + r] = v; }, warp, lane);
    __syncthreads();
    gemm_fwd<VH2, VH3>(sW3, [&](int k, int r) { return act_f<ACT>(sA2[k * VLD + r]); },
                       [&](int j, int r, float v) { sY[j * VLD + r] = v; }, warp, lane);
    __syncthreads();
    // ---- input gradient (vhjb.py:201-202) ----
    gemm_bwd<VH3, VH2>(sW3, [&](int o, int r) { return 2.f * sY[o * VLD + r]; },
                       [&](int i, int r, float v) { sB2[i * VLD + r] = v; }, warp, lane);
    __syncthreads();
    gemm_bwd<VH2, VH1>(sW2, [&](int o, int r) { return sB2[o * VLD + r] * act_d1<ACT>(sA2[o * VLD + r]); },
                       [&](int i, int r, float v) { sB1[i * VLD + r] = v; }, warp, lane);
    __syncthreads();
    gemm_bwd<VH1, N>(sW1, [&](int o, int r) { return sB1[o * VLD + r] * act_d1<ACT>(sA1[o * VLD + r]); },
                     [&](int i, int r, float v) { sP[i * VLD + r] = v; }, warp, lane);
    __syncthreads();
    // ---- per-state epilogue: control, Hamiltonian residual, adjoint seeds (vhjb.py:204-253) ----
    if (warp == 0) {
      float z[N], p[N];
      float V = 0.f, zz = 0.f;
#pragma unroll
      for (int j = 0; j < VH3; ++j) { const float y = sY[j * VLD + lane]; V = fmaf(y, y, V); }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        z[i] = sZ[i * VLD + lane];
        zz = fmaf(z[i], z[i], zz);
        p[i] = fmaf(sP[i * VLD + lane], a.inv_std[i], 2.f * a.eps_s * z[i]);
      }
      V = fmaf(a.eps_s, zz, V);
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
        if constexpr (UFORM == HJB_U_CLIPPED) {
          float ur = a.uf[k];
#pragma unroll
          for (int j = 0; j < M; ++j) ur = fmaf(-0.5f * a.Rinv[k * M + j], c[j], ur);
          inside[k] = (ur > a.sys.umin[k]) && (ur < a.sys.umax[k]);
          u[k] = clampf(ur, a.sys.umin[k], a.sys.umax[k]);
        } else {
          inside[k] = false;
          u[k] = -sign0(c[k]);
        }
        du[k] = u[k] - a.uf[k];
      }
      float xdot[N], vdot = 0.f;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float s = f[i];
#pragma unroll
        for (int k = 0; k < M; ++k) s = fmaf(G[i * M + k], u[k], s);
        xdot[i] = s;
        vdot = fmaf(p[i], s, vdot);
      }
      const float done = valid ? __ldg(a.dones + idx) : 0.f;
      const float cost = valid ? __ldg(a.costs + idx) : 1.f;
      float r, pbar[N], Vbar = 0.f;
      if constexpr (RFORM == HJB_RES_NORMALIZED) {
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
        const float den = l + a.eps;
        const float iden = 1.0f / den;
        r = fmaf(vdot, iden, 1.f);
        const float tq = V / (cost + a.eps) - 1.f;
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
            for (int j = 0; j < M; ++j) ub = fmaf(lbar * a.Rsym[k * M + j], du[j], ub);
            t[k] = inside[k] ? ub : 0.f;   // u_raw-bar
          }
#pragma unroll
          for (int i = 0; i < N; ++i) pbar[i] = vbar * xdot[i];
#pragma unroll
          for (int j = 0; j < M; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < M; ++k) s = fmaf(t[k], a.Rinv[k * M + j], s);
            s *= -0.5f;
#pragma unroll
            for (int i = 0; i < N; ++i) pbar[i] = fmaf(G[i * M + j], s, pbar[i]);
          }
          Vbar = valid ? a.reg * done * inv_norm1 * sign0(tq) / (cost + a.eps) : 0.f;
        }
      } else {  // MIN_TIME: |vdot + l_i|, plain mean, u = -sign(.) carries no gradient
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
      if constexpr (GRAD) {
#pragma unroll
        for (int i = 0; i < N; ++i) sP[i * VLD + lane] = pbar[i] * a.inv_std[i];   // g0-bar
        sVb[lane] = Vbar;
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
                         [&](int j, int r, float v) {
                           sYb[j * VLD + r] = 2.f * fmaf(sY[j * VLD + r], sVb[r], v);               // y-bar
                         }, warp, lane);
      __syncthreads();
      // weight gradients that need g1-bar / g2-bar before they are overwritten
      wgrad<8, 8>(acc2, [&](int i, int r) { return sT1[i * VLD + r] * act_d1<ACT>(sA1[i * VLD + r]); },
                  [&](int o, int r) { return sB2[o * VLD + r] * act_d1<ACT>(sA2[o * VLD + r]); }, ti, to);
      wgrad<8, 4>(acc3, [&](int i, int r) { return sT2[i * VLD + r] * act_d1<ACT>(sA2[i * VLD + r]); },
                  [&](int o, int r) { return 2.f * sY[o * VLD + r]; }, ti, to);
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
  float* part = a.partial + (int64_t)(blockIdx.x + a.part_slot0) * a.pstride;
  if constexpr (GRAD) {
    // (deferred-state mode: hjb_sum / term_sum are the loss terms of the deferred states — the tensor kernel left them out)
#pragma unroll
    for (int q = 0; q < NA1; ++q) {
      const int i = ih1 + 2 * q;
      if (i < N) part[i * VH1 + o1] = acc1[q];
    }
    float* p2 = part + N * VH1;
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
      for (int y = 0; y < 8; ++y) p2[(ti + 16 * x) * VH2 + (to + 16 * y)] = acc2[x][y];
    float* p3 = p2 + VH1 * VH2;
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) p3[(ti + 16 * x) * VH3 + (to + 16 * y)] = acc3[x][y];
  }
  if (warp == 0) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      hjb_sum += __shfl_xor_sync(0xffffffffu, hjb_sum, s);
      term_sum += __shfl_xor_sync(0xffffffffu, term_sum, s);
    }
    if (lane == 0) {
      part[vhjb_param_count(N)] = hjb_sum;
      part[vhjb_param_count(N) + 1] = term_sum;
      part[vhjb_param_count(N) + 2] = 0.f;   // saturation count (tensor-core kernel only)
    }
  }
}

struct VhjbLaunch {
  int grid;
  bool grad;
};

template <class S, int ACT, int UFORM, int RFORM>
inline cudaError_t launch_vhjb_variant(const VhjbArgs& a, const VhjbLaunch& l, cudaStream_t st) {
  const size_t smem = sizeof(float) * vhjb_smem_floats(S::N);
  cudaError_t e;
  if (l.grad) {
    auto k = vhjb_kernel<S, ACT, UFORM, RFORM, true>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<l.grid, VTHREADS, smem, st>>>(a);
  } else {
    auto k = vhjb_kernel<S, ACT, UFORM, RFORM, false>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<l.grid, VTHREADS, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// NORMALIZED+CLIPPED for every system; MIN_TIME+BANGBANG only where ALLOW_MIN_TIME (linear systems)
template <class S, bool ALLOW_MIN_TIME>
inline cudaError_t launch_vhjb_system(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st) {
  if (uform == HJB_U_CLIPPED && rform == HJB_RES_NORMALIZED) {
    switch (act) {
      case HJB_ACT_RELU: return launch_vhjb_variant<S, HJB_ACT_RELU, HJB_U_CLIPPED, HJB_RES_NORMALIZED>(a, l, st);
      case HJB_ACT_TANH: return launch_vhjb_variant<S, HJB_ACT_TANH, HJB_U_CLIPPED, HJB_RES_NORMALIZED>(a, l, st);
      case HJB_ACT_SIN: return launch_vhjb_variant<S, HJB_ACT_SIN, HJB_U_CLIPPED, HJB_RES_NORMALIZED>(a, l, st);
    }
  }
  if constexpr (ALLOW_MIN_TIME) {
    if (uform == HJB_U_BANGBANG && rform == HJB_RES_MIN_TIME) {
      switch (act) {
        case HJB_ACT_RELU: return launch_vhjb_variant<S, HJB_ACT_RELU, HJB_U_BANGBANG, HJB_RES_MIN_TIME>(a, l, st);
        case HJB_ACT_TANH: return launch_vhjb_variant<S, HJB_ACT_TANH, HJB_U_BANGBANG, HJB_RES_MIN_TIME>(a, l, st);
        case HJB_ACT_SIN: return launch_vhjb_variant<S, HJB_ACT_SIN, HJB_U_BANGBANG, HJB_RES_MIN_TIME>(a, l, st);
      }
    }
  }
  return cudaErrorNotSupported;
}

cudaError_t vhjb_launch_linear21(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st);
cudaError_t vhjb_launch_cartpole(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st);
cudaError_t vhjb_launch_quad2d(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st);
cudaError_t vhjb_launch_quad10d(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st);

}  // namespace hjb

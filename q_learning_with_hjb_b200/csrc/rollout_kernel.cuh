// K1 — batched closed-loop rollout: one thread = one environment, whole horizon in one launch.
//
// State, RK4 stages and the cost accumulator live in registers; system constants, gains and cost
// matrices are read from the kernel parameter space (constant bank operands, no register cost).
// Trajectories are written time-major ([t][env][i]) so that a warp's stores of one step are contiguous;
// stores are vectorised (16 B for n = 4, 8 B for even n) and streaming (st.global.cs); rows of awkward width (n = 10, m = 3)
// are staged per warp in shared memory and written by one TMA bulk store per 32-row block.
#pragma once
#include <type_traits>

#include "systems.cuh"

namespace hjb {

// COST_UNIT: Q = I, R = I and a goal whose non-angle components are 0 (every reference notebook and gin file) — one FFMA per
// state component, FADD + FFMA per input, instead of two FFMAs each
enum CostMode { COST_NONE = 0, COST_DIAG = 1, COST_DENSE = 2, COST_UNIT = 3 };

struct RolloutArgs {
  DevSys sys;
  DevCtl ctl;
  DevCost cost;
  DevBox box;
  const float* x0;
  float* xs;
  float* us;
  float* x_final;
  float* cost_out;
  int32_t* steps_out;
  int64_t N;
  int32_t T;
  int32_t stride;  // record stride (>= 1 when xs/us requested)
  int32_t n_rec;   // T / stride recorded intervals
  int32_t staged;  // trajectory rows may go through shared memory + cp.async.bulk (alignment checked by the launcher)
};

// error coordinate of internal state z w.r.t. a goal whose non-angle part is xf and whose angle part differs
// from the internal offset by dang: d_i = z_i - xf_i, d_th = wrap(z_th + dang) (no wrap needed when dang == 0)
template <class S>
__device__ __forceinline__ void error_coords(const float* z, const float* xf, const float* dang, float* d) {
#pragma unroll
  for (int i = 0; i < S::N; ++i) d[i] = z[i] - xf[i];
#pragma unroll
  for (int k = 0; k < S::NANG; ++k) {
    const int i = S::ang(k);
    d[i] = z[i];
    if (dang[k] != 0.f) d[i] = wrap_pi_<S::kFast>(z[i] + dang[k]);
  }
}

// l(x, u) = dx^T Q dx + (u - uf)^T R (u - uf)
// CWRAP: some angle of the cost's goal differs from the internal offset (a wrap per step); the loop-invariant test is made
// once per launch, outside the step loop
// c0r / r0r: the additive constants of the diagonal form, held in registers by the caller (an FFMA takes ONE constant-bank
// operand; left to itself the compiler re-loads the second one with an LDC inside the step loop — 12 of the 112
// instructions of a 10-D quadcopter step)
template <class S, int COST, bool CWRAP>
__device__ __forceinline__ float running_cost(const DevCost& pc, const float* c0r, const float* r0r, const float* z,
                                              const float* u, float l) {
  if constexpr (COST == COST_UNIT) {
#pragma unroll
    for (int i = 0; i < S::N; ++i) {
      float d = z[i];
#pragma unroll
      for (int k = 0; k < S::NANG; ++k)
        if (CWRAP && S::ang(k) == i && pc.dang[k] != 0.f) d = wrap_pi_<S::kFast>(z[i] + pc.dang[k]);
      l = fmaf(d, d, l);
    }
#pragma unroll
    for (int k = 0; k < S::M; ++k) {
      const float y = u[k] + pc.r0[k];                       // r0 = -uf
      l = fmaf(y, y, l);
    }
  } else if constexpr (COST == COST_DIAG) {
#pragma unroll
    for (int i = 0; i < S::N; ++i) {
      bool is_ang = false;
#pragma unroll
      for (int k = 0; k < S::NANG; ++k) is_ang = is_ang || (S::ang(k) == i);
      float y;
      if (is_ang) {
        float d = z[i];
#pragma unroll
        for (int k = 0; k < S::NANG; ++k)
          if (CWRAP && S::ang(k) == i && pc.dang[k] != 0.f) d = wrap_pi_<S::kFast>(z[i] + pc.dang[k]);
        y = d * pc.sq[i];
      } else {
        y = fmaf(z[i], pc.sq[i], c0r[i]);
      }
      l = fmaf(y, y, l);
    }
#pragma unroll
    for (int k = 0; k < S::M; ++k) {
      const float y = fmaf(u[k], pc.sr[k], r0r[k]);
      l = fmaf(y, y, l);
    }
  } else {
    float dx[S::N], du[S::M];
    if constexpr (CWRAP) error_coords<S>(z, pc.xf, pc.dang, dx);
    else {
#pragma unroll
      for (int i = 0; i < S::N; ++i) dx[i] = z[i] - pc.xf[i];
#pragma unroll
      for (int k = 0; k < S::NANG; ++k) dx[S::ang(k)] = z[S::ang(k)];
    }
#pragma unroll
    for (int k = 0; k < S::M; ++k) du[k] = u[k] - pc.uf[k];
#pragma unroll
    for (int i = 0; i < S::N; ++i) {
      float row = 0.f;
#pragma unroll
      for (int j = 0; j < S::N; ++j) row = fmaf(pc.Q[i * S::N + j], dx[j], row);
      l = fmaf(dx[i], row, l);
    }
#pragma unroll
    for (int k = 0; k < S::M; ++k) {
      float row = 0.f;
#pragma unroll
      for (int j = 0; j < S::M; ++j) row = fmaf(pc.R[k * S::M + j], du[j], row);
      l = fmaf(du[k], row, l);
    }
  }
  return l;
}

// one integration step of Dynamics.simulate AFTER the clip, on the internal state: z <- wrap(step(z, u))
template <class S, int INTEG, class TC>
__device__ __forceinline__ void integrate(const DevSys& ps, float* x, const typename S::Trig& tr0, const float* u, const TC& tc) {
  constexpr int N = S::N;
  if constexpr (INTEG == HJB_INT_DISCRETE) {
    float d[N];
    S::xdot(ps, x, tr0, u, d);
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = d[i];
  } else if constexpr (INTEG == HJB_INT_EULER) {
    if constexpr (S::kFusedEuler) {
      S::euler(ps, x, tr0, u);          // dt folded into the system's constants (systems.cuh)
    } else {
      float d[N];
      S::xdot(ps, x, tr0, u, d);
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] = fmaf(d[i], ps.dt, x[i]);
    }
  } else {
    const float h = ps.dt, hh = 0.5f * ps.dt;
    float k[N], acc[N], xt[N];
    typename S::Trig tr;
    S::xdot(ps, x, tr0, u, k);
#pragma unroll
    for (int i = 0; i < N; ++i) { acc[i] = k[i]; xt[i] = fmaf(hh, k[i], x[i]); }
    // (stage states are not wrapped: GUARD = true sends arguments beyond the trig table's range to the in-line path)
    S::template trig<true>(ps, xt, tr, tc);
    S::xdot(ps, xt, tr, u, k);
#pragma unroll
    for (int i = 0; i < N; ++i) { acc[i] = fmaf(2.f, k[i], acc[i]); xt[i] = fmaf(hh, k[i], x[i]); }
    S::template trig<true>(ps, xt, tr, tc);
    S::xdot(ps, xt, tr, u, k);
#pragma unroll
    for (int i = 0; i < N; ++i) { acc[i] = fmaf(2.f, k[i], acc[i]); xt[i] = fmaf(h, k[i], x[i]); }
    S::template trig<true>(ps, xt, tr, tc);
    S::xdot(ps, xt, tr, u, k);
    const float h6 = ps.dt * (1.0f / 6.0f);
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = fmaf(h6, acc[i] + k[i], x[i]);
  }
  wrap_state<S>(x);
}

// ---- trajectory write-back through shared memory + TMA bulk stores ----------------------------------------------
// A warp's 32 rows of one recorded time slice are contiguous in global memory (time-major layout): 128 W bytes.  Each
// lane writes its row into the warp's staging buffer and ONE lane issues a single cp.async.bulk (shared -> global) for
// the whole block, so the memory system sees full lines whatever the row width (per-thread stores of odd widths —
// n = 10 is 5 x 8 bytes at a 40-byte stride, m = 3 is 3 x 4 bytes at 12 — touch every 32-byte sector several times:
// measured 35 % of the HBM rate for the 10-D quadcopter).  Two buffers per stream of rows; before a buffer is
// rewritten the issuing lane waits until at most one bulk group is still reading.
__device__ __forceinline__ void bulk_store_issue(float* gdst, const float* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int W>
__device__ __forceinline__ void staged_store(float* stage, int lane, float* gblock, const float* v) {
  if (lane == 0) bulk_store_wait_read_1();      // the bulk store that last read this buffer is older than the latest one
  __syncwarp();
  float* row = stage + lane * W;
  if constexpr (W % 4 == 0) {
#pragma unroll
    for (int i = 0; i < W / 4; ++i) reinterpret_cast<float4*>(row)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if constexpr (W % 2 == 0) {
#pragma unroll
    for (int i = 0; i < W / 2; ++i) reinterpret_cast<float2*>(row)[i] = make_float2(v[2 * i], v[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < W; ++i) row[i] = v[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) bulk_store_issue(gblock, stage, 32u * W * 4u);
}

template <class S>
__device__ __forceinline__ bool inside_box(const DevBox& b, const float* z) {
  float dx[S::N];
  error_coords<S>(z, b.xf, b.dang, dx);
  bool in = true;
#pragma unroll
  for (int i = 0; i < S::N; ++i) in = in && !(dx[i] > b.hi[i]) && !(dx[i] < b.lo[i]);
  return in;
}

// (sin, cos)(k 2^-7 + aoff) in double, rounded to fp32, for the `nt` angle arguments of a system.  One copy per translation
// unit: fp64 sincos carries a large slow path.
// `wide`: entries (S, C, -S/2, -C/2) as float4 (TableTrigT<true>), else (S, C) as float2.
static __device__ __noinline__ void build_trig_tables(void* tab, bool wide, int nt, float aoff0, float aoff1) {
  for (int i = threadIdx.x; i < nt * kTrigSize; i += blockDim.x) {
    const int k = i / kTrigSize, j = i - k * kTrigSize;
    const double off = k == 2 ? (double)aoff0 + (double)aoff1 : (double)(k == 1 ? aoff1 : aoff0);
    double sv, cv;
    sincos((double)(j - kTrigHalf) * (1.0 / (double)(1 << kTrigLog2)) + off, &sv, &cv);
    const float sf = (float)sv, cf = (float)cv;
    if (wide) static_cast<float4*>(tab)[i] = make_float4(sf, cf, -0.5f * sf, -0.5f * cf);
    else static_cast<float2*>(tab)[i] = make_float2(sf, cf);
  }
}

// Persistent CTAs: the grid is at most kRolloutCtasPerSm x SM count and every CTA walks over blocks of 256 environments
// (block b, b + gridDim.x, ...).  What a CTA sets up once — the trig tables in shared memory — is then paid ~1200 times
// per launch instead of once per 256 environments.
constexpr int kRolloutCtasPerSm = 8;

template <class S, class C, int INTEG, bool REC, int COST, bool BOX>
__global__ void __launch_bounds__(256) rollout_kernel(const __grid_constant__ RolloutArgs a) {
  constexpr int N = S::N, M = S::M;
  // recorded rows go through the warp's staging buffers when the whole warp is in range and every time slice starts
  // on a 16-byte boundary (a.staged: N W divisible by 4 floats, 16-byte aligned bases); otherwise per-thread stores
  // Measured on B200 (bench.py --record-stride 1, staged against direct): n = 10 / m = 3 rows 6248 against 2399 GB/s;
  // n = 6 rows +4 %; n = 4 / m = 1 rows (already 16-byte stores) lose to the two proxy fences per step (3894 against
  // 6439 GB/s).  So only rows whose width is neither a vector width nor small are staged: the 10-D quadcopter's.
  constexpr bool kStageX = REC && (N % 4 != 0) && (N > 6);
  constexpr bool kStageU = REC && (M % 2 != 0) && (M > 1);
  __shared__ __align__(128) float stage_x[kStageX ? 8 * 2 * 32 * N : 1];
  __shared__ __align__(128) float stage_u[kStageU ? 8 * 2 * 32 * M : 1];
  // fast instantiations of systems with angles: sin / cos from per-CTA tables (hjb_common.cuh::sincos_tab)
  constexpr bool kTab = S::kFast && S::NANG > 0;
  constexpr int NT = kTab ? trig_tables<S>() : 0;
  // wide entries (one FMA-pipe instruction less per evaluation) for the systems with ONE table (cart-pole, quad-2D: 14 KB x
  // 8 CTAs per SM); two wide tables would not fit 8 times per SM (acrobot: 230 KB) or next to the staging buffers of the
  // recorded variants within the 48 KB of static shared memory (10-D quadcopter)
  // (... and only where the loop is bound by its instructions: the recorded variants are bound by their stores to HBM and
  // lose 4-5 % with the larger shared-memory carve-out — C4 with every step recorded: 1.55e11 -> 1.48e11 env-steps/s)
  constexpr bool kWide = kTab && NT == 1 && !REC;
  using TabTrig = TableTrigT<kWide>;
  __shared__ __align__(16) typename TabTrig::Entry trig_tab[kTab ? NT * kTrigSize : 1];
  using TC = std::conditional_t<kTab, TabTrig, DirectTrig<S::kFast>>;
  TC tc;
  if constexpr (kTab) {
    build_trig_tables(trig_tab, kWide, NT, a.sys.aoff[0], a.sys.aoff[1]);
    __syncthreads();
    tc.tab = trig_tab;
  }
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;

  // (+ a zero that only the launch knows: a plain copy of a kernel parameter is re-loaded by ptxas wherever it is used)
  float c0r[N], r0r[M];
  // Measured: worth it for the 13 terms of the 10-D quadcopter (112 -> 103 instructions per step, 2.91e11 -> 3.10e11
  // env-steps/s); for the narrower systems the compiler keeps the constants in registers by itself and the detour costs
  // ~3 % (C4), so those read the parameters directly.
  constexpr bool kRegConsts = COST == COST_DIAG && (N + M) > 8;
  const float opaque0 = kRegConsts && a.T < 0 ? 1.f : 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) c0r[i] = COST == COST_DIAG ? (kRegConsts ? a.cost.c0[i] + opaque0 : a.cost.c0[i]) : 0.f;
#pragma unroll
  for (int k = 0; k < M; ++k) r0r[k] = COST == COST_DIAG ? (kRegConsts ? a.cost.r0[k] + opaque0 : a.cost.r0[k]) : 0.f;
  bool cost_wraps = false;
  if constexpr (COST != COST_NONE) {
#pragma unroll
    for (int k = 0; k < S::NANG; ++k) cost_wraps = cost_wraps || (a.cost.dang[k] != 0.f);
  }

  for (int64_t blk = blockIdx.x; blk * 256 < a.N; blk += gridDim.x) {
  const int64_t env = blk * 256 + threadIdx.x;
  const int64_t env0 = env - lane;
  bool staged = false;
  if constexpr (kStageX || kStageU) staged = a.staged != 0 && env0 + 32 <= a.N;
  if (env >= a.N) break;    // the ragged tail of the last block (no block-wide barrier below this point)

  float z[N];  // internal state
  {
    float x[N];
    load_row<N>(a.x0, env, x);
    if constexpr (REC) {
      if (a.xs) store_row<N>(a.xs, env, x);
    }
    to_internal<S>(a.sys, x, z);
  }
  float J = 0.f;  // sum of l (times dt at the end)
  int32_t nsteps = 0;
  bool alive = true;
  int64_t rec = 1;      // next trajectory slot
  int32_t phase = 0;    // steps since the last recorded state

  auto run = [&](auto cwrap) {
  constexpr bool CWRAP = decltype(cwrap)::value;
  // the plain step loop is issue-bound: four steps per trip cost 54.3 instead of 58 instructions per quad-2D Euler step
  // (loop control, and the rounding of a / 2 pi shared between a step's wrap and the next step's table index); recorded
  // and boxed rollouts are bound by their stores and stay rolled
  constexpr int kUnroll = (REC || BOX) ? 1 : 4;
#pragma unroll kUnroll
  for (int32_t t = 0; t < a.T; ++t) {
    if constexpr (BOX) alive = alive && inside_box<S>(a.box, z);
    typename S::Trig tr;
    S::template trig<false>(a.sys, z, tr, tc);
    float u[M];
    C::template control<S>(a.sys, a.ctl, z, tr, u, t);
    if constexpr (REC) {
      if (phase == 0 && a.us && rec <= a.n_rec) {
        if (kStageU && staged) staged_store<M>(stage_u + (wib * 2 + (int)(rec & 1)) * 32 * M, lane, a.us + ((rec - 1) * a.N + env0) * M, u);
        else store_row<M>(a.us, (rec - 1) * a.N + env, u);
      }
    }
    if constexpr (BOX) {
      float zn[N];
#pragma unroll
      for (int i = 0; i < N; ++i) zn[i] = z[i];
      float l = 0.f;
      if constexpr (COST != COST_NONE) l = running_cost<S, COST, CWRAP>(a.cost, c0r, r0r, z, u, 0.f);
      if constexpr (!C::kClips) clip_u<S>(a.sys, u);
      integrate<S, INTEG>(a.sys, zn, tr, u, tc);
      if (alive) {
#pragma unroll
        for (int i = 0; i < N; ++i) z[i] = zn[i];
        J += l;
        ++nsteps;
      }
    } else {
      if constexpr (COST != COST_NONE) J = running_cost<S, COST, CWRAP>(a.cost, c0r, r0r, z, u, J);
      // Dynamics.simulate's own clip (dynamics_basic.py:118); idempotent when the controller already clipped
      if constexpr (!C::kClips) clip_u<S>(a.sys, u);
      integrate<S, INTEG>(a.sys, z, tr, u, tc);
    }
    if constexpr (REC) {
      if (++phase == a.stride) {
        phase = 0;
        if (a.xs) {
          float x[N];
          to_external<S>(a.sys, z, x);
          if (kStageX && staged) staged_store<N>(stage_x + (wib * 2 + (int)(rec & 1)) * 32 * N, lane, a.xs + (rec * a.N + env0) * N, x);
          else store_row<N>(a.xs, rec * a.N + env, x);
        }
        ++rec;
      }
    }
  }
  };
  if (COST != COST_NONE && cost_wraps) run(std::true_type{});
  else run(std::false_type{});
  if constexpr (kStageX || kStageU) {
    if (staged && lane == 0) bulk_store_wait_all();   // the staging buffers must outlive the bulk stores that read them
  }
  if (a.x_final) {
    float x[N];
    to_external<S>(a.sys, z, x);
    store_row<N>(a.x_final, env, x);
  }
  if (a.cost_out) a.cost_out[env] = J * a.sys.dt;
  if (a.steps_out) a.steps_out[env] = BOX ? nsteps : a.T;
  }
}

// ------------------------------------------------------------------------------------------------
// launchers (one translation unit per problem instantiates these)
// ------------------------------------------------------------------------------------------------
inline int rollout_sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
      cached = sms;
  }
  return cached > 0 ? cached : 148;
}

struct RolloutVariant {
  int integrator;  // hjb_integrator
  bool rec;
  int cost;        // CostMode
  bool box;
};

template <class S, class C, int INTEG, bool REC, int COST, bool BOX>
inline cudaError_t launch_one(const RolloutArgs& a, cudaStream_t st) {
  const int block = 256;
  int64_t grid = (a.N + block - 1) / block;
  const int64_t resident = (int64_t)kRolloutCtasPerSm * rollout_sm_count();
  if (grid > resident) grid = resident;
  rollout_kernel<S, C, INTEG, REC, COST, BOX><<<(unsigned)grid, block, 0, st>>>(a);
  return cudaGetLastError();
}

// Compiled combinations.  The specialised cost forms matter where the kernel is compute-bound and common: without a box,
//   final state (+ cost):   NONE, UNIT, DIAG, DENSE
//   recorded trajectories:  NONE, UNIT, DENSE        (diagonal costs run as DENSE)
//   box termination:        DENSE only               (every cost form is a dense (Q, R); without a cost output the sum is
//                                                     computed and dropped) — rollouts of this kind are short
// — 9 instantiations per (system, controller, trig, integrator) instead of 16.
template <class S, class C, int INTEG, bool REC>
inline cudaError_t launch_cost(const RolloutArgs& a, const RolloutVariant& v, cudaStream_t st) {
  if (v.box) return launch_one<S, C, INTEG, REC, COST_DENSE, true>(a, st);
  switch (v.cost) {
    case COST_NONE: return launch_one<S, C, INTEG, REC, COST_NONE, false>(a, st);
    case COST_UNIT: return launch_one<S, C, INTEG, REC, COST_UNIT, false>(a, st);
    case COST_DIAG:
      if constexpr (!REC) return launch_one<S, C, INTEG, REC, COST_DIAG, false>(a, st);
      else return launch_one<S, C, INTEG, REC, COST_DENSE, false>(a, st);
    default: return launch_one<S, C, INTEG, REC, COST_DENSE, false>(a, st);
  }
}
template <class S, class C, int INTEG>
inline cudaError_t launch_rec(const RolloutArgs& a, const RolloutVariant& v, cudaStream_t st) {
  return v.rec ? launch_cost<S, C, INTEG, true>(a, v, st) : launch_cost<S, C, INTEG, false>(a, v, st);
}
// ALLOW_DISCRETE: only LINEAR systems have the exact-ZOH update
template <class S, class C, bool ALLOW_DISCRETE>
inline cudaError_t launch_rollout(const RolloutArgs& a, const RolloutVariant& v, cudaStream_t st) {
  switch (v.integrator) {
    case HJB_INT_EULER: return launch_rec<S, C, HJB_INT_EULER>(a, v, st);
    case HJB_INT_RK4: return launch_rec<S, C, HJB_INT_RK4>(a, v, st);
    case HJB_INT_DISCRETE:
      if constexpr (ALLOW_DISCRETE) return launch_rec<S, C, HJB_INT_DISCRETE>(a, v, st);
      else return cudaErrorNotSupported;
    default: return cudaErrorInvalidValue;
  }
}

// problem ids dispatched by api.cu; each is defined in its own rollout_<name>.cu
#define HJB_DECLARE_PROBLEM(name) \
  cudaError_t rollout_##name(const RolloutArgs& a, const RolloutVariant& v, bool fast, cudaStream_t st)

HJB_DECLARE_PROBLEM(linear21_fb);
HJB_DECLARE_PROBLEM(linear41_fb);
HJB_DECLARE_PROBLEM(linear42_fb);
HJB_DECLARE_PROBLEM(cartpole_fb);
HJB_DECLARE_PROBLEM(cartpole_es);
HJB_DECLARE_PROBLEM(acrobot_fb);
HJB_DECLARE_PROBLEM(acrobot_es);
HJB_DECLARE_PROBLEM(quad2d_fb);
HJB_DECLARE_PROBLEM(quad2d_track);
HJB_DECLARE_PROBLEM(quad10d_fb);

#define HJB_DEFINE_PROBLEM(name, SYS_T, CTL, ALLOW_DISCRETE)                                              \
  cudaError_t rollout_##name(const RolloutArgs& a, const RolloutVariant& v, bool fast, cudaStream_t st) { \
    if (fast) return launch_rollout<SYS_T(true), CTL, ALLOW_DISCRETE>(a, v, st);                          \
    return launch_rollout<SYS_T(false), CTL, ALLOW_DISCRETE>(a, v, st);                                   \
  }
// state-feedback problems: the controller's clip flag selects a compile-time variant
#define HJB_DEFINE_FB_PROBLEM(name, SYS_T, ALLOW_DISCRETE)                                                \
  cudaError_t rollout_##name(const RolloutArgs& a, const RolloutVariant& v, bool fast, cudaStream_t st) { \
    if (a.ctl.clip) {                                                                                     \
      if (fast) return launch_rollout<SYS_T(true), FeedbackCtl<true>, ALLOW_DISCRETE>(a, v, st);          \
      return launch_rollout<SYS_T(false), FeedbackCtl<true>, ALLOW_DISCRETE>(a, v, st);                   \
    }                                                                                                     \
    if (fast) return launch_rollout<SYS_T(true), FeedbackCtl<false>, ALLOW_DISCRETE>(a, v, st);           \
    return launch_rollout<SYS_T(false), FeedbackCtl<false>, ALLOW_DISCRETE>(a, v, st);                    \
  }

}  // namespace hjb

// vhjb entry points: argument folding, launch of the fused pass, deterministic cross-CTA reduction, done-counting
// and the Adam update.  Kernel bodies: vhjb_simt.cuh; per-system instantiations: vhjb_<system>.cu.
#include <cmath>
#include <cstring>

#include <cstdio>
#include <cstdlib>

#include "vhjb_simt.cuh"
#include "vhjb_tc.cuh"

namespace hjb {

constexpr int kMaxCtas = 160;  // >= SM count (148 on B200): per-CTA partial slots of the main launch
// Workspace: [kMaxCtas slots: main launch][kMaxCtas slots: fp32 pass over the deferred states][4 tail words]
//            (tail: 8 floats)  [int defer_count[4 kMaxCtas]][int defer_index[4 kMaxCtas][kDeferCap]]
constexpr int kSlots = 2 * kMaxCtas;
constexpr int kDeferLists = 4 * kMaxCtas;

// Kernel selection: the tcgen05 kernel (vhjb_tc.cuh) for every compiled combination (relu / tanh / sin value nets);
// the CUDA-core kernel (vhjb_simt.cuh) is the fp32 reference implementation on the device.  HJB_VHJB_IMPL=simt forces the CUDA-core kernel (A/B measurements, parity).
static bool use_tensor_path(const hjb_vnet* net, int64_t B) {
  if (B <= 0 || net->impl == 1) return false;
  const char* e = std::getenv("HJB_VHJB_IMPL");
  return !(e && std::strcmp(e, "simt") == 0);
}

static int64_t pstride_of(int n) { return ((int64_t)vhjb_param_count(n) + 2 + 3) / 4 * 4; }

// grad[j] = sum over CTAs (fixed order) of partial[cta][j]; j in [first, first + count)
// (accumulate: out[j] += the sum — a batch processed as several launches, in launch order)
// dtail != null: the fp32 pass over the deferred states ran behind the main launch; its CTAs 0 .. dtail[3] - 1 wrote
// partials into the slots kMaxCtas onwards, summed after the main launch's (same fixed order in every run)
__global__ void __launch_bounds__(256) vhjb_reduce_kernel(const float* __restrict__ partial, int64_t pstride, int ncta,
                                                          int first, int count, float* __restrict__ out, int accumulate,
                                                          const float* __restrict__ dtail) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  float s = 0.f;
  if (dtail != nullptr && dtail[5] != 0.f) ncta = 0;     // the fp32 pass redid the whole batch (vhjb_simt.cuh: redo_all)
  for (int c = 0; c < ncta; ++c) s += partial[(int64_t)c * pstride + first + j];
  if (dtail != nullptr) {
    const int nd = (int)dtail[3];
    for (int c = 0; c < nd; ++c) s += partial[(int64_t)(kMaxCtas + c) * pstride + first + j];
  }
  out[j] = accumulate ? out[j] + s : s;
}

// saturation count of one launch -> tail[0] (this batch; added when the batch arrives in pieces) and tail[1] (running total).
// streamed launches: the number of warps whose wait for a piece of the batch gave up -> tail[2] (sticky until
// hjb_vhjb_stream_failures resets it); when it is non-zero the loss sums of this step are poisoned with NaN and the
// guarded Adam update (hjb_vhjb_adam_guarded) skips the step.
__global__ void vhjb_sat_kernel(const float* __restrict__ partial, int64_t pstride, int ncta, int at, float* __restrict__ tail,
                                int accumulate, int streamed, float* __restrict__ sums, const float* __restrict__ dtail) {
  // one warp: lane l sums the CTAs l, l + 32, ... (counts are small integers: exact in any order), then a shuffle tree
  float s = 0.f, f = 0.f;
  const bool redone = dtail != nullptr && dtail[5] != 0.f;   // what the tensor launch could not hold was recomputed in fp32
  for (int c = threadIdx.x; c < ncta; c += 32) {
    if (!redone) s += partial[(int64_t)c * pstride + at];
    if (streamed) f += partial[(int64_t)c * pstride + at + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    f += __shfl_xor_sync(0xffffffffu, f, o);
  }
  if (threadIdx.x == 0) {
    tail[0] = accumulate ? tail[0] + s : s;
    tail[1] += s;
    if (streamed && f > 0.f) {
      tail[2] += f;
      if (sums) sums[0] = sums[1] = __int_as_float(0x7fc00000);
    }
  }
}

// optax.adam, skipped (weights, mu, nu untouched) while the workspace's stream-failure word is raised
__global__ void __launch_bounds__(256) adam_guarded_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                                                           const float* __restrict__ g, int64_t len, float lr, float b1, float b2,
                                                           float eps, float bc1, float bc2, const float* __restrict__ guard) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len || *guard != 0.f) return;
  const float gi = g[i];
  const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
  const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  const float mhat = mi / bc1, vhat = vi / bc2;
  w[i] = w[i] - lr * mhat / (sqrtf(vhat) + eps);
}

// ---- sum(1 - done), sum(done): two-stage, fixed order ----
__global__ void __launch_bounds__(256) count_stage1(const float* __restrict__ dones, int64_t B, float* __restrict__ part) {
  float s1 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x)
    s1 += __ldg(dones + i);
  __shared__ float sh[8];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s1;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    part[blockIdx.x] = t;
  }
}
// min_time: the plain mean over the batch of the double-integrator notebook (cell 11) — no boundary term, norm[1] = 1
__global__ void count_stage2(const float* __restrict__ part, int nblk, int64_t B, float eps, int min_time,
                             float* __restrict__ norm) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nblk; ++i) t += (double)part[i];
    norm[0] = (float)((double)B - t + (double)eps);
    norm[1] = min_time ? 1.0f : (float)(t + (double)eps);
  }
}

// Small batches (the reference trains with minibatches of 256): one block counts, adds eps and writes norm — one
// launch instead of two.  Done flags are 0 / 1, so the float sums are exact integers whatever the order.
__global__ void __launch_bounds__(1024) count_one_block(const float* __restrict__ dones, int64_t B, float eps0, float eps1,
                                                        int min_time, float* __restrict__ norm) {
  float s1 = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) s1 += __ldg(dones + i);
  __shared__ float sh[32];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s1;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)sh[w];
    norm[0] = (float)((double)B - t + (double)eps0);
    norm[1] = min_time ? 1.0f : (float)(t + (double)eps1);
  }
}

// grad[j] = sum over CTAs (same fixed order as vhjb_reduce_kernel) followed by the optax.adam update of element j; the
// thread that owns j = 0 also reduces the loss sums and the saturation count.  One launch instead of four.
__global__ void __launch_bounds__(256) vhjb_reduce_adam_kernel(const float* __restrict__ partial, int64_t pstride, int ncta, int P,
                                                               float* __restrict__ grad, float* __restrict__ sums,
                                                               float* __restrict__ sat, float* __restrict__ w,
                                                               float* __restrict__ m, float* __restrict__ v, float lr, float b1,
                                                               float b2, float eps, float bc1, float bc2,
                                                               const float* __restrict__ norm, float reg,
                                                               float* __restrict__ loss_acc, const float* __restrict__ dtail) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (dtail != nullptr && dtail[5] != 0.f) ncta = 0;     // the fp32 pass redid the whole batch (vhjb_simt.cuh: redo_all)
  if (j < P) {
    float g = 0.f;
    for (int c = 0; c < ncta; ++c) g += partial[(int64_t)c * pstride + j];
    if (dtail != nullptr) {   // the fp32 pass over the deferred states (see vhjb_reduce_kernel)
      const int nd = (int)dtail[3];
      for (int c = 0; c < nd; ++c) g += partial[(int64_t)(kMaxCtas + c) * pstride + j];
    }
    grad[j] = g;
    const float mi = fmaf(b1, m[j], (1.f - b1) * g);
    const float vi = fmaf(b2, v[j], (1.f - b2) * g * g);
    m[j] = mi;
    v[j] = vi;
    const float mhat = mi / bc1, vhat = vi / bc2;
    w[j] = w[j] - lr * mhat / (sqrtf(vhat) + eps);
  }
  if (j < 3) {
    float t = 0.f;
    for (int c = 0; c < ncta; ++c) t += partial[(int64_t)c * pstride + P + j];
    if (dtail != nullptr && j < 2) {   // loss terms of the deferred states
      const int nd = (int)dtail[3];
      for (int c = 0; c < nd; ++c) t += partial[(int64_t)(kMaxCtas + c) * pstride + P + j];
    }
    if (j < 2) sums[j] = t;
    else { sat[0] = t; sat[1] += t; }   // [0]: this launch, [1]: running total (hjb_vhjb_saturation_total)
  }
  if (j == 0 && loss_acc != nullptr) {
    // running sums of the step losses (total, hjb, term) of params_update, controller/vhjb.py:282-288: a training loop
    // reads them once per epoch instead of doing tensor arithmetic on two scalars after every update
    float s0 = 0.f, s1 = 0.f;
    for (int c = 0; c < ncta; ++c) { s0 += partial[(int64_t)c * pstride + P]; s1 += partial[(int64_t)c * pstride + P + 1]; }
    if (dtail != nullptr) {
      const int nd = (int)dtail[3];
      for (int c = 0; c < nd; ++c) {
        s0 += partial[(int64_t)(kMaxCtas + c) * pstride + P];
        s1 += partial[(int64_t)(kMaxCtas + c) * pstride + P + 1];
      }
    }
    const float hjb = s0 / norm[0], term = s1 / norm[1];
    loss_acc[0] += hjb + reg * term;
    loss_acc[1] += hjb;
    loss_acc[2] += term;
  }
}


// ---- several GPUs: reduce -> exchange over NVLink peer memory -> Adam, in ONE kernel -------------------------------------
// The multi-GPU train step used to end with three reduction launches, an NCCL all-reduce of [grad | loss sums] (~101 KB:
// pure latency on NVLink / NVSwitch) and an Adam launch.  Here CTA b owns elements [256 b, 256 b + 256) of the exchanged
// vector q = [grad (P) | hjb sum, term sum | this rank's done-counts of the NEXT batch (2)]:
//   1. it sums ITS elements over the per-CTA partial slots (same fixed order as vhjb_reduce_kernel),
//   2. stores them into slot `rank` of EVERY rank's exchange buffer (plain stores to peer memory: NVLink writes),
//   3. publishes: __threadfence_system, then st.release.sys of the step number to flag (rank, b) on every rank,
//   4. waits (bounded) until the flags (r, b) of all ranks r on ITS OWN GPU carry the step number,
//   5. sums the W slots in rank order — the same order on every rank, so every rank holds the same bits — and
//   6. applies the optax.adam update to its elements of the weights.
// No CTA waits for another CTA of its own GPU (no co-residency requirement), only for the same CTA of the peers.  Buffers
// are double-buffered by the parity of the step: a rank can only write step k + 2 after it saw every peer's step k + 1,
// which a peer publishes after its own step-k kernel completed.  A wait that gives up raises the workspace's failure word,
// poisons the loss sums with NaN and skips the update (never a hung device, never a silent garbage step).
struct PeerExchange {
  float* const* bufs;        // device array [world]: rank r's exchange buffer, [2][world][qpad] floats
  uint32_t* const* flags;    // device array [world]: rank r's flags, [world][nblk]
  int rank, world;
  uint32_t seq;              // step number (> 0, increasing)
  int qpad;
  const float* next_counts;  // this rank's [sum(1 - done), sum(done)] of the next batch (or zeros)
  float* counts_out;         // the next batch's GLOBAL normalisers: counts + eps ([1] = 1 for the min-time form)
  float norm_eps;
  int min_time;
  unsigned poll_limit;
};

__global__ void __launch_bounds__(256) vhjb_reduce_exchange_adam_kernel(const float* __restrict__ partial, int64_t pstride, int ncta,
                                                                        int P, const float* __restrict__ dtail, PeerExchange x,
                                                                        float* __restrict__ grad, float* __restrict__ sums,
                                                                        float* __restrict__ tail, float* __restrict__ w,
                                                                        float* __restrict__ m, float* __restrict__ v, float lr,
                                                                        float b1, float b2, float eps, float bc1, float bc2,
                                                                        const float* __restrict__ norm, float reg,
                                                                        float* __restrict__ loss_acc) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int Q = P + 4, nblk = gridDim.x;
  const int parity = (int)(x.seq & 1u);
  if (dtail != nullptr && dtail[5] != 0.f) ncta = 0;     // the fp32 pass redid the whole batch (vhjb_simt.cuh: redo_all)
  float mine = 0.f;
  if (j < P + 2) {
    for (int c = 0; c < ncta; ++c) mine += partial[(int64_t)c * pstride + j];
    if (dtail != nullptr) {
      const int nd = (int)dtail[3];
      for (int c = 0; c < nd; ++c) mine += partial[(int64_t)(kMaxCtas + c) * pstride + j];
    }
  } else if (j < Q) {
    mine = x.next_counts ? x.next_counts[j - (P + 2)] : 0.f;
  }
  if (j < Q) {
    const size_t off = ((size_t)parity * x.world + x.rank) * x.qpad + j;
    for (int r = 0; r < x.world; ++r) x.bufs[r][off] = mine;
  }
  if (blockIdx.x == 0 && threadIdx.x == 32) {   // saturation count of this rank's launch (local, like vhjb_sat_kernel)
    float t = 0.f;
    for (int c = 0; c < ncta; ++c) t += partial[(int64_t)c * pstride + P + 2];
    tail[0] = t;
    tail[1] += t;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int failed;
  if (threadIdx.x == 0) failed = 0;
  if (threadIdx.x < x.world) {
    uint32_t* f = x.flags[threadIdx.x] + (size_t)x.rank * nblk + blockIdx.x;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(x.seq) : "memory");
  }
  __syncthreads();
  if (threadIdx.x < x.world) {
    const uint32_t* f = x.flags[x.rank] + (size_t)threadIdx.x * nblk + blockIdx.x;
    uint32_t got = 0;
    unsigned spins = 0;
    for (; spins < x.poll_limit; ++spins) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(f) : "memory");
      if (got == x.seq) break;
      __nanosleep(200);
    }
    if (got != x.seq) atomicAdd(&failed, 1);
  }
  __syncthreads();
  const bool bad = failed != 0;
  if (bad && threadIdx.x == 0) atomicAdd(tail + 2, 1.0f);
  float g = 0.f;
  if (j < Q) {
    const float* mybuf = x.bufs[x.rank] + (size_t)parity * x.world * x.qpad;
    for (int r = 0; r < x.world; ++r) g += __ldcg(mybuf + (size_t)r * x.qpad + j);
  }
  const float nanv = __int_as_float(0x7fc00000);
  if (j < P) {
    grad[j] = g;
    if (!bad) {
      const float mi = fmaf(b1, m[j], (1.f - b1) * g);
      const float vi = fmaf(b2, v[j], (1.f - b2) * g * g);
      m[j] = mi;
      v[j] = vi;
      const float mhat = mi / bc1, vhat = vi / bc2;
      w[j] = w[j] - lr * mhat / (sqrtf(vhat) + eps);
    }
  } else if (j < P + 2) {
    sums[j - P] = bad ? nanv : g;
  } else if (j < Q) {
    if (x.counts_out) x.counts_out[j - (P + 2)] = (x.min_time && j == P + 3) ? 1.0f : g + x.norm_eps;
  }
  if (loss_acc != nullptr && j == P) {          // the thread that owns the hjb sum also reads the term sum's slots
    const float* mybuf = x.bufs[x.rank] + (size_t)parity * x.world * x.qpad;
    float s1 = 0.f;
    for (int r = 0; r < x.world; ++r) s1 += __ldcg(mybuf + (size_t)r * x.qpad + P + 1);
    const float hjb = g / norm[0], term = s1 / norm[1];
    loss_acc[0] += hjb + reg * term;
    loss_acc[1] += hjb;
    loss_acc[2] += term;
  }
}

// ---- optax.adam ----
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                                                   const float* __restrict__ g, int64_t len, float lr, float b1, float b2,
                                                   float eps, float bc1, float bc2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const float gi = g[i];
  const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
  const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  const float mhat = mi / bc1, vhat = vi / bc2;
  w[i] = w[i] - lr * mhat / (sqrtf(vhat) + eps);
}

static int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
      cached = sms;
  }
  return cached > 0 ? (cached < kMaxCtas ? cached : kMaxCtas) : 148;
}

struct AdamTail {   // when set, run_vhjb ends with vhjb_reduce_adam_kernel instead of the three reductions
  float *w, *m, *v;
  float lr, b1, b2, eps, bc1, bc2;
  float* loss_acc;
  const PeerExchange* peer = nullptr;   // several GPUs: reduce + exchange over peer memory + Adam in one kernel
};

static int run_vhjb(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs, const float* dones,
                    const float* costs, int64_t B, const float* norm, float reg, float* V, float* p, float* u, float* r,
                    float* grad, float* sums, void* workspace, bool want_grad, bool accumulate, cudaStream_t st,
                    const AdamTail* tail = nullptr, const int32_t* ready = nullptr, int64_t piece_states = 0) {
  if (!sys || !net || !task || B < 0) return HJB_ERR_BAD_ARG;
  if (net->n != sys->n || net->features[0] != VH1 || net->features[1] != VH2 || net->features[2] != VH3)
    return HJB_ERR_UNSUPPORTED;
  if (!net->params || !workspace) return HJB_ERR_BAD_ARG;
  if (B > 0 && (!xs || !dones || !costs)) return HJB_ERR_BAD_ARG;
  if (want_grad && (!grad || !norm)) return HJB_ERR_BAD_ARG;
  const int n = sys->n, m = sys->m;

  VhjbArgs a;
  std::memset(&a, 0, sizeof(a));
  make_dev_sys(sys, a.sys);  // aoff = 0: same folding as the rollout path
  a.params = net->params;
  for (int i = 0; i < n; ++i) {
    a.mean[i] = net->mean[i];
    a.inv_std[i] = (float)(1.0 / (double)net->std[i]);
    a.xf[i] = net->xf[i];
  }
  a.eps_s = net->eps_s;
  for (int i = 0; i < n * n; ++i) a.Q[i] = task->Q[i];
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      a.R[i * m + j] = task->R[i * m + j];
      a.Rsym[i * m + j] = task->R[i * m + j] + task->R[j * m + i];
      a.Rinv[i * m + j] = task->Rinv[i * m + j];
    }
  for (int i = 0; i < m; ++i) a.uf[i] = task->uf[i];
  a.eps = task->eps;
  a.xs = xs; a.dones = dones; a.costs = costs; a.B = B;
  a.norm = norm; a.reg = reg;
  a.V = V; a.p = p; a.u = u; a.r = r;
  a.partial = static_cast<float*>(workspace);
  a.pstride = pstride_of(n);
  a.tail = a.partial + (int64_t)kSlots * a.pstride;
  a.defer_count = reinterpret_cast<int*>(a.tail + 8);
  a.defer_index = a.defer_count + kDeferLists;
  const bool tensor = use_tensor_path(net, B);
  const int tile = tensor ? tc::TS : VBM;
  a.n_tiles = (B + tile - 1) / tile;
  if (ready) {   // streamed batch: tensor-core gradient kernel only, pieces of whole tiles
    if (!tensor || !want_grad || piece_states <= 0 || piece_states % tc::TS != 0) return HJB_ERR_UNSUPPORTED;
    a.ready = ready;
    a.piece_tiles = piece_states / tc::TS;
    a.poll_limit = 1u << 22;
    if (const char* pl = std::getenv("HJB_STREAM_POLL_LIMIT")) {
      const long v = std::atol(pl);
      if (v > 0) a.poll_limit = (unsigned)v;
    }
  }

  static long long* dbg_buf = nullptr;
  const bool dbg = tensor && std::getenv("HJB_TC_DEBUG_TIMING") != nullptr;
  if (dbg) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 64 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 64 * sizeof(long long), st);
    a.dbg = dbg_buf;
  }

  VhjbLaunch l;
  l.grad = want_grad;
  // the residual-only tensor kernel runs two tiles per CTA (vhjb_tc.cuh)
  const int64_t cta_tiles = (tensor && !want_grad) ? (a.n_tiles + 1) / 2 : a.n_tiles;
  int64_t grid = cta_tiles < sm_count() ? cta_tiles : sm_count();
  if (grid < 1) grid = 1;
  l.grid = (int)grid;
  a.defer_lists = 4 * l.grid;

  cudaError_t e = cudaErrorNotSupported;
  if (tensor) {
    switch (sys->kind) {
      case HJB_SYS_LINEAR:
        if (n == 2 && m == 1) e = vhjb_tc_launch_linear21(a, l, net->act, task->control_form, task->residual_form, st);
        break;
      case HJB_SYS_CARTPOLE: e = vhjb_tc_launch_cartpole(a, l, net->act, task->control_form, task->residual_form, st); break;
      case HJB_SYS_QUAD2D: e = vhjb_tc_launch_quad2d(a, l, net->act, task->control_form, task->residual_form, st); break;
      case HJB_SYS_QUAD10D: e = vhjb_tc_launch_quad10d(a, l, net->act, task->control_form, task->residual_form, st); break;
      default: break;
    }
  } else
  switch (sys->kind) {
    case HJB_SYS_LINEAR:
      if (n == 2 && m == 1) e = vhjb_launch_linear21(a, l, net->act, task->control_form, task->residual_form, st);
      break;
    case HJB_SYS_CARTPOLE: e = vhjb_launch_cartpole(a, l, net->act, task->control_form, task->residual_form, st); break;
    case HJB_SYS_QUAD2D: e = vhjb_launch_quad2d(a, l, net->act, task->control_form, task->residual_form, st); break;
    case HJB_SYS_QUAD10D: e = vhjb_launch_quad10d(a, l, net->act, task->control_form, task->residual_form, st); break;
    default: break;
  }
  if (e == cudaErrorNotSupported) return HJB_ERR_UNSUPPORTED;
  if (e != cudaSuccess) return (int)e;
  // The tensor-core gradient kernel leaves the states whose adjoint seeds lie beyond its fp16 range management to this
  // launch: the CUDA-core kernel over exactly those states (fp32), partials into the second half of the slots.  CTAs
  // without work return before touching their weights: an empty pass costs one launch.
  const bool deferred = tensor && want_grad;
  if (deferred) {
    VhjbArgs d = a;
    d.defer_gather = 1;
    d.part_slot0 = kMaxCtas;
    d.ready = nullptr;
    d.dbg = nullptr;
    VhjbLaunch ld;
    ld.grad = true;
    const int64_t dt = (B + VBM - 1) / VBM;
    ld.grid = (int)(dt < sm_count() ? dt : sm_count());
    cudaError_t de = cudaErrorNotSupported;
    switch (sys->kind) {
      case HJB_SYS_LINEAR:
        if (n == 2 && m == 1) de = vhjb_launch_linear21(d, ld, net->act, task->control_form, task->residual_form, st);
        break;
      case HJB_SYS_CARTPOLE: de = vhjb_launch_cartpole(d, ld, net->act, task->control_form, task->residual_form, st); break;
      case HJB_SYS_QUAD2D: de = vhjb_launch_quad2d(d, ld, net->act, task->control_form, task->residual_form, st); break;
      case HJB_SYS_QUAD10D: de = vhjb_launch_quad10d(d, ld, net->act, task->control_form, task->residual_form, st); break;
      default: break;
    }
    if (de == cudaErrorNotSupported) return HJB_ERR_UNSUPPORTED;
    if (de != cudaSuccess) return (int)de;
  }
  const float* dtail = deferred ? a.tail : nullptr;
  if (dbg) {  // developer probe: clock64 at (after wait_mma, end of pass) of every step of CTA 0's third tile
    long long h[64];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    std::fprintf(stderr, "[hjb tc timing] grad=%d:", (int)want_grad);
    if (h[63] > h[61])
      std::fprintf(stderr, " [CTA 0: %lld cycles in %lld ns = %.0f MHz]", h[62] - h[60], h[63] - h[61],
                   1e3 * (double)(h[62] - h[60]) / (double)(h[63] - h[61]));
    for (int i = 1; i < 58 && h[i] != 0; ++i) std::fprintf(stderr, " %s%lld", (i & 1) ? "w" : "p", h[i] - h[i - 1]);
    std::fprintf(stderr, "\n");
  }
  const int P = vhjb_param_count(n);
  if (tail && tail->peer) {
    vhjb_reduce_exchange_adam_kernel<<<(P + 4 + 255) / 256, 256, 0, st>>>(a.partial, a.pstride, l.grid, P, dtail, *tail->peer, grad, sums,
                                                                          a.tail, tail->w, tail->m, tail->v, tail->lr, tail->b1,
                                                                          tail->b2, tail->eps, tail->bc1, tail->bc2, norm, reg,
                                                                          tail->loss_acc);
    e = cudaGetLastError();
    return e == cudaSuccess ? HJB_OK : (int)e;
  }
  if (tail) {
    vhjb_reduce_adam_kernel<<<(P + 255) / 256, 256, 0, st>>>(a.partial, a.pstride, l.grid, P, grad, sums, a.tail, tail->w, tail->m,
                                                             tail->v, tail->lr, tail->b1, tail->b2, tail->eps, tail->bc1, tail->bc2,
                                                             norm, reg, tail->loss_acc, dtail);
    e = cudaGetLastError();
    return e == cudaSuccess ? HJB_OK : (int)e;
  }
  if (want_grad) {
    vhjb_reduce_kernel<<<(P + 255) / 256, 256, 0, st>>>(a.partial, a.pstride, l.grid, 0, P, grad, (int)accumulate, dtail);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  if (sums) {
    vhjb_reduce_kernel<<<1, 256, 0, st>>>(a.partial, a.pstride, l.grid, P, 2, sums, (int)accumulate, dtail);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  if (want_grad) {  // saturation count of the fp16 range management (vhjb_tc.cuh) -> workspace tail
    vhjb_sat_kernel<<<1, 32, 0, st>>>(a.partial, a.pstride, l.grid, P + 2, a.tail, (int)accumulate, ready != nullptr, sums, dtail);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return HJB_OK;
}

}  // namespace hjb

using namespace hjb;

extern "C" {

int64_t hjb_vhjb_param_count(int32_t n) { return n > 0 && n <= HJB_MAX_N ? vhjb_param_count(n) : -1; }

int64_t hjb_vhjb_workspace_bytes(int32_t n) {
  if (n <= 0 || n > HJB_MAX_N) return -1;
  return ((int64_t)kSlots * pstride_of(n) + 8) * (int64_t)sizeof(float) +
         ((int64_t)kDeferLists + (int64_t)kDeferLists * kDeferCap) * (int64_t)sizeof(int);
}

int hjb_vhjb_saturation(const void* workspace, int32_t n, float* count, void* stream) {
  if (!workspace || !count || n <= 0 || n > HJB_MAX_N) return HJB_ERR_BAD_ARG;
  const float* tail = static_cast<const float*>(workspace) + (int64_t)kSlots * pstride_of(n);
  cudaError_t e = cudaMemcpyAsync(count, tail, sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  return e == cudaSuccess ? HJB_OK : (int)e;
}

int hjb_vhjb_saturation_total(void* workspace, int32_t n, float* count, int32_t reset, void* stream) {
  if (!workspace || n <= 0 || n > HJB_MAX_N) return HJB_ERR_BAD_ARG;
  float* tail = static_cast<float*>(workspace) + (int64_t)kSlots * pstride_of(n);
  cudaError_t e = cudaSuccess;
  if (count) e = cudaMemcpyAsync(count, tail + 1, sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e == cudaSuccess && reset) e = cudaMemsetAsync(tail + 1, 0, sizeof(float), (cudaStream_t)stream);
  return e == cudaSuccess ? HJB_OK : (int)e;
}

int hjb_vhjb_deferred(const void* workspace, int32_t n, float* count, void* stream) {
  if (!workspace || !count || n <= 0 || n > HJB_MAX_N) return HJB_ERR_BAD_ARG;
  const float* tail = static_cast<const float*>(workspace) + (int64_t)kSlots * pstride_of(n);
  cudaError_t e = cudaMemcpyAsync(count, tail + 4, sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream);
  return e == cudaSuccess ? HJB_OK : (int)e;
}

int hjb_vhjb_stream_failures(void* workspace, int32_t n, float* count, int32_t reset, void* stream) {
  if (!workspace || n <= 0 || n > HJB_MAX_N) return HJB_ERR_BAD_ARG;
  float* tail = static_cast<float*>(workspace) + (int64_t)kSlots * pstride_of(n);
  cudaError_t e = cudaSuccess;
  if (count) e = cudaMemcpyAsync(count, tail + 2, sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream);
  if (e == cudaSuccess && reset) e = cudaMemsetAsync(tail + 2, 0, sizeof(float), (cudaStream_t)stream);
  return e == cudaSuccess ? HJB_OK : (int)e;
}

int hjb_vhjb_adam_guarded(float* params, float* m, float* v, const float* grad, int32_t n, float lr, float b1, float b2, float eps,
                          int32_t step, const void* workspace, void* stream) {
  if (!params || !m || !v || !grad || !workspace || n <= 0 || n > HJB_MAX_N || step < 1) return HJB_ERR_BAD_ARG;
  const int64_t len = vhjb_param_count(n);
  const float bc1 = (float)(1.0 - std::pow((double)b1, (double)step));
  const float bc2 = (float)(1.0 - std::pow((double)b2, (double)step));
  const float* guard = static_cast<const float*>(workspace) + (int64_t)kSlots * pstride_of(n) + 2;
  adam_guarded_kernel<<<(unsigned)((len + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, m, v, grad, len, lr, b1, b2, eps, bc1,
                                                                                       bc2, guard);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HJB_OK : (int)e;
}

int hjb_vhjb_count(const float* dones, int64_t B, float eps, float* norm, void* workspace, void* stream) {
  if (!norm || !workspace || B < 0 || (B > 0 && !dones)) return HJB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = sm_count();
  float* part = static_cast<float*>(workspace);
  count_stage1<<<nblk, 256, 0, st>>>(dones, B, part);
  count_stage2<<<1, 32, 0, st>>>(part, nblk, B, eps, 0, norm);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HJB_OK : (int)e;
}

int hjb_vhjb_residual(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs, const float* dones,
                      const float* costs, int64_t B, float* V, float* p, float* u, float* r, float* sums, void* workspace,
                      void* stream) {
  return run_vhjb(sys, net, task, xs, dones, costs, B, nullptr, 0.f, V, p, u, r, nullptr, sums, workspace, false, false,
                  (cudaStream_t)stream);
}

int hjb_vhjb_loss_grad(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs, const float* dones,
                       const float* costs, int64_t B, const float* norm, float reg, float* grad, float* sums, void* workspace,
                       void* stream) {
  return run_vhjb(sys, net, task, xs, dones, costs, B, norm, reg, nullptr, nullptr, nullptr, nullptr, grad, sums, workspace,
                  true, false, (cudaStream_t)stream);
}

int hjb_vhjb_loss_grad_accumulate(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs,
                                  const float* dones, const float* costs, int64_t B, const float* norm, float reg,
                                  float* grad, float* sums, void* workspace, void* stream) {
  return run_vhjb(sys, net, task, xs, dones, costs, B, norm, reg, nullptr, nullptr, nullptr, nullptr, grad, sums, workspace,
                  true, true, (cudaStream_t)stream);
}

int hjb_vhjb_stream_batch(const float* xs_host, const float* costs_host, float* xs, float* costs, int64_t B, int32_t n,
                          int64_t piece_states, int32_t* ready, const int32_t* ones_host, void* copy_stream) {
  if (!xs_host || !costs_host || !xs || !costs || !ready || !ones_host || B < 0 || n <= 0 || n > HJB_MAX_N || piece_states <= 0)
    return HJB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)copy_stream;
  int k = 0;
  for (int64_t lo = 0; lo < B; lo += piece_states, ++k) {
    const int64_t cnt = (B - lo < piece_states) ? B - lo : piece_states;
    cudaError_t e = cudaMemcpyAsync(xs + lo * n, xs_host + lo * n, (size_t)cnt * n * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(costs + lo, costs_host + lo, (size_t)cnt * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ready + k, ones_host + k, sizeof(int32_t), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
  }
  return HJB_OK;
}

int hjb_vhjb_loss_grad_streamed(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs,
                                const float* dones, const float* costs, int64_t B, const float* norm, float reg, float* grad,
                                float* sums, void* workspace, const int32_t* ready, int64_t piece_states, void* stream) {
  if (!ready) return HJB_ERR_BAD_ARG;
  return run_vhjb(sys, net, task, xs, dones, costs, B, norm, reg, nullptr, nullptr, nullptr, nullptr, grad, sums, workspace,
                  true, false, (cudaStream_t)stream, nullptr, ready, piece_states);
}

int hjb_vhjb_train_step(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs, const float* dones,
                        const float* costs, int64_t B, float reg, float lr, float b1, float b2, float adam_eps, int32_t step,
                        float* m, float* v, float* norm, float* grad, float* sums, float* loss_acc, void* workspace,
                        void* stream) {
  if (!sys || !net || !task || !net->params || !m || !v || !norm || !grad || !sums || !workspace || B <= 0 || step < 1)
    return HJB_ERR_BAD_ARG;
  if (!xs || !dones || !costs) return HJB_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int min_time = task->residual_form == HJB_RES_MIN_TIME;
  if (B <= (int64_t)1 << 18) {
    count_one_block<<<1, 1024, 0, st>>>(dones, B, min_time ? 0.f : task->eps, task->eps, min_time, norm);
  } else {
    const int nblk = sm_count();
    float* part = static_cast<float*>(workspace);
    count_stage1<<<nblk, 256, 0, st>>>(dones, B, part);
    count_stage2<<<1, 32, 0, st>>>(part, nblk, B, min_time ? 0.f : task->eps, min_time, norm);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  AdamTail tail;
  tail.w = const_cast<float*>(net->params);
  tail.m = m;
  tail.v = v;
  tail.lr = lr; tail.b1 = b1; tail.b2 = b2; tail.eps = adam_eps;
  tail.bc1 = (float)(1.0 - std::pow((double)b1, (double)step));
  tail.bc2 = (float)(1.0 - std::pow((double)b2, (double)step));
  tail.loss_acc = loss_acc;
  return run_vhjb(sys, net, task, xs, dones, costs, B, norm, reg, nullptr, nullptr, nullptr, nullptr, grad, sums, workspace, true,
                  false, st, &tail);
}

int64_t hjb_vhjb_peer_exchange_floats(int32_t n, int32_t world) {
  if (n <= 0 || n > HJB_MAX_N || world < 1) return -1;
  const int64_t qpad = (vhjb_param_count(n) + 4 + 63) / 64 * 64;
  return 2 * (int64_t)world * qpad;
}
int64_t hjb_vhjb_peer_exchange_flags(int32_t n, int32_t world) {
  if (n <= 0 || n > HJB_MAX_N || world < 1) return -1;
  return (int64_t)world * ((vhjb_param_count(n) + 4 + 255) / 256);
}

int hjb_vhjb_train_step_peer(const hjb_system* sys, const hjb_vnet* net, const hjb_task* task, const float* xs, const float* dones,
                             const float* costs, int64_t B, float reg, float lr, float b1, float b2, float adam_eps, int32_t step,
                             float* m, float* v, const float* norm, float* grad, float* sums, float* loss_acc,
                             const float* next_counts, float* counts_out, void* const* peer_bufs, void* const* peer_flags,
                             int32_t rank, int32_t world, void* workspace, void* stream) {
  if (!sys || !net || !task || !net->params || !m || !v || !norm || !grad || !sums || !workspace || B < 0 || step < 1)
    return HJB_ERR_BAD_ARG;
  if (!peer_bufs || !peer_flags || world < 1 || world > 64 || rank < 0 || rank >= world) return HJB_ERR_BAD_ARG;
  if (B > 0 && (!xs || !dones || !costs)) return HJB_ERR_BAD_ARG;
  PeerExchange x;
  x.bufs = reinterpret_cast<float* const*>(peer_bufs);
  x.flags = reinterpret_cast<uint32_t* const*>(peer_flags);
  x.rank = rank; x.world = world;
  x.seq = (uint32_t)step;
  x.qpad = (int)((vhjb_param_count(sys->n) + 4 + 63) / 64 * 64);
  x.next_counts = next_counts;
  x.counts_out = counts_out;
  x.min_time = task->residual_form == HJB_RES_MIN_TIME;
  x.norm_eps = x.min_time ? 0.f : task->eps;
  x.poll_limit = 1u << 23;   // x 200 ns: ~2 s
  if (const char* pl = std::getenv("HJB_PEER_POLL_LIMIT")) {
    const long pv = std::atol(pl);
    if (pv > 0) x.poll_limit = (unsigned)pv;
  }
  AdamTail tail;
  tail.w = const_cast<float*>(net->params);
  tail.m = m;
  tail.v = v;
  tail.lr = lr; tail.b1 = b1; tail.b2 = b2; tail.eps = adam_eps;
  tail.bc1 = (float)(1.0 - std::pow((double)b1, (double)step));
  tail.bc2 = (float)(1.0 - std::pow((double)b2, (double)step));
  tail.loss_acc = loss_acc;
  tail.peer = &x;
  return run_vhjb(sys, net, task, xs, dones, costs, B, norm, reg, nullptr, nullptr, nullptr, nullptr, grad, sums, workspace, true,
                  false, (cudaStream_t)stream, &tail);
}

int hjb_adam(float* params, float* m, float* v, const float* grad, int64_t len, float lr, float b1, float b2, float eps,
             int32_t step, void* stream) {
  if (!params || !m || !v || !grad || len < 0 || step < 1) return HJB_ERR_BAD_ARG;
  if (len == 0) return HJB_OK;
  const float bc1 = (float)(1.0 - std::pow((double)b1, (double)step));
  const float bc2 = (float)(1.0 - std::pow((double)b2, (double)step));
  adam_kernel<<<(unsigned)((len + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, m, v, grad, len, lr, b1, b2, eps, bc1, bc2);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HJB_OK : (int)e;
}

}  // extern "C"

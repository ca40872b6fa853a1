// K2 on the tensor cores, second generation (relu value nets): the residual-only pass (rows V1-V5 of SURVEY.md 8a) with
// the STATES ON THE TMEM LANES and the activations fed to the MMAs FROM TENSOR MEMORY (".ts" form).
//
// Round 1's residual kernel (vhjb_tc.cuh, still used for tanh / sin nets) puts the features on the lanes and 64 states on
// the MMA's N dimension; its A operand is a resident weight matrix in shared memory, and an SS-mode MMA re-reads that
// 4 KB A tile for every instruction: 48.6 cycles at N = 64 against a math floor of 32 (measured, tests/cuda/umma_rate.cu)
// — a 67 % ceiling by construction.  Here a tile is 128 states, one per TMEM lane:
//
//     D[s][j] (+)= sum_k A[s][k] B[k][j]     A = activations of the tile, in TENSOR MEMORY (packed fp16 pairs, written by
//                                            the thread that owns lane s with tcgen05.st — they never touch shared memory),
//                                            B = a resident weight matrix in shared memory, N = 128 / 64 / 16
//
// and every MMA runs at its math floor (tests/cuda/ts_probe.cu: 64.5 / 32.5 / 9.6 cycles for N = 128 / 64 / 16 with A in
// TMEM, against 64.6 / 48.6 / 39.5 from shared memory): 39.7 tensor-pipe cycles per state instead of 68.9.  The six GEMMs
// of a tile, each as hi hi + lo hi + hi lo over fp16 pieces (same products, same precision as round 1):
//
//     G0  a1 = h0 W1        K = 16   N = 128      P1  h1 = relu(a1), mask1      -> A
//     G1  a2 = h1 W2        K = 128  N = 128      P2  h2 = relu(a2), mask2      -> A
//     G2  y  = h2 W3        K = 128  N = 64       P3  V = |y|^2, gy = 2 y       -> A
//     G3  b2 = gy W3^T      K = 64   N = 128      P4  g2 = b2 . mask2           -> A
//     G4  b1 = g2 W2^T      K = 128  N = 128      P5  g1 = b1 . mask1           -> A
//     G5  g0 = g1 W1^T      K = 128  N = 16       P6  per-state epilogue (control, Hamiltonian residual, outputs)
//
// The relu masks are 64 bits per thread and layer, kept in registers (thread = one state x 64 features), so the
// pre-activations need not survive in TMEM: a group of 8 warps needs 128 accumulator columns + 128 operand columns, and two
// groups (two tiles in flight, ping-pong: one group's GEMM runs under the other group's element-wise pass) fill the 512.
// Shared memory holds nothing but the split weights (104 KB).  Smooth activations need sigma'(a) again in P4 / P5
// (256 more columns per group): they stay on the round-1 kernel.
#pragma once
#include "vhjb_tc.cuh"

namespace hjb {
namespace tc {

constexpr int TSM = 128;                                   // states per tile = TMEM lanes
constexpr int kRes2Groups = 2;
constexpr int kRes2Threads = 32 * kComputeWarps * kRes2Groups;
constexpr uint32_t r2D = 0, r2Ahi = 128, r2Alo = 192, kRes2Cols = 256;   // TMEM columns of a group
constexpr uint32_t kRes2Misc = kF0;                        // first byte after the resident weights
constexpr uint32_t kRes2SmemBytes = kRes2Misc + 2048;      // float sV[2][128]; u64 bars[2]; u32 tmem; float sLoss[16]

// D (+)= A B^T with A in tensor memory: three passes over the fp16 pieces — (hi, hi), (lo, hi), (hi, lo)
template <class B, uint32_t IDESC, int KSTEPS, int I>
__device__ __forceinline__ void mma_ts_step(uint32_t tmg, uint32_t sbd) {
  constexpr int pr = I / KSTEPS, k = I % KSTEPS;
  constexpr uint32_t acol = (pr == 1 ? r2Alo : r2Ahi) + 8u * k;
  constexpr uint32_t b_lo = ((B::addr + (pr == 2 ? B::piece : 0u) + k * B::kadv) >> 4) | ((B::lbo >> 4) << 16);
  constexpr uint32_t b_hi = (B::sbo >> 4) | (1u << 14);
  mma_ts_imm<acol, b_lo, b_hi, IDESC, r2D>(tmg, sbd, I == 0 ? 0u : 1u);
}
template <class B, uint32_t IDESC, int KSTEPS, int... I>
__device__ __forceinline__ void gemm3_ts_seq(uint32_t tmg, uint32_t sbd, std::integer_sequence<int, I...>) {
  (mma_ts_step<B, IDESC, KSTEPS, I>(tmg, sbd), ...);
}
template <int KSTEPS, class B, uint32_t IDESC>
__device__ __forceinline__ void gemm3_ts(uint32_t tmg, uint32_t sbd) {
  gemm3_ts_seq<B, IDESC, KSTEPS>(tmg, sbd, std::make_integer_sequence<int, 3 * KSTEPS>{});
}

template <int FMT>
struct Res2Ops {
  using W2_mn = MnMaj<kW2, kW2Piece, kRbW2>; using W2_k = KMaj<kW2, kW2Piece, kRbW2>;
  using W3_mn = MnMaj<kW3, kW3Piece, kRbW3>; using W3_k = KMaj<kW3, kW3Piece, kRbW3>;
  using W1_mn = MnMaj<kW1, kW1Piece, kRbW1>; using W1_k = KMaj<kW1, kW1Piece, kRbW1>;
  static constexpr uint32_t idN128_mn = idesc_f16(128, 128, FMT, FMT, 0, 1), idN128_k = idesc_f16(128, 128, FMT, FMT, 0, 0),
                            idN64_mn = idesc_f16(128, VH3, FMT, FMT, 0, 1), idN16_k = idesc_f16(128, 16, FMT, FMT, 0, 0);
  // the six GEMMs of a tile, in order (tmg = TMEM base of the group)
  static __device__ __forceinline__ void issue(int step, uint32_t tmg, uint32_t sb) {
    switch (step) {
      case 0: gemm3_ts<1, W1_mn, idN128_mn>(tmg, sb); break;   // a1 = h0 W1
      case 1: gemm3_ts<8, W2_mn, idN128_mn>(tmg, sb); break;   // a2 = h1 W2
      case 2: gemm3_ts<8, W3_mn, idN64_mn>(tmg, sb); break;    // y  = h2 W3
      case 3: gemm3_ts<4, W3_k, idN128_k>(tmg, sb); break;     // b2 = gy W3^T
      case 4: gemm3_ts<8, W2_k, idN128_k>(tmg, sb); break;     // b1 = g2 W2^T
      default: gemm3_ts<8, W1_k, idN16_k>(tmg, sb); break;     // g0 = g1 W1^T
    }
  }
};

template <class S, int UFORM, int RFORM, int FMT>
__global__ void __launch_bounds__(kRes2Threads, 1) vhjb_tc_residual2_kernel(const __grid_constant__ VhjbArgs a) {
  constexpr int N = S::N;
  static_assert(N <= 16, "state dimension padded to one K = 16 step");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRes2Misc + 1024);   // [g]: this group's GEMM done
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + kRes2Misc + 1024 + 32);
  float* sLoss = reinterpret_cast<float*>(smem + kRes2Misc + 1024 + 64);   // [g][q][2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  {  // weights -> shared memory (split, core-matrix layout), once per CTA
    const float* W1 = a.params;
    const float* W2 = W1 + N * VH1;
    const float* W3 = W2 + VH1 * VH2;
    for (int c = tid; c < VH1 * (VH2 / 8); c += kRes2Threads) {
      const int k = c / (VH2 / 8), jb = c % (VH2 / 8);
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(W2 + k * VH2 + 8 * jb));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(W2 + k * VH2 + 8 * jb) + 1);
      const float o[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      store8<FMT>(smem, kW2, kW2Piece, kRbW2, k, 8 * jb, o);
    }
    for (int c = tid; c < VH2 * (VH3 / 8); c += kRes2Threads) {
      const int k = c / (VH3 / 8), cb = c % (VH3 / 8);
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(W3 + k * VH3 + 8 * cb));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(W3 + k * VH3 + 8 * cb) + 1);
      const float o[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      store8<FMT>(smem, kW3, kW3Piece, kRbW3, k, 8 * cb, o);
    }
    for (int c = tid; c < 16 * (VH1 / 8); c += kRes2Threads) {
      const int i = c / (VH1 / 8), jb = c % (VH1 / 8);
      float o[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) o[t] = i < N ? __ldg(W1 + i * VH1 + 8 * jb + t) : 0.f;
      store8<FMT>(smem, kW1, kW1Piece, kRbW1, i, 8 * jb, o);
    }
  }
  if (tid == 0) {
    for (int g = 0; g < kRes2Groups; ++g) mbar_init(bars + g, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tptr, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tptr;
  const uint32_t sb = smem_u32(smem) >> 4;
  const int64_t n_tiles = (a.B + TSM - 1) / TSM;

  {
    const int g = warp / kComputeWarps, wg = warp % kComputeWarps;
    const int q = wg & 3, hh = wg >> 2;
    const int sl = 32 * q + lane;                              // the state (TMEM lane) this thread owns
    const int c0 = 64 * hh;                                    // its 64 feature columns of a 128-wide layer
    const uint32_t tmg = tm + (uint32_t)g * kRes2Cols;         // the group's columns (lane 0): MMA operands
    const uint32_t tl = tmg + ((uint32_t)(32 * q) << 16);      // ... at this warp's lane quarter: loads / stores
    float* sV = reinterpret_cast<float*>(smem + kRes2Misc) + g * TSM;
    uint64_t* bar_mma = bars + g;
    const bool epi_warp = hh == 0, load_warp = hh == 1;
    const int64_t first = 2 * (int64_t)blockIdx.x + g, stride = 2 * (int64_t)gridDim.x;
    const int64_t n_iter = n_tiles > first ? (n_tiles - first + stride - 1) / stride : 0;
    uint32_t ph = 0;
    int step = 0;                                              // next GEMM of this group: 0..5, cyclic
    // end of a pass: the operand columns are written (tcgen05.wait::st), the group meets, one lane of its first warp
    // issues the group's next GEMM
    auto pass_done = [&]() {
      tc_wait_st();
      tc_fence_before();
      if (g == 0) asm volatile("bar.sync 3, 256;" ::: "memory");
      else asm volatile("bar.sync 4, 256;" ::: "memory");
      if (wg == 0) {
        tc_fence_after();
        if (elect_one()) {
          Res2Ops<FMT>::issue(step, tmg, sb);
          mma_commit(bar_mma);
        }
        __syncwarp();
      }
      step = step == 5 ? 0 : step + 1;
    };
    auto wait_mma = [&]() {
      mbar_wait(bar_mma, ph);
      ph ^= 1u;
      tc_fence_after();
    };
    // 16 accumulator columns -> 8 packed operand columns (hi and lo piece); bit position `sh` of `bits` onwards holds the
    // relu mask of these columns.  MASK_IN: apply the stored mask, else take relu and record the mask.
    // (16 columns at a time: the kernel runs 16 warps at 128 registers, and the epilogue warps carry a state's x, f, g)
    auto pack_store = [&](uint32_t dcol, uint32_t& bits, int sh, uint32_t pcol, auto mask_in) {
      uint32_t d[16], hi[8], lo[8];
      tmem_ld16(tl + r2D + dcol, d);
      tc_wait_ld();
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        float x0 = __uint_as_float(d[2 * t]), x1 = __uint_as_float(d[2 * t + 1]);
        if constexpr (decltype(mask_in)::value) {
          x0 = (bits >> (sh + 2 * t)) & 1u ? x0 : 0.f;
          x1 = (bits >> (sh + 2 * t + 1)) & 1u ? x1 : 0.f;
        } else {
          bits |= (x0 > 0.f ? 1u : 0u) << (sh + 2 * t);
          bits |= (x1 > 0.f ? 1u : 0u) << (sh + 2 * t + 1);
          x0 = fmaxf(x0, 0.f);
          x1 = fmaxf(x1, 0.f);
        }
        Fm<FMT>::pack2(x0, x1, hi[t], lo[t]);
      }
      tmem_st8(tl + r2Ahi + pcol, hi);
      tmem_st8(tl + r2Alo + pcol, lo);
    };
    // a 128-wide layer: this thread's 64 columns of D -> operand columns [c0 / 2, c0 / 2 + 32)
    auto layer_pass = [&](uint32_t& m0, uint32_t& m1, auto mask_in) {
      if constexpr (!decltype(mask_in)::value) { m0 = 0u; m1 = 0u; }
      pack_store(c0, m0, 0, c0 / 2, mask_in);
      pack_store(c0 + 16, m0, 16, c0 / 2 + 8, mask_in);
      pack_store(c0 + 32, m1, 0, c0 / 2 + 16, mask_in);
      pack_store(c0 + 48, m1, 16, c0 / 2 + 24, mask_in);
    };

    float xraw[N], z[N], fdyn[N], Gdyn[N * S::M];
    float xnext[N];                         // epilogue warps: next tile's raw states (issued one tile ahead)
    bool vnext = false;
    float dnext = 0.f, cnext = 1.f;
    float lz = 0.f, zz = 0.f, done = 0.f, cost = 1.f, Vsum = 0.f;
    float hjb_sum = 0.f, term_sum = 0.f;
    bool valid = false;
    int64_t idx = 0;
    uint32_t m1a = 0, m1b = 0, m2a = 0, m2b = 0;   // relu masks of layer 1 / 2: this thread's 2 x 32 columns
    // global loads of a tile's states are ISSUED one tile ahead (fetch_raw) and consumed later (to_error)
    auto fetch_raw = [&](int64_t tile) {
      idx = tile * TSM + sl;
      valid = idx < a.B;
      if (valid) load_row<N>(a.xs, idx, xraw);
      else {
#pragma unroll
        for (int i = 0; i < N; ++i) xraw[i] = a.xf[i];
      }
    };
    auto to_error = [&]() {
#pragma unroll
      for (int i = 0; i < N; ++i) z[i] = xraw[i] - a.xf[i];
      wrap_state<S>(z);
    };
    // normalised input h0 = (z - mu) / sd (vhjb.py:45) -> operand columns 0..7 (K = 16: N values, zero padding)
    auto store_h0 = [&]() {
      float h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = 0.f;
#pragma unroll
      for (int i = 0; i < N; ++i) h[i] = (z[i] - a.mean[i]) * a.inv_std[i];
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) Fm<FMT>::pack2(h[2 * t], h[2 * t + 1], hi[t], lo[t]);
      tmem_st8(tl + r2Ahi, hi);
      tmem_st8(tl + r2Alo, lo);
    };

    if (n_iter > 0) {
      if (load_warp) {
        fetch_raw(first);
        to_error();
        store_h0();
      }
      pass_done();                                              // -> G0 of the first tile
    }
    for (int64_t it = 0; it < n_iter; ++it) {
      const int64_t tile = first + it * stride;
      const bool more = it + 1 < n_iter;
      if (load_warp && more) fetch_raw(tile + stride);          // consumed in P6
      wait_mma();
      layer_pass(m1a, m1b, std::false_type{});                  // P1: h1 = relu(a1)
      pass_done();                                              // -> G1
      if (epi_warp) {   // what the epilogue needs from x alone (under G1; loads were issued a tile ahead)
        if (it == 0) {
          fetch_raw(tile);
          done = valid ? __ldg(a.dones + idx) : 0.f;
          cost = valid ? __ldg(a.costs + idx) : 1.f;
        } else {
          idx = tile * TSM + sl;
          valid = vnext;
          done = dnext;
          cost = cnext;
#pragma unroll
          for (int i = 0; i < N; ++i) xraw[i] = xnext[i];
        }
        to_error();
        float zi[N];
        to_internal<S>(a.sys, xraw, zi);
        typename S::Trig tr;
        S::trig(a.sys, zi, tr);
        S::fg(a.sys, zi, tr, fdyn, Gdyn);
        zz = 0.f;
        lz = 0.f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          zz = fmaf(z[i], z[i], zz);
          if constexpr (RFORM == HJB_RES_NORMALIZED) {
            float row = 0.f;
#pragma unroll
            for (int jj = 0; jj < N; ++jj) row = fmaf(a.Q[i * N + jj], z[jj], row);
            lz = fmaf(z[i], row, lz);
          }
        }
      }
      wait_mma();
      layer_pass(m2a, m2b, std::false_type{});                  // P2: h2 = relu(a2)
      pass_done();                                              // -> G2
      wait_mma();
      {                                                         // P3: V = |y|^2, gy = 2 y -> operand (K = 64)
        float v = 0.f;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t yv[16], hi[8], lo[8];
          tmem_ld16(tl + r2D + 32 * hh + 16 * hf, yv);
          tc_wait_ld();
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float y0 = __uint_as_float(yv[2 * t]), y1 = __uint_as_float(yv[2 * t + 1]);
            v = fmaf(y0, y0, v);
            v = fmaf(y1, y1, v);
            Fm<FMT>::pack2(2.f * y0, 2.f * y1, hi[t], lo[t]);
          }
          tmem_st8(tl + r2Ahi + 16 * hh + 8 * hf, hi);
          tmem_st8(tl + r2Alo + 16 * hh + 8 * hf, lo);
        }
        if (hh == 1) sV[sl] = v;
        if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
        else asm volatile("bar.sync 2, 256;" ::: "memory");
        if (hh == 0) Vsum = v + sV[sl];
      }
      pass_done();                                              // -> G3
      wait_mma();
      layer_pass(m2a, m2b, std::true_type{});                   // P4: g2 = b2 sigma'(a2)
      pass_done();                                              // -> G4
      wait_mma();
      if (epi_warp && more) {                                   // issue the next tile's loads: consumed at its top
        const int64_t nidx = (tile + stride) * TSM + sl;
        vnext = nidx < a.B;
        if (vnext) load_row<N>(a.xs, nidx, xnext);
        else {
#pragma unroll
          for (int i = 0; i < N; ++i) xnext[i] = a.xf[i];
        }
        dnext = vnext ? __ldg(a.dones + nidx) : 0.f;
        cnext = vnext ? __ldg(a.costs + nidx) : 1.f;
      }
      layer_pass(m1a, m1b, std::true_type{});                   // P5: g1 = b1 sigma'(a1)
      pass_done();                                              // -> G5
      wait_mma();
      if (epi_warp) {                                           // P6: control, residual, outputs
        uint32_t gv[16];
        tmem_ld16(tl + r2D, gv);
        tc_wait_ld();
        float g0v[N], pbar[N], Vbar;
#pragma unroll
        for (int i = 0; i < N; ++i) g0v[i] = __uint_as_float(gv[i]);
        state_epilogue<S, UFORM, RFORM, false>(a, g0v, Vsum, z, zz, lz, fdyn, Gdyn, done, cost, valid, idx, 0.f, 0.f, hjb_sum,
                                               term_sum, pbar, Vbar);
      }
      if (more) {
        if (load_warp) {                                        // next tile's input while the other warps run the epilogue
          to_error();
          store_h0();
        }
        pass_done();                                            // -> G0 of the next tile
      }
    }
    // ---- loss sums of this CTA (both groups) ----
    if (epi_warp) {
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        hjb_sum += __shfl_xor_sync(0xffffffffu, hjb_sum, sft);
        term_sum += __shfl_xor_sync(0xffffffffu, term_sum, sft);
      }
      if (lane == 0) {
        sLoss[(g * 4 + q) * 2] = hjb_sum;
        sLoss[(g * 4 + q) * 2 + 1] = term_sum;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    float* part = a.partial + (int64_t)blockIdx.x * a.pstride;
    float h = 0.f, tt = 0.f;
    for (int i = 0; i < 8; ++i) { h += sLoss[2 * i]; tt += sLoss[2 * i + 1]; }
    part[vhjb_param_count(N)] = h;
    part[vhjb_param_count(N) + 1] = tt;
    part[vhjb_param_count(N) + 2] = 0.f;
    part[vhjb_param_count(N) + 3] = 0.f;
  }
  if (warp == 0) tmem_dealloc(tm, 512);
}

}  // namespace tc
}  // namespace hjb

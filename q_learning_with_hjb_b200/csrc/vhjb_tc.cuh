// K2/K3 (tensor-core version) — the fused vhjb pass on tcgen05: value-MLP forward, input gradient dV/dx, optimal
// control, Hamiltonian residual and (GRAD) the full parameter gradient, 64 sampled states per tile, every GEMM of
// SURVEY.md 8a-V6 (all 17, the n-wide ones padded to 16) as tcgen05.mma with fp32 accumulation in TMEM.
//
// Precision: every operand x is split x = hi + lo into two 16-bit pieces and each product is issued as
// A_hi B_hi + A_lo B_hi + A_hi B_lo ("x3").  fp16 pieces (22 significant bits, weights pre-scaled by 64 so that
// the lo piece stays normal) for the residual-only pass -> fp32-grade V, p, u, r;  bf16 pieces (16 bits, fp32
// exponent range — adjoints span many decades) for the gradient pass.  (Mixed fp16 x bf16 operands are an illegal
// instruction on B200 — tests/cuda/umma_probe.cu.)
//
// Orientation: FEATURES on the 128 TMEM lanes, the tile's STATES on the MMA N dimension:
//   chain GEMMs      D[j][s] = sum_k Wt[j][k] X[k][s]     A = a weight matrix (smem, resident), B = activations (smem)
//   weight gradients D[i][j] = sum_s X[i][s] Y[j][s]      both operands are activation buffers, K = states
//   per-state GEMMs  D[s][c] = sum_k X[k][s] W[k][c]      (y = h2 W3, dV/dz = g1 W1^T, ...): M = 128 is issued over
//                    the 64 valid state rows; lanes 64..127 accumulate rows read past the tile — never read back.
// One activation buffer X[feature][state] (8x8 core matrices, no swizzle) serves as MN-major B operand of the chain,
// as K-major A or B operand of the weight gradients and as MN-major A operand of the per-state GEMMs: every
// intermediate is written to shared memory exactly once.  Measured on B200 (tests/cuda/umma_probe.cu): an SS-mode MMA
// (M=128, K=16) costs 32 + N/4 cycles — the 4 KB A tile is re-read from shared memory — so N = 64 runs at 48 cycles.
//
// Roles: warps 0..7 run the element-wise passes between GEMMs (TMEM -> registers -> activation / mask / split ->
// shared memory); lane l of warp w owns feature 32 (w % 4) + l and the states 32 (w / 4) .. +31 of the tile; warps
// 0, 1 additionally own one state each for the per-state epilogue (identical arithmetic to vhjb_simt.cuh).  Warp 8
// issues every MMA from one elected lane.  Passes and GEMM groups alternate through two mbarriers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "umma.cuh"
#include "vhjb_simt.cuh"

namespace hjb {
namespace tc {
using namespace umma;

constexpr int TS = 64;  // states per tile
constexpr int kComputeWarps = 8;
constexpr int kThreads = 32 * (kComputeWarps + 1);

// ---- shared-memory map (bytes); every 16-bit matrix is stored as [hi piece | lo piece] ----
constexpr uint32_t kW2 = 0, kW2Piece = VH1 * VH2 * 2;
constexpr uint32_t kW3 = kW2 + 2 * kW2Piece, kW3Piece = VH2 * VH3 * 2;
constexpr uint32_t kW1 = kW3 + 2 * kW3Piece, kW1Piece = 16 * VH1 * 2;
constexpr uint32_t kF0 = kW1 + 2 * kW1Piece, kFPiece = 128 * TS * 2;
constexpr uint32_t kF1 = kF0 + 2 * kFPiece, kF2 = kF1 + 2 * kFPiece;
constexpr uint32_t kY0 = kF2 + 2 * kFPiece, kYPiece = TS * VH3 * 2;
constexpr uint32_t kH0 = kY0 + 2 * kYPiece, kHPiece = TS * 16 * 2;
constexpr uint32_t kG0 = kH0 + 2 * kHPiece;
constexpr uint32_t kMisc = kG0 + 2 * kHPiece;  // float sV[64], sVb[64]; u64 bars[2]; u32 tmem; float sF[64], sYm[64]
constexpr uint32_t kSmemBytes = kMisc + 1152;
static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");

// row-block strides of the core-matrix layouts: X[R][C] -> (C / 8) * 128 bytes
constexpr uint32_t kRbW2 = (VH2 / 8) * 128, kRbW3 = (VH3 / 8) * 128, kRbW1 = (VH1 / 8) * 128, kRbF = (TS / 8) * 128,
                   kRbY = (VH3 / 8) * 128, kRbH = (16 / 8) * 128;

// ---- TMEM columns (512 allocated) ----
constexpr uint32_t cW2g = 0, cW3g = 128, cW1g = 192, cG0 = 208, cA1 = 224, cA2 = 288, cY = 352, cWk = 416;

struct Op { uint32_t addr, piece, lbo, sbo, kadv; };
// storage X[R][C]; K-major use: MN = row, K = column.  MN-major use: K = row, MN = column.
__device__ __forceinline__ constexpr Op kmaj(uint32_t addr, uint32_t piece, uint32_t rb) { return Op{addr, piece, 128u, rb, 256u}; }
__device__ __forceinline__ constexpr Op mnmaj(uint32_t addr, uint32_t piece, uint32_t rb) { return Op{addr, piece, rb, 128u, 2u * rb}; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// D (+)= A B^T as three passes over the 16-bit pieces; acc0 = accumulate flag of the very first MMA
// sbd = (shared-memory base address) >> 4; every operand offset is a multiple of 16 bytes
template <int KSTEPS>
__device__ __forceinline__ void gemm3(uint32_t sbd, uint32_t d, const Op A, const Op B, uint32_t idesc, uint32_t acc0) {
  const uint32_t a_hi = (A.sbo >> 4) | (1u << 14), b_hi = (B.sbo >> 4) | (1u << 14);
#pragma unroll
  for (int pr = 0; pr < 3; ++pr) {
    const uint32_t aa = ((A.addr + (pr == 1 ? A.piece : 0u)) >> 4) | ((A.lbo >> 4) << 16);
    const uint32_t bb = ((B.addr + (pr == 2 ? B.piece : 0u)) >> 4) | ((B.lbo >> 4) << 16);
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k)
      mma_ss2(d, sbd + (aa + ((k * A.kadv) >> 4)), a_hi, sbd + (bb + ((k * B.kadv) >> 4)), b_hi, idesc, (pr | k) ? 1u : acc0);
  }
}

template <int FMT> struct Fm;
template <> struct Fm<kBF16> {
  static constexpr float ws = 1.f, iws = 1.f;
  static __device__ __forceinline__ void pack2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
};
template <> struct Fm<kF16> {
  static constexpr float ws = 64.f, iws = 1.f / 64.f;
  static __device__ __forceinline__ void pack2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
};

// 8 consecutive columns [c0, c0 + 8) of row r of X[R][C] (row-block stride rb) -> one 16-byte chunk per piece
template <int FMT>
__device__ __forceinline__ void store8(uint8_t* smem, uint32_t buf, uint32_t piece, uint32_t rb, int r, int c0, const float* o) {
  uint4 hi, lo;
  Fm<FMT>::pack2(o[0], o[1], hi.x, lo.x);
  Fm<FMT>::pack2(o[2], o[3], hi.y, lo.y);
  Fm<FMT>::pack2(o[4], o[5], hi.z, lo.z);
  Fm<FMT>::pack2(o[6], o[7], hi.w, lo.w);
  const uint32_t off = buf + (uint32_t)(r >> 3) * rb + ((uint32_t)(c0 >> 3) << 7) + ((uint32_t)(r & 7) << 4);
  *reinterpret_cast<uint4*>(smem + off) = hi;
  *reinterpret_cast<uint4*>(smem + off + piece) = lo;
}

template <class S, int ACT, int UFORM, int RFORM, bool GRAD, int FMT>
__global__ void __launch_bounds__(kThreads, 1) vhjb_tc_kernel(const __grid_constant__ VhjbArgs a) {
  static_assert(ACT == HJB_ACT_RELU, "tensor-core path: relu value nets (sigma'' = 0)");
  constexpr int N = S::N, M = S::M;
  static_assert(N <= 16, "state dimension padded to one K = 16 step");
  constexpr float iws = Fm<FMT>::iws;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sV = reinterpret_cast<float*>(smem + kMisc);
  float* sVb = sV + TS;
  uint64_t* bar_pass = reinterpret_cast<uint64_t*>(smem + kMisc + 512);
  uint64_t* bar_mma = bar_pass + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + kMisc + 544);
  float* sF = reinterpret_cast<float*>(smem + kMisc + 576);    // 2^f_s: weight-gradient scale of the forward operands
  float* sYm = reinterpret_cast<float*>(smem + kMisc + 832);   // max_c |2 y_c| of the state (second column half)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- weights -> shared memory (scaled, split, core-matrix layout), once per CTA ----
  {
    const float* W1 = a.params;
    const float* W2 = W1 + N * VH1;
    const float* W3 = W2 + VH1 * VH2;
    constexpr float ws = Fm<FMT>::ws;
    for (int c = tid; c < VH1 * (VH2 / 8); c += kThreads) {
      const int k = c / (VH2 / 8), jb = c % (VH2 / 8);
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(W2 + k * VH2 + 8 * jb));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(W2 + k * VH2 + 8 * jb) + 1);
      const float o[8] = {v0.x * ws, v0.y * ws, v0.z * ws, v0.w * ws, v1.x * ws, v1.y * ws, v1.z * ws, v1.w * ws};
      store8<FMT>(smem, kW2, kW2Piece, kRbW2, k, 8 * jb, o);
    }
    for (int c = tid; c < VH2 * (VH3 / 8); c += kThreads) {
      const int k = c / (VH3 / 8), cb = c % (VH3 / 8);
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(W3 + k * VH3 + 8 * cb));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(W3 + k * VH3 + 8 * cb) + 1);
      const float o[8] = {v0.x * ws, v0.y * ws, v0.z * ws, v0.w * ws, v1.x * ws, v1.y * ws, v1.z * ws, v1.w * ws};
      store8<FMT>(smem, kW3, kW3Piece, kRbW3, k, 8 * cb, o);
    }
    for (int c = tid; c < 16 * (VH1 / 8); c += kThreads) {
      const int i = c / (VH1 / 8), jb = c % (VH1 / 8);
      float o[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) o[t] = i < N ? __ldg(W1 + i * VH1 + 8 * jb + t) * ws : 0.f;
      store8<FMT>(smem, kW1, kW1Piece, kRbW1, i, 8 * jb, o);
    }
  }
  if (tid == 0) {
    mbar_init(bar_pass, kComputeWarps);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == kComputeWarps) tmem_alloc(tptr, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tptr;
  const uint32_t sb = smem_u32(smem) >> 4;
  const int64_t n_iter = a.n_tiles > blockIdx.x ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  float* part = a.partial + (int64_t)blockIdx.x * a.pstride;

  if (warp == kComputeWarps) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idN64_mn_mn = idesc_f16(128, TS, FMT, FMT, 1, 1), idN64_mn_k = idesc_f16(128, TS, FMT, FMT, 1, 0),
                       idN64_k_k = idesc_f16(128, TS, FMT, FMT, 0, 0), idN64_k_mn = idesc_f16(128, TS, FMT, FMT, 0, 1),
                       idY_mn_mn = idesc_f16(128, VH3, FMT, FMT, 1, 1), idN16_mn_k = idesc_f16(128, 16, FMT, FMT, 1, 0),
                       idN16_k_mn = idesc_f16(128, 16, FMT, FMT, 0, 1), idN128_k_k = idesc_f16(128, VH2, FMT, FMT, 0, 0),
                       idW3g_k_mn = idesc_f16(128, VH3, FMT, FMT, 0, 1);
    constexpr Op W2_mn = mnmaj(kW2, kW2Piece, kRbW2), W2_k = kmaj(kW2, kW2Piece, kRbW2);
    constexpr Op W3_mn = mnmaj(kW3, kW3Piece, kRbW3), W3_k = kmaj(kW3, kW3Piece, kRbW3);
    constexpr Op W1_mn = mnmaj(kW1, kW1Piece, kRbW1), W1_k = kmaj(kW1, kW1Piece, kRbW1);
    constexpr Op F0_mn = mnmaj(kF0, kFPiece, kRbF), F0_k = kmaj(kF0, kFPiece, kRbF);
    constexpr Op F1_mn = mnmaj(kF1, kFPiece, kRbF), F1_k = kmaj(kF1, kFPiece, kRbF);
    constexpr Op F2_mn = mnmaj(kF2, kFPiece, kRbF), F2_k = kmaj(kF2, kFPiece, kRbF);
    constexpr Op Y0_mn = mnmaj(kY0, kYPiece, kRbY), Y0_k = kmaj(kY0, kYPiece, kRbY);
    constexpr Op H0_mn = mnmaj(kH0, kHPiece, kRbH), H0_k = kmaj(kH0, kHPiece, kRbH);
    constexpr Op G0_mn = mnmaj(kG0, kHPiece, kRbH), G0_k = kmaj(kG0, kHPiece, kRbH);
    uint32_t ph = 0;
#define HJB_TC_GROUP(...)                 \
  do {                                    \
    mbar_wait(bar_pass, ph);              \
    ph ^= 1u;                             \
    tc_fence_after();                     \
    if (elect_one()) {                    \
      __VA_ARGS__;                        \
      mma_commit(bar_mma);                \
    }                                     \
    __syncwarp();                         \
  } while (0)
    for (int64_t it = 0; it < n_iter; ++it) {
      const uint32_t acc = it > 0 ? 1u : 0u;
      // G0: a1^T = W1^T h0^T
      HJB_TC_GROUP(gemm3<1>(sb, tm + cA1, W1_mn, H0_k, idN64_mn_k, 0u));
      // G1: a2^T = W2^T h1^T
      HJB_TC_GROUP(gemm3<8>(sb, tm + cA2, W2_mn, F0_mn, idN64_mn_mn, 0u));
      // G2: y = h2 W3                      (lanes = states)
      HJB_TC_GROUP(gemm3<8>(sb, tm + cY, F1_mn, W3_mn, idY_mn_mn, 0u));
      // G3: b2^T = W3 gy^T
      HJB_TC_GROUP(gemm3<4>(sb, tm + cWk, W3_k, Y0_k, idN64_k_k, 0u));
      // G4: b1^T = W2 g2^T
      HJB_TC_GROUP(gemm3<8>(sb, tm + cWk, W2_k, F0_mn, idN64_k_mn, 0u));
      // G5: g0 = g1 W1^T                   (lanes = states, 16 columns)
      HJB_TC_GROUP(gemm3<8>(sb, tm + cG0, F1_mn, W1_k, idN16_mn_k, 0u));
      if constexpr (GRAD) {
        // G6: W1bar^T += g1^T g0bar ; g1bar^T = W1^T g0bar^T
        HJB_TC_GROUP(gemm3<4>(sb, tm + cW1g, F1_k, G0_mn, idN16_k_mn, acc);
                     gemm3<1>(sb, tm + cWk, W1_mn, G0_k, idN64_mn_k, 0u));
        // G7: g2bar^T = W2^T b1bar^T ; W2bar += b1bar^T g2
        HJB_TC_GROUP(gemm3<8>(sb, tm + cWk, W2_mn, F2_mn, idN64_mn_mn, 0u);
                     gemm3<4>(sb, tm + cW2g, F2_k, F0_k, idN128_k_k, acc));
        // G8: gybar = b2bar W3 (lanes = states) ; W3bar += b2bar^T gy
        HJB_TC_GROUP(gemm3<8>(sb, tm + cWk, F1_mn, W3_mn, idY_mn_mn, 0u);
                     gemm3<4>(sb, tm + cW3g, F1_k, Y0_mn, idW3g_k_mn, acc));
        // G9: a2bar_pre^T = W3 ybar^T ; W3bar += h2^T ybar
        HJB_TC_GROUP(gemm3<4>(sb, tm + cWk, W3_k, Y0_k, idN64_k_k, 0u);
                     gemm3<4>(sb, tm + cW3g, F0_k, Y0_mn, idW3g_k_mn, 1u));
        // G10: a1bar_pre^T = W2 a2bar^T ; W2bar += h1^T a2bar
        HJB_TC_GROUP(gemm3<8>(sb, tm + cWk, W2_k, F2_mn, idN64_k_mn, 0u);
                     gemm3<4>(sb, tm + cW2g, F1_k, F2_k, idN128_k_k, 1u));
        // G11: W1bar^T += a1bar^T h0
        HJB_TC_GROUP(gemm3<4>(sb, tm + cW1g, F0_k, H0_mn, idN16_k_mn, 1u));
      }
    }
#undef HJB_TC_GROUP
  } else {
    // ================================ element-wise passes ================================
    const int q = warp & 3, hh = warp >> 2;
    const int j = 32 * q + lane;                 // feature == TMEM lane
    const int sc0 = 32 * hh;                     // first state column of this thread
    const uint32_t tl = tm + ((uint32_t)(32 * q) << 16);
    const bool state_warp = q < 2;               // lanes 0..63 hold state rows of the per-state GEMMs
    const bool epi_warp = state_warp && hh == 0; // owns state j for the epilogue
    uint32_t ph = 0;
    auto pass_done = [&]() {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pass);
    };
    auto wait_mma = [&]() {
      mbar_wait(bar_mma, ph);
      ph ^= 1u;
      tc_fence_after();
    };
    // feature pass: out(j, s) = fn(D(j, s), stash(j, s)) for the 32 states of this thread -> X[feature][state]
    auto feature_pass = [&](uint32_t cD, uint32_t cStash, uint32_t buf, auto fn) {
      uint32_t d[32], st[32];
      tmem_ld32(tl + cD + sc0, d);
      tmem_ld32(tl + cStash + sc0, st);
      tc_wait_ld();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = fn(__uint_as_float(d[8 * g + t]), __uint_as_float(st[8 * g + t]));
        store8<FMT>(smem, buf, kFPiece, kRbF, j, sc0 + 8 * g, o);
      }
    };
    auto act_pass = [&](uint32_t cStash, uint32_t buf, auto scaled) {   // h = sigma(a) [* 2^f_s]
      uint32_t st[32];
      tmem_ld32(tl + cStash + sc0, st);
      tc_wait_ld();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = act_f<ACT>(__uint_as_float(st[8 * g + t]) * iws);
        if constexpr (decltype(scaled)::value) {
          const float4 f0 = *reinterpret_cast<const float4*>(sF + sc0 + 8 * g), f1 = *reinterpret_cast<const float4*>(sF + sc0 + 8 * g + 4);
          o[0] *= f0.x; o[1] *= f0.y; o[2] *= f0.z; o[3] *= f0.w;
          o[4] *= f1.x; o[5] *= f1.y; o[6] *= f1.z; o[7] *= f1.w;
        }
        store8<FMT>(smem, buf, kFPiece, kRbF, j, sc0 + 8 * g, o);
      }
    };
    // in-place X[feature][state] *= 2^f_s (exact: power-of-two scaling of both pieces)
    auto rescale_feature_buf = [&](uint32_t buf) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 f0 = *reinterpret_cast<const float4*>(sF + sc0 + 8 * g), f1 = *reinterpret_cast<const float4*>(sF + sc0 + 8 * g + 4);
        const __half2 m0 = __floats2half2_rn(f0.x, f0.y), m1 = __floats2half2_rn(f0.z, f0.w), m2 = __floats2half2_rn(f1.x, f1.y),
                      m3 = __floats2half2_rn(f1.z, f1.w);
        const uint32_t off = buf + (uint32_t)(j >> 3) * kRbF + ((uint32_t)((sc0 + 8 * g) >> 3) << 7) + ((uint32_t)(j & 7) << 4);
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
          uint4* ptr = reinterpret_cast<uint4*>(smem + off + pc * kFPiece);
          uint4 v = *ptr;
          __half2 h0 = *reinterpret_cast<__half2*>(&v.x), h1 = *reinterpret_cast<__half2*>(&v.y), h2 = *reinterpret_cast<__half2*>(&v.z),
                  h3 = *reinterpret_cast<__half2*>(&v.w);
          h0 = __hmul2(h0, m0); h1 = __hmul2(h1, m1); h2 = __hmul2(h2, m2); h3 = __hmul2(h3, m3);
          v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
          v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
          *ptr = v;
        }
      }
    };
    auto masked = [](float d, float st) { return d * (iws * act_d1<ACT>(st * iws)); };

    float xraw[N], z[N];
    float Vsum = 0.f, Vbar = 0.f, gymax = 0.f;
    float hjb_sum = 0.f, term_sum = 0.f, sat_count = 0.f;
    float inv_norm0 = 0.f, inv_norm1 = 0.f;
    // Gradient pass, fp16 range management.  The reverse pass of one state is linear in its adjoint seeds
    // (g0bar, Vbar), whose size varies by many decades over a batch (1 / (l + eps), 1 / (cost + eps), 1 / batch).
    // Per state: seeds = 2^k_s x (numbers in [1/2, 1)); the adjoint chain carries 2^(a_s - k_s) x true values, the
    // forward partners of the weight-gradient GEMMs carry 2^f_s, with a_s + f_s = k_s - E and one launch-wide
    // exponent E = exponent(typical seed weight) + 12, so the TMEM accumulators hold 2^-E x gradient.  Both operand
    // families stay inside fp16's range for every seed up to 2^(14+E); smaller seeds lose bits gradually (they
    // contribute proportionally less); larger ones (|x - xf| and |u - uf| below ~1e-4 with the default eps) are
    // under-weighted and COUNTED in partial[P + 2] (hjb_vhjb_saturation): nothing overflows silently.
    int expE = 0;
    if constexpr (GRAD) {
      inv_norm0 = 1.0f / __ldg(a.norm);
      inv_norm1 = 1.0f / __ldg(a.norm + 1);
      const float wt = RFORM == HJB_RES_NORMALIZED ? fmaxf(inv_norm0, fabsf(a.reg) * inv_norm1) : inv_norm0;
      expE = (int)((__float_as_uint(wt) >> 23) & 0xffu) - 126 + 12;
      expE = max(-100, min(100, expE));
    }
    bool valid = false;
    int64_t idx = 0;
    int mark = 0;
    auto tmark = [&](int64_t it) {
      if (a.dbg != nullptr && blockIdx.x == 0 && it == 2 && tid == 0) a.dbg[mark++] = clock64();
    };

    // P0: states of the tile -> error coordinates, normalised input (vhjb.py:39, :45) -> H0[s][16]
    auto load_tile = [&](int64_t tile) {
      if (epi_warp) {
        idx = tile * TS + j;
        valid = idx < a.B;
        if (valid) load_row<N>(a.xs, idx, xraw);
        else {
#pragma unroll
          for (int i = 0; i < N; ++i) xraw[i] = a.xf[i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) z[i] = xraw[i] - a.xf[i];
        wrap_state<S>(z);
        float h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = 0.f;
#pragma unroll
        for (int i = 0; i < N; ++i) h[i] = (z[i] - a.mean[i]) * a.inv_std[i];
        store8<FMT>(smem, kH0, kHPiece, kRbH, j, 0, h);
        store8<FMT>(smem, kH0, kHPiece, kRbH, j, 8, h + 8);
      }
    };

    if (n_iter > 0) load_tile(blockIdx.x);
    for (int64_t it = 0; it < n_iter; ++it) {
      tmark(it);
      pass_done();                                              // -> G0
      // P1: h1 = sigma(a1) -> F0
      wait_mma();
      tmark(it);
      act_pass(cA1, kF0, std::false_type{});
      tmark(it);
      pass_done();                                              // -> G1
      // P2: h2 = sigma(a2) -> F1
      wait_mma();
      tmark(it);
      act_pass(cA2, kF1, std::false_type{});
      tmark(it);
      pass_done();                                              // -> G2
      // P3 (state warps): V = |y|^2 (+ eps_s |z|^2 later), gy = 2 y -> Y0[s][c]
      wait_mma();
      tmark(it);
      if (state_warp) {
        uint32_t yv[32];
        tmem_ld32(tl + cY + sc0, yv);
        tc_wait_ld();
        float v = 0.f, ym = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float o[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float y = __uint_as_float(yv[8 * g + t]) * iws;
            v = fmaf(y, y, v);
            o[t] = 2.f * y;
            ym = fmaxf(ym, fabsf(o[t]));
          }
          store8<FMT>(smem, kY0, kYPiece, kRbY, j, sc0 + 8 * g, o);
        }
        if (hh == 1) { sV[j] = v; sYm[j] = ym; }
        asm volatile("bar.sync 1, 128;" ::: "memory");          // warps 0, 1, 4, 5
        if (hh == 0) { Vsum = v + sV[j]; gymax = fmaxf(ym, sYm[j]); }
      }
      tmark(it);
      pass_done();                                              // -> G3
      // P4: g2 = b2 sigma'(a2) -> F0
      wait_mma();
      tmark(it);
      feature_pass(cWk, cA2, kF0, masked);
      tmark(it);
      pass_done();                                              // -> G4
      // P5: g1 = b1 sigma'(a1) -> F1
      wait_mma();
      tmark(it);
      feature_pass(cWk, cA1, kF1, masked);
      tmark(it);
      pass_done();                                              // -> G5
      // P6 (epilogue warps): control, Hamiltonian residual, adjoint seeds (vhjb.py:204-253)
      wait_mma();
      tmark(it);
      if (epi_warp) {
        uint32_t gv[16];
        tmem_ld16(tl + cG0, gv);
        tc_wait_ld();
        float p[N];
        float V = Vsum, zz = 0.f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          zz = fmaf(z[i], z[i], zz);
          p[i] = fmaf(__uint_as_float(gv[i]) * iws, a.inv_std[i], 2.f * a.eps_s * z[i]);
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
            for (int jj = 0; jj < M; ++jj) ur = fmaf(-0.5f * a.Rinv[k * M + jj], c[jj], ur);
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
        float r, pbar[N];
#pragma unroll
        for (int i = 0; i < N; ++i) pbar[i] = 0.f;
        Vbar = 0.f;
        if constexpr (RFORM == HJB_RES_NORMALIZED) {
          float l = 0.f;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            float row = 0.f;
#pragma unroll
            for (int jj = 0; jj < N; ++jj) row = fmaf(a.Q[i * N + jj], z[jj], row);
            l = fmaf(z[i], row, l);
          }
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
            Vbar = valid ? a.reg * done * inv_norm1 * sign0(tq) / (cost + a.eps) : 0.f;
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
        if constexpr (GRAD) {
          float gb[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) gb[i] = 0.f;
          float ms = fabsf(Vbar) * gymax;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            gb[i] = pbar[i] * a.inv_std[i];    // g0-bar
            ms = fmaxf(ms, fabsf(gb[i]));
          }
          const int eb = (int)((__float_as_uint(ms) >> 23) & 0xffu);       // ms = f 2^(eb - 126), f in [1/2, 1)
          float lam = 0.f, fs = 0.f;
          if (eb > 8 && eb < 226) {
            const int ks = eb - 126, ts = ks - expE;
            const int as = max(-8, min(8, ts >> 1));
            const int fe = max(-16, min(6, ts - as));
            if (ts - as > 6) sat_count += 1.f;                               // seed beyond 2^(14+E): under-weighted
            lam = __uint_as_float((uint32_t)(as - ks + 127) << 23);          // 2^(a_s - k_s), exponent in [19, 252]
            fs = exp2f((float)fe);
          }
#pragma unroll
          for (int i = 0; i < N; ++i) gb[i] *= lam;
          store8<FMT>(smem, kG0, kHPiece, kRbH, j, 0, gb);
          store8<FMT>(smem, kG0, kHPiece, kRbH, j, 8, gb + 8);
          sVb[j] = Vbar * lam;
          sF[j] = fs;
          // h0 * 2^f_s for the W1bar GEMM of step 11 (G0 consumed the unscaled copy long ago)
          float h[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) h[i] = 0.f;
#pragma unroll
          for (int i = 0; i < N; ++i) h[i] = (z[i] - a.mean[i]) * a.inv_std[i] * fs;
          store8<FMT>(smem, kH0, kHPiece, kRbH, j, 0, h);
          store8<FMT>(smem, kH0, kHPiece, kRbH, j, 8, h + 8);
        }
      }
      if constexpr (!GRAD) {
        if (it + 1 < n_iter) load_tile(blockIdx.x + (it + 1) * (int64_t)gridDim.x);   // H0 is free: G0 completed long ago
      } else {
        asm volatile("bar.sync 3, 256;" ::: "memory");          // sF published by the epilogue warps
        rescale_feature_buf(kF1);                               // g1 2^f_s  (W1bar, step 6)
        rescale_feature_buf(kF0);                               // g2 2^f_s  (W2bar, step 7)
        if (state_warp) {                                       // gy 2^f_s  (W3bar, step 8): rows = states
          const __half2 m = __float2half2_rn(sF[j]);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t off = kY0 + (uint32_t)(j >> 3) * kRbY + ((uint32_t)((sc0 + 8 * g) >> 3) << 7) + ((uint32_t)(j & 7) << 4);
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
              uint4* ptr = reinterpret_cast<uint4*>(smem + off + pc * kYPiece);
              uint4 v = *ptr;
              __half2 h0 = *reinterpret_cast<__half2*>(&v.x), h1 = *reinterpret_cast<__half2*>(&v.y),
                      h2 = *reinterpret_cast<__half2*>(&v.z), h3 = *reinterpret_cast<__half2*>(&v.w);
              h0 = __hmul2(h0, m); h1 = __hmul2(h1, m); h2 = __hmul2(h2, m); h3 = __hmul2(h3, m);
              v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
              v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
              *ptr = v;
            }
          }
        }
        tmark(it);
        pass_done();                                            // -> G6
        // P7: b1bar = g1bar sigma'(a1) -> F2
        wait_mma();
      tmark(it);
        feature_pass(cWk, cA1, kF2, masked);
        tmark(it);
      pass_done();                                            // -> G7
        // P8: b2bar = g2bar sigma'(a2) -> F1
        wait_mma();
      tmark(it);
        feature_pass(cWk, cA2, kF1, masked);
        tmark(it);
      pass_done();                                            // -> G8
        // P9: (state warps) ybar = 2 gybar + 2 y Vbar -> Y0 ; (all) h2 = sigma(a2) -> F0
        wait_mma();
      tmark(it);
        if (state_warp) {
          uint32_t gv[32], yv[32];
          tmem_ld32(tl + cWk + sc0, gv);
          tmem_ld32(tl + cY + sc0, yv);
          tc_wait_ld();
          const float vb = sVb[j];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float o[8];
#pragma unroll
            for (int t = 0; t < 8; ++t)
              o[t] = (2.f * iws) * fmaf(__uint_as_float(yv[8 * g + t]), vb, __uint_as_float(gv[8 * g + t]));
            store8<FMT>(smem, kY0, kYPiece, kRbY, j, sc0 + 8 * g, o);
          }
        }
        act_pass(cA2, kF0, std::true_type{});
        tmark(it);
      pass_done();                                            // -> G9
        // P10: a2bar = a2bar_pre sigma'(a2) -> F2 ; h1 = sigma(a1) -> F1
        wait_mma();
      tmark(it);
        feature_pass(cWk, cA2, kF2, masked);
        act_pass(cA1, kF1, std::true_type{});
        tmark(it);
      pass_done();                                            // -> G10
        // P11: a1bar = a1bar_pre sigma'(a1) -> F0
        wait_mma();
      tmark(it);
        feature_pass(cWk, cA1, kF0, masked);
        tmark(it);
      pass_done();                                            // -> G11
        // next tile's P0 overwrites H0, which G11 reads
        wait_mma();
      tmark(it);
        if (it + 1 < n_iter) load_tile(blockIdx.x + (it + 1) * (int64_t)gridDim.x);
      }
    }

    // ---- per-CTA partials: weight-gradient accumulators (TMEM) and the two loss sums ----
    if constexpr (GRAD) {
      const int P1 = N * VH1;
      const float unscale = exp2f((float)expE);
      if (n_iter > 0) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {   // W2bar[k = lane][col]: this thread's 64 columns
          uint32_t v[32];
          tmem_ld32(tl + cW2g + 64 * hh + 32 * half, v);
          tc_wait_ld();
          float4* dst = reinterpret_cast<float4*>(part + P1 + j * VH2 + 64 * hh + 32 * half);
#pragma unroll
          for (int t = 0; t < 8; ++t)
            dst[t] = make_float4(__uint_as_float(v[4 * t]) * unscale, __uint_as_float(v[4 * t + 1]) * unscale,
                                 __uint_as_float(v[4 * t + 2]) * unscale, __uint_as_float(v[4 * t + 3]) * unscale);
        }
        {
          uint32_t v[32];
          tmem_ld32(tl + cW3g + 32 * hh, v);
          tc_wait_ld();
          float4* dst = reinterpret_cast<float4*>(part + P1 + VH1 * VH2 + j * VH3 + 32 * hh);
#pragma unroll
          for (int t = 0; t < 8; ++t)
            dst[t] = make_float4(__uint_as_float(v[4 * t]) * unscale, __uint_as_float(v[4 * t + 1]) * unscale,
                                 __uint_as_float(v[4 * t + 2]) * unscale, __uint_as_float(v[4 * t + 3]) * unscale);
        }
        if (hh == 0) {
          uint32_t v[16];
          tmem_ld16(tl + cW1g, v);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < N; ++i) part[i * VH1 + j] = __uint_as_float(v[i]) * unscale;
        }
      } else {
        for (int i = tid; i < vhjb_param_count(N); i += 32 * kComputeWarps) part[i] = 0.f;
      }
    }
    if (warp < 2) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        hjb_sum += __shfl_xor_sync(0xffffffffu, hjb_sum, s);
        term_sum += __shfl_xor_sync(0xffffffffu, term_sum, s);
        sat_count += __shfl_xor_sync(0xffffffffu, sat_count, s);
      }
      if (lane == 0) {
        sV[4 * warp] = hjb_sum;
        sV[4 * warp + 1] = term_sum;
        sV[4 * warp + 2] = sat_count;
      }
      asm volatile("bar.sync 2, 64;" ::: "memory");
      if (tid == 0) {
        part[vhjb_param_count(N)] = sV[0] + sV[4];
        part[vhjb_param_count(N) + 1] = sV[1] + sV[5];
        part[vhjb_param_count(N) + 2] = sV[2] + sV[6];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kComputeWarps) tmem_dealloc(tm, 512);
}

template <class S, int ACT, int UFORM, int RFORM>
inline cudaError_t launch_vhjb_tc_variant(const VhjbArgs& a, const VhjbLaunch& l, cudaStream_t st) {
  cudaError_t e;
  if (l.grad) {
    auto k = vhjb_tc_kernel<S, ACT, UFORM, RFORM, true, kF16>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    k<<<l.grid, kThreads, kSmemBytes, st>>>(a);
  } else {
    auto k = vhjb_tc_kernel<S, ACT, UFORM, RFORM, false, kF16>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    k<<<l.grid, kThreads, kSmemBytes, st>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace tc

// tensor-core launchers (relu nets); cudaErrorNotSupported when the variant is not compiled
cudaError_t vhjb_tc_launch_linear21(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st);
cudaError_t vhjb_tc_launch_cartpole(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st);
cudaError_t vhjb_tc_launch_quad2d(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st);
cudaError_t vhjb_tc_launch_quad10d(const VhjbArgs& a, const VhjbLaunch& l, int act, int uform, int rform, cudaStream_t st);

}  // namespace hjb

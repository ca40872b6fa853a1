// tcgen05 / TMEM / mbarrier building blocks (sm_100a inline PTX) for the tensor-core vhjb kernels.
//
// Conventions used throughout:
//   * cta_group::1, UMMA M = 128: accumulator row i lives in TMEM lane i; a 32-bit TMEM address is
//     (lane << 16) | column; warp w of a warpgroup may only touch lanes [32 (w % 4), 32 (w % 4) + 32).
//   * kind::f16 with bf16 inputs, fp32 accumulation; fp32 accuracy is recovered with the BF16x3 split
//     x = hi + lo (hi = bf16(x), lo = bf16(x - hi)):  A B ~= A_hi B_hi + A_hi B_lo + A_lo B_hi.
//   * shared-memory operands use the NO-SWIZZLE canonical layout: 8 x 8 bf16 "core matrices" (8 rows of 16 B,
//     128 B each), core (rb, cb) of a logical [R][C] matrix at byte offset (rb * C/8 + cb) * 128.  The same bytes
//     serve as a K-major operand (MN = R, K = C: SBO = C/8 * 128, LBO = 128) and as an MN-major operand
//     (K = R, MN = C: SBO = 128, LBO = C/8 * 128) — so one copy of a weight matrix feeds both W and W^T GEMMs.
//   * the A operand may instead live in TMEM: row m in lane m, 32-bit column c holds (k = 2c | k = 2c + 1 << 16).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hjb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}

// ---- proxies / fences ----------------------------------------------------------------------------------
// generic-proxy st.shared -> visible to the async proxy (UMMA operand fetch)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM allocation (one full warp; ncols power of two >= 32) ---------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// ---- descriptors ---------------------------------------------------------------------------------------
// no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor: version = 1 at bit 46, layout_type 0)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D fp32; A/B format 0 = fp16, 1 = bf16 (cute::UMMA::InstrDescriptor:
// c_format [4,6), a_format [7,10), b_format [10,13), a_major 15, b_major 16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29))
constexpr int kF16 = 0, kBF16 = 1;
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_fmt, int b_fmt, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return idesc_f16(M, N, kBF16, kBF16, a_mn_major, b_mn_major);
}

// ---- MMA issue (ONE thread) ----------------------------------------------------------------------------
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// same, descriptors given as (lo, hi) 32-bit halves: lo = (smem address >> 4) | (LBO >> 4) << 16, hi = (SBO >> 4) | 1 << 14
// (bit 46 = descriptor version 1) — one integer add per descriptor on the issuing thread
__device__ __forceinline__ void mma_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, {%7, %7, %7, %7}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// same with every descriptor field an immediate: lo = sbd + A_LO where sbd = (shared-memory base) >> 4 is the only
// register input besides the TMEM base — nothing for the compiler to hoist out of the issue loop and spill
template <uint32_t A_LO, uint32_t A_HI, uint32_t B_LO, uint32_t B_HI, uint32_t IDESC, uint32_t DCOL>
__device__ __forceinline__ void mma_imm(uint32_t tmem_base, uint32_t sbd, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 al, ah, bl, bh, dd, id;\n\t"
      ".reg .b64 da, db;\n\t"
      "add.u32 al, %1, %3;\n\t"
      "mov.u32 ah, %4;\n\t"
      "add.u32 bl, %1, %5;\n\t"
      "mov.u32 bh, %6;\n\t"
      "add.u32 dd, %0, %8;\n\t"
      "mov.u32 id, %7;\n\t"
      "mov.b64 da, {al, ah};\n\t"
      "mov.b64 db, {bl, bh};\n\t"
      "setp.ne.b32 p, %2, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [dd], da, db, id, p;\n\t"
      "}\n" ::"r"(tmem_base),
      "r"(sbd), "r"(accumulate), "n"(A_LO), "n"(A_HI), "n"(B_LO), "n"(B_HI), "n"(IDESC), "n"(DCOL)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// A operand in TENSOR MEMORY (row m of the M = 128 tile in lane m; 32-bit column c holds k = 2c in the low half and
// k = 2c + 1 in the high half; one K = 16 step = 8 columns), B from shared memory; every field but the two bases an
// immediate (see mma_imm).  tests/cuda/ts_probe.cu checks the layout against a CPU GEMM and measures the issue rate.
template <uint32_t ACOL, uint32_t B_LO, uint32_t B_HI, uint32_t IDESC, uint32_t DCOL>
__device__ __forceinline__ void mma_ts_imm(uint32_t tmem_base, uint32_t sbd, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 bl, bh, dd, aa, id;\n\t"
      ".reg .b64 db;\n\t"
      "add.u32 bl, %1, %4;\n\t"
      "mov.u32 bh, %5;\n\t"
      "add.u32 dd, %0, %7;\n\t"
      "add.u32 aa, %0, %3;\n\t"
      "mov.u32 id, %6;\n\t"
      "mov.b64 db, {bl, bh};\n\t"
      "setp.ne.b32 p, %2, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [dd], [aa], db, id, p;\n\t"
      "}\n" ::"r"(tmem_base),
      "r"(sbd), "r"(accumulate), "n"(ACOL), "n"(B_LO), "n"(B_HI), "n"(IDESC), "n"(DCOL)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM <-> registers: 32 lanes x 32 bit, N consecutive columns per thread ------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// ---- BF16x3 split ----------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// pack two consecutive-K values: low half = even k
__device__ __forceinline__ void split_pack2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat16 h0, l0, h1, l1;
  split_bf16(x0, h0, l0);
  split_bf16(x1, h1, l1);
  hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

// byte offset of element (r, c) of a logical [R][C] bf16 matrix in the core-matrix layout
__device__ __forceinline__ uint32_t core_off(int r, int c, int C) {
  return (uint32_t)(((r >> 3) * (C >> 3) + (c >> 3)) * 128 + (r & 7) * 16 + (c & 7) * 2);
}

}  // namespace umma
}  // namespace hjb

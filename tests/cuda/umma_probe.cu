// Standalone probe for the tcgen05 building blocks of csrc/umma.cuh (run on a B200 through gpurun):
//   * checks every shared-memory operand layout / major-ness / format combination the vhjb tensor-core kernel relies
//     on against a CPU reference (D = A B^T, fp32 accumulate),
//   * including the "M = 128 over fewer valid rows" trick (fp16 x bf16 MIXED operands were tried here and raise
//     'illegal instruction' on B200: both operands of a kind::f16 MMA must have the same format),
//   * and measures the issue/throughput constants the kernel schedule is designed around.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/build/umma_probe tests/cuda/umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>

#include "../../q_learning_with_hjb_b200/csrc/umma.cuh"

using namespace hjb::umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct ProbeArgs {
  const uint8_t* a_img; const uint8_t* b_img;
  int a_bytes, b_bytes;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, idesc;
  int ksteps; uint32_t a_kadv, b_kadv;
  int ncols;
  float* out;  // [128][ncols]
};

__global__ void __launch_bounds__(128, 1) gemm_probe(ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((p.a_bytes + 1023) / 1024) * 1024;
  for (int i = tid * 16; i < p.a_bytes; i += 128 * 16) *(uint4*)(sa + i) = *(const uint4*)(p.a_img + i);
  for (int i = tid * 16; i < p.b_bytes; i += 128 * 16) *(uint4*)(sb + i) = *(const uint4*)(p.b_img + i);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    for (int k = 0; k < p.ksteps; ++k) {
      const uint64_t ad = smem_desc(smem_u32(sa) + k * p.a_kadv, p.a_lbo, p.a_sbo);
      const uint64_t bd = smem_desc(smem_u32(sb) + k * p.b_kadv, p.b_lbo, p.b_sbo);
      mma_ss(tm, ad, bd, p.idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < p.ncols; c0 += 8) {
    uint32_t v[8];
    tmem_ld8(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc_wait_ld();
    for (int j = 0; j < 8; ++j) p.out[tid * p.ncols + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

// ---- host side -------------------------------------------------------------------------------------------
static uint16_t to_fmt(float x, int fmt) {
  if (fmt == kBF16) { __nv_bfloat16 h = __float2bfloat16_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }
  __half h = __float2half_rn(x); uint16_t u; memcpy(&u, &h, 2); return u;
}
static float from_fmt(uint16_t u, int fmt) {
  if (fmt == kBF16) { __nv_bfloat16 h; memcpy(&h, &u, 2); return __bfloat162float(h); }
  __half h; memcpy(&h, &u, 2); return __half2float(h);
}
// X[R][C] (C contiguous) -> core-matrix image; element (r, c) at ((r/8)(C/8) + c/8) 128 + (r%8) 16 + (c%8) 2
static std::vector<uint8_t> pack(const std::vector<float>& X, int R, int C, int fmt, int pad_bytes) {
  std::vector<uint8_t> img((size_t)R * C * 2 + pad_bytes, 0);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) {
      size_t off = ((size_t)(r / 8) * (C / 8) + c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2;
      uint16_t u = to_fmt(X[(size_t)r * C + c], fmt);
      memcpy(&img[off], &u, 2);
    }
  // padding: finite garbage
  for (size_t i = (size_t)R * C * 2; i + 1 < img.size(); i += 2) { uint16_t u = to_fmt(0.25f, fmt); memcpy(&img[i], &u, 2); }
  return img;
}

struct Case { const char* name; int Mvalid, N, K; int afmt, bfmt; int a_mn, b_mn; int mma_m = 128; };

static bool run_case(const Case& c) {
  const int M = c.Mvalid, N = c.N, K = c.K;
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(1234 + M * 7 + N * 3 + K + c.afmt * 11 + c.bfmt * 13 + c.a_mn * 17 + c.b_mn * 19);
  auto rnd = []() { return (float)((rand() % 2001) - 1000) / 500.0f; };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  for (auto& v : A) v = from_fmt(to_fmt(v, c.afmt), c.afmt);
  for (auto& v : B) v = from_fmt(to_fmt(v, c.bfmt), c.bfmt);
  ProbeArgs p{};
  std::vector<uint8_t> ai, bi;
  const int pad = 4096;
  if (!c.a_mn) {  // K-major: X[R = M][C = K]
    ai = pack(A, M, K, c.afmt, pad + (128 - M) * K * 2);
    p.a_sbo = (K / 8) * 128; p.a_lbo = 128; p.a_kadv = 256;
  } else {        // MN-major: X[R = K][C = M]
    std::vector<float> T((size_t)K * M);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) T[(size_t)k * M + m] = A[(size_t)m * K + k];
    ai = pack(T, K, M, c.afmt, pad);
    p.a_sbo = 128; p.a_lbo = (M / 8) * 128; p.a_kadv = 2 * p.a_lbo;
  }
  if (!c.b_mn) {
    bi = pack(B, N, K, c.bfmt, pad);
    p.b_sbo = (K / 8) * 128; p.b_lbo = 128; p.b_kadv = 256;
  } else {
    std::vector<float> T((size_t)K * N);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) T[(size_t)k * N + n] = B[(size_t)n * K + k];
    bi = pack(T, K, N, c.bfmt, pad);
    p.b_sbo = 128; p.b_lbo = (N / 8) * 128; p.b_kadv = 2 * p.b_lbo;
  }
  p.idesc = idesc_f16(c.mma_m, N, c.afmt, c.bfmt, c.a_mn, c.b_mn);
  p.ksteps = K / 16;
  p.ncols = N;
  p.a_bytes = (int)ai.size(); p.b_bytes = (int)bi.size();
  uint8_t *da, *db; float* dout;
  CK(cudaMalloc(&da, ai.size())); CK(cudaMalloc(&db, bi.size())); CK(cudaMalloc(&dout, 128 * N * 4));
  CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, 128 * N * 4));
  p.a_img = da; p.b_img = db; p.out = dout;
  size_t smem = ((ai.size() + 1023) / 1024) * 1024 + bi.size() + 1024;
  CK(cudaFuncSetAttribute(gemm_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gemm_probe<<<1, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s LAUNCH FAILED: %s\n", c.name, cudaGetErrorString(e)); exit(3); }
  std::vector<float> out((size_t)128 * N);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
      // accumulator row -> TMEM lane: M = 128: lane = row; M = 64: 16 rows per 32-lane quarter (cute tmem_frg_1sm)
      const int lane_of_m = c.mma_m == 64 ? (m % 16) + 32 * (m / 16) : m;
      double d = fabs(s - (double)out[(size_t)lane_of_m * N + n]);
      if (!(d <= maxerr)) maxerr = d;   // NaN-propagating
      if (fabs(s) > maxref) maxref = fabs(s);
    }
  bool ok = maxerr <= 1e-3 * (maxref + 1);
  printf("%-44s M=%3d N=%3d K=%3d  maxerr %.3e (ref max %.2f)  %s\n", c.name, M, N, K, maxerr, maxref, ok ? "OK" : "FAIL");
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return ok;
}

// ---- timing probes -----------------------------------------------------------------------------------------
// one thread issues `reps` MMAs (M=128, N, K=16) on resident operands; cycles from first issue to commit-arrival
__global__ void __launch_bounds__(128, 1) mma_rate_probe(int N, int reps, int a_mn, int b_mn, long long* cycles, int mma_m) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = idesc_f16(mma_m, N, kF16, kF16, a_mn, b_mn);
    const uint32_t a_lbo = a_mn ? 2048 : 128, a_sbo = a_mn ? 128 : 2048;
    const uint32_t b_lbo = b_mn ? (N / 8) * 128 : 128, b_sbo = b_mn ? 128 : 2048;
    const uint32_t a_kadv = a_mn ? 2 * a_lbo : 256, b_kadv = b_mn ? 2 * b_lbo : 256;
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 32768;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int k = r & 7;
      mma_ss(tm + (uint32_t)((r >> 3) & 1) * 256, smem_desc(sa + k * a_kadv, a_lbo, a_sbo), smem_desc(sb + k * b_kadv, b_lbo, b_sbo),
             idesc, r >= 16);
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// nwarps warps each issue `reps` tcgen05.ld 32x32b.x32 (+ wait) back to back
__global__ void __launch_bounds__(256, 1) ldtm_rate_probe(int reps, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint32_t v[32];
    tmem_ld32(tm + ((r * 32 + (warp >> 2) * 64) & 255), v);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; j += 8) acc += __uint_as_float(v[j]);
  }
  long long t1 = clock64();
  __syncthreads();
  if (tid == 0) cycles[0] = t1 - t0;
  sink[tid] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  std::vector<Case> cases = {
    {"bf16 A:K  B:K", 128, 64, 64, kBF16, kBF16, 0, 0},
    {"bf16 A:MN B:K", 128, 64, 64, kBF16, kBF16, 1, 0},
    {"bf16 A:K  B:MN", 128, 64, 64, kBF16, kBF16, 0, 1},
    {"bf16 A:MN B:MN", 128, 64, 64, kBF16, kBF16, 1, 1},
    {"fp16 A:K  B:K", 128, 64, 64, kF16, kF16, 0, 0},
    {"fp16 A:MN B:MN N=32", 128, 32, 128, kF16, kF16, 1, 1},
    {"fp16 A:K  B:MN N=32", 128, 32, 128, kF16, kF16, 0, 1},
    {"fp16 A:MN B:K  N=128 (wgrad-like)", 128, 128, 32, kF16, kF16, 1, 0},
    {"fp16 A:K  B:K  N=128 K=32 (wgrad)", 128, 128, 32, kF16, kF16, 0, 0},
    {"partial-M 64 valid, A:MN B:MN N=64", 64, 64, 128, kF16, kF16, 1, 1},
    {"partial-M 32 valid, A:MN B:MN N=64", 32, 64, 128, kF16, kF16, 1, 1},
    {"partial-M 32 valid, A:MN B:K  N=16", 32, 16, 128, kF16, kF16, 1, 0},
    {"partial-M 64 valid, A:K  B:K  N=64", 64, 64, 128, kF16, kF16, 0, 0},
    {"K=16 A:MN B:K N=32 (W1 fwd)", 128, 32, 16, kF16, kF16, 1, 0},
    {"M=64 MMA, A:MN B:MN N=64 (y = h2 W3)", 64, 64, 128, kF16, kF16, 1, 1, 64},
    {"M=64 MMA, A:MN B:K  N=16 (g0 = g1 W1^T)", 64, 16, 128, kF16, kF16, 1, 0, 64},
    {"M=64 MMA, A:K  B:K  N=64", 64, 64, 64, kF16, kF16, 0, 0, 64},
    {"K=32 A:K  B:MN N=16 (W1 grad)", 128, 16, 32, kBF16, kBF16, 0, 1},
  };
  int fails = 0;
  for (auto& c : cases) fails += !run_case(c);
  printf("layout cases failed: %d of %zu\n", fails, cases.size());

  long long* dc; float* sink; CK(cudaMalloc(&dc, 8)); CK(cudaMalloc(&sink, 1024));
  CK(cudaFuncSetAttribute(mma_rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
  for (int amn = 0; amn < 2; ++amn)
    for (int bmn = 0; bmn < 2; ++bmn)
      for (int N : {16, 32, 64, 128, 256}) {
        if (N == 256 && bmn) continue;
        long long best = 1ll << 60;
        for (int it = 0; it < 3; ++it) {
          mma_rate_probe<<<1, 128, 132 * 1024>>>(N, 512, amn, bmn, dc, 128);
          CK(cudaDeviceSynchronize());
          long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
          if (c < best) best = c;
        }
        printf("mma rate: A:%s B:%s N=%3d : %.1f cycles/MMA (512 MMAs, K=16, M=128)\n", amn ? "MN" : "K ", bmn ? "MN" : "K ", N, best / 512.0);
      }
  for (int threads : {128, 256}) {
    long long best = 1ll << 60;
    for (int it = 0; it < 3; ++it) {
      ldtm_rate_probe<<<1, threads>>>(256, dc, sink);
      CK(cudaDeviceSynchronize());
      long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
      if (c < best) best = c;
    }
    printf("tcgen05.ld 32x32b.x32: %d warps, %.1f cycles per load+wait per warp (4 KB each)\n", threads / 32, best / 256.0);
  }
  return fails ? 1 : 0;
}

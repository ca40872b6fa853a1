// Probe of tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form), the building block of the states-on-M residual
// kernel: D[s][j] (+)= sum_k A[s][k] B[k][j], A = activations of 128 states (TMEM: row s in lane s, 32-bit column c holds
// k = 2c in the low half and k = 2c + 1 in the high half, written by the lane's own thread with tcgen05.st), B = a resident
// weight matrix in shared memory (core-matrix layout, either major-ness), fp32 accumulators in other TMEM columns.
//   1. correctness against a CPU GEMM for N in {128, 64, 16}, K = 128, B K-major and MN-major;
//   2. cycles per MMA (K = 16) back to back, for N in {16, 32, 64, 128}: is the M = 128 / N = 64 shape math-bound (32
//      cycles) once the 4 KB A tile no longer comes from shared memory?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/build/ts_probe tests/cuda/ts_probe.cu
#include <cuda_fp16.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../q_learning_with_hjb_b200/csrc/umma.cuh"
using namespace hjb::umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

struct TsArgs {
  const float* A;        // [128][K] row-major fp32 (values exactly representable in fp16)
  const uint8_t* b_img;  // core-matrix image of B
  int b_bytes;
  uint32_t b_lbo, b_sbo, b_kadv, idesc;
  int K, N;
  float* out;            // [128][N]
};

__global__ void __launch_bounds__(128, 1) ts_gemm(TsArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid * 16; i < p.b_bytes; i += 128 * 16) *(uint4*)(smem + i) = *(const uint4*)(p.b_img + i);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t cA = 256;   // A operand columns [256, 256 + K/2)
  // this thread's row -> TMEM, 8 packed columns (16 k values) at a time
  for (int c0 = 0; c0 < p.K / 2; c0 += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) {
      const __half2 h = __floats2half2_rn(p.A[tid * p.K + 2 * (c0 + j)], p.A[tid * p.K + 2 * (c0 + j) + 1]);
      v[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + cA + c0),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    for (int k = 0; k < p.K / 16; ++k) {
      const uint64_t bd = smem_desc(smem_u32(smem) + k * p.b_kadv, p.b_lbo, p.b_sbo);
      mma_ts(tm, tm + cA + 8 * k, bd, p.idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < p.N; c0 += 8) {
    uint32_t v[8];
    tmem_ld8(lane_base + c0, v);
    tc_wait_ld();
    for (int j = 0; j < 8; ++j) p.out[tid * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int TS>
__global__ void __launch_bounds__(128, 1) ts_rate(int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 96 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base;
  {  // A operand columns [384, 448): ones
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = 0x3c003c00u;
    tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + 384, v);
    tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + 416, v);
    tc_wait_st();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 1) {
    constexpr uint32_t idesc = idesc_f16(128, N, kF16, kF16, 0, 1);
    constexpr uint32_t b_lbo = (N / 8) * 128, b_sbo = 128, b_kadv = 2 * b_lbo;   // B MN-major: W[k][j]
    const uint64_t ad0 = smem_desc(smem_u32(smem) + 65536, 128, 2048), bd0 = smem_desc(smem_u32(smem), b_lbo, b_sbo);
    long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (TS) mma_ts(tm + (r & 1) * 128, tm + 384 + 8 * k, bd0 + ((k * b_kadv) >> 4), idesc, (r > 1 || k > 0) ? 1u : 0u);
          else mma_ss(tm + (r & 1) * 128, ad0 + ((k * 256) >> 4), bd0 + ((k * b_kadv) >> 4), idesc, (r > 1 || k > 0) ? 1u : 0u);
        }
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if ((tid & 31) == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static uint16_t h16(float x) { __half h = __float2half_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }
static float f16(float x) { __half h = __float2half_rn(x); return __half2float(h); }
static std::vector<uint8_t> pack(const std::vector<float>& X, int R, int C) {
  std::vector<uint8_t> img((size_t)R * C * 2 + 4096, 0);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) {
      size_t off = ((size_t)(r / 8) * (C / 8) + c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2;
      uint16_t u = h16(X[(size_t)r * C + c]);
      memcpy(&img[off], &u, 2);
    }
  return img;
}

static bool run_case(int N, int K, int b_mn) {
  std::vector<float> A((size_t)128 * K), B((size_t)N * K);   // D[m][n] = sum_k A[m][k] B[n][k]
  srand(77 + N + 3 * K + b_mn);
  for (auto& v : A) v = f16((float)((rand() % 2001) - 1000) / 500.0f);
  for (auto& v : B) v = f16((float)((rand() % 2001) - 1000) / 500.0f);
  TsArgs p{};
  std::vector<uint8_t> bi;
  if (!b_mn) {   // stored [n][k]: K-major
    bi = pack(B, N, K);
    p.b_sbo = (K / 8) * 128; p.b_lbo = 128; p.b_kadv = 256;
  } else {       // stored [k][n]: MN-major (a Flax (in, out) kernel used as is)
    std::vector<float> T((size_t)K * N);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) T[(size_t)k * N + n] = B[(size_t)n * K + k];
    bi = pack(T, K, N);
    p.b_sbo = 128; p.b_lbo = (N / 8) * 128; p.b_kadv = 2 * p.b_lbo;
  }
  p.idesc = idesc_f16(128, N, kF16, kF16, 0, b_mn);
  p.K = K; p.N = N; p.b_bytes = (int)bi.size();
  float *dA, *dO; uint8_t* dB;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dO, 128 * N * 4)); CK(cudaMalloc(&dB, bi.size()));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, bi.data(), bi.size(), cudaMemcpyHostToDevice));
  p.A = dA; p.b_img = dB; p.out = dO;
  CK(cudaFuncSetAttribute(ts_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  ts_gemm<<<1, 128, 64 * 1024>>>(p);
  CK(cudaDeviceSynchronize());
  std::vector<float> O((size_t)128 * N);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      maxerr = fmax(maxerr, fabs(r - O[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(r));
    }
  const bool ok = maxerr < 1e-4 * maxref;
  printf("TS  A:TMEM B:%s  M=128 N=%3d K=%3d  maxerr %.3e (ref max %.2f)  %s\n", b_mn ? "MN" : "K ", N, K, maxerr, maxref, ok ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dO); cudaFree(dB);
  return ok;
}

template <int N, int TS> void rate(long long* dc) {
  CK(cudaFuncSetAttribute(ts_rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  long long best = 1ll << 60;
  for (int it = 0; it < 3; ++it) {
    ts_rate<N, TS><<<1, 128, 100 * 1024>>>(64, dc);
    CK(cudaDeviceSynchronize());
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    if (c < best) best = c;
  }
  printf("rate M=128 A:%s B:MN N=%3d: %.1f cycles/MMA (K = 16)\n", TS ? "TMEM" : "smem", N, best / 512.0);
}

int main() {
  int bad = 0;
  for (int b_mn = 0; b_mn < 2; ++b_mn)
    for (int N : {128, 64, 16}) bad += !run_case(N, 128, b_mn);
  bad += !run_case(128, 16, 1);
  bad += !run_case(128, 64, 0);
  printf("TS layout cases failed: %d\n", bad);
  long long* dc; CK(cudaMalloc(&dc, 8));
  rate<16, 1>(dc); rate<32, 1>(dc); rate<64, 1>(dc); rate<128, 1>(dc);
  rate<16, 0>(dc); rate<64, 0>(dc); rate<128, 0>(dc);
  return bad != 0;
}

#include <cstdio>
// Issue-rate probe: cycles per tcgen05.mma (SS mode, K = 16, fp16) for M in {64, 128} and N in {16..128}, issued back to
// back by one elected lane with precomputed descriptors.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -std=c++17 -o tests/cuda/build/umma_rate tests/cuda/umma_rate.cu
#include "../../q_learning_with_hjb_b200/csrc/umma.cuh"
using namespace hjb::umma;
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <int M, int N, int AMN, int BMN>
__global__ void __launch_bounds__(128, 1) rate2(int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base;
  if (warp == 1) {
    constexpr uint32_t idesc = idesc_f16(M, N, kF16, kF16, AMN, BMN);
    constexpr uint32_t a_lbo = AMN ? 2048 : 128, a_sbo = AMN ? 128 : 2048;
    constexpr uint32_t b_lbo = BMN ? (N / 8) * 128 : 128, b_sbo = BMN ? 128 : 2048;
    constexpr uint32_t a_kadv = AMN ? 2 * a_lbo : 256, b_kadv = BMN ? 2 * b_lbo : 256;
    const uint64_t ad0 = smem_desc(smem_u32(smem), a_lbo, a_sbo), bd0 = smem_desc(smem_u32(smem) + 32768, b_lbo, b_sbo);
    long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          mma_ss(tm + (r & 1) * 256, ad0 + ((k * a_kadv) >> 4), bd0 + ((k * b_kadv) >> 4), idesc, (r > 1 || k > 0) ? 1u : 0u);
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
template <int M, int N, int AMN, int BMN> void go(long long* dc) {
  cudaFuncSetAttribute(rate2<M, N, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024);
  long long best = 1ll << 60;
  for (int it = 0; it < 3; ++it) {
    rate2<M, N, AMN, BMN><<<1, 128, 132 * 1024>>>(64, dc);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("fail\n"); exit(1); }
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    if (c < best) best = c;
  }
  printf("M=%3d A:%s B:%s N=%3d: %.1f cycles/MMA\n", M, AMN ? "MN" : "K ", BMN ? "MN" : "K ", N, best / 512.0);
}
int main() {
  long long* dc; cudaMalloc(&dc, 8);
  go<128, 16, 0, 0>(dc); go<128, 32, 0, 0>(dc); go<128, 64, 0, 0>(dc); go<128, 128, 0, 0>(dc);
  go<128, 16, 1, 1>(dc); go<128, 32, 1, 1>(dc); go<128, 64, 1, 1>(dc); go<128, 128, 1, 1>(dc);
  go<128, 32, 1, 0>(dc); go<128, 64, 0, 1>(dc);
  go<64, 16, 1, 0>(dc); go<64, 32, 1, 1>(dc); go<64, 64, 1, 1>(dc); go<64, 128, 1, 1>(dc); go<64, 64, 0, 0>(dc); go<64, 16, 0, 1>(dc);
  return 0;
}

// Cost probe for the element-wise passes of vhjb_tc.cuh: cycles per pass of one SM for W warps, by ingredient.
//   mode 0: TMEM loads only (2 x 32 columns per thread) + a trivial reduction
//   mode 1: + relu mask + hi piece (one pack per 2 elements) stored to shared memory
//   mode 2: + lo piece (unpack, subtract, pack) — the full store8 of the kernels
//   mode 3: full, but ONE TMEM load (mask from a register bit pattern)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tests/cuda/build/pass_probe tests/cuda/pass_probe.cu
#include <cstdio>

#include "../../q_learning_with_hjb_b200/csrc/vhjb_tc.cuh"
using namespace hjb;
using namespace hjb::tc;
using namespace hjb::umma;

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(int reps, int active_warps, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base;
  const int grp = warp >> 3, wg = warp & 7, q = wg & 3, hh = wg >> 2;
  const int j = 32 * q + lane, sc0 = 32 * hh;
  const uint32_t tl = tm + ((uint32_t)(32 * q) << 16) + (uint32_t)grp * 256;
  const uint32_t buf = (uint32_t)grp * 2 * kFPiece;
  float acc = 0.f;
  uint32_t bits = 0x5a5a5a5au ^ tid;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < active_warps) {
    for (int r = 0; r < reps; ++r) {
      uint32_t d[32], st[32];
      tmem_ld32(tl + sc0, d);
      if (MODE != 3) tmem_ld32(tl + 64 + sc0, st);
      tc_wait_ld();
      if (MODE == 0) {
#pragma unroll
        for (int t = 0; t < 32; ++t) acc += __uint_as_float(d[t]) + __uint_as_float(st[t]);
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float o[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const bool on = MODE == 3 ? ((bits >> (8 * g + t)) & 1u) != 0 : __uint_as_float(st[8 * g + t]) > 0.f;
            o[t] = on ? __uint_as_float(d[8 * g + t]) : 0.f;
          }
          if (MODE == 1) {
            uint4 hi;
            hi.x = __float_as_uint(0.f);
            const __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]), h2 = __floats2half2_rn(o[4], o[5]),
                          h3 = __floats2half2_rn(o[6], o[7]);
            hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
            hi.z = *reinterpret_cast<const uint32_t*>(&h2); hi.w = *reinterpret_cast<const uint32_t*>(&h3);
            const uint32_t off = buf + (uint32_t)(j >> 3) * kRbF + ((uint32_t)((sc0 + 8 * g) >> 3) << 7) + ((uint32_t)(j & 7) << 4);
            *reinterpret_cast<uint4*>(smem + off) = hi;
          } else {
            store8<kF16>(smem, buf, kFPiece, kRbF, j, sc0 + 8 * g, o);
          }
        }
        bits = bits * 1664525u + 1013904223u;
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (tid == 0) cycles[0] = t1 - t0;
  if (acc == 123.456f) sink[tid] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int MODE>
void run(const char* name, long long* dc, float* sink) {
  const int reps = 2000;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int w : {2, 4, 8, 16}) {
    probe<MODE><<<1, 512, 200 * 1024>>>(reps, w, dc, sink);
    long long c = 0;
    cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-34s warps=%2d: %8.1f cycles per pass (%s)\n", name, w, (double)c / reps, cudaGetErrorString(e));
  }
}

int main() {
  long long* dc; float* sink;
  cudaMalloc(&dc, 64); cudaMalloc(&sink, 4096);
  run<0>("tmem loads only (2 x 32 cols)", dc, sink);
  run<1>("+ mask + hi piece + STS", dc, sink);
  run<2>("+ lo piece (full store8)", dc, sink);
  run<3>("full, one tmem load (mask bits)", dc, sink);
  return 0;
}

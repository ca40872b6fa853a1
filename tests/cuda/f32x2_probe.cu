// Does the packed fp32 FMA of sm_100 (fma.rn.f32x2 -> FFMA2: two FMAs per lane and instruction) buy anything for an
// ISSUE-bound kernel?  Four variants of the same arithmetic (independent FMA chains, 16 warps / SM sub-partition... and a
// mix with ALU-pipe instructions like the rollout loop's): scalar FFMA, packed FFMA2, each with and without interleaved
// integer / min-max work.  Prints G FMA/s per variant.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/build/f32x2_probe tests/cuda/f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

template <int MODE>   // 0 scalar, 1 packed, 2 scalar + alu mix, 3 packed + alu mix
__global__ void __launch_bounds__(256) probe(int iters, float* sink) {
  float2 x[8];
  const float s = 1.0f + 1e-7f * threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(0.1f * i + 1e-3f * threadIdx.x, 0.2f * i);
  const float2 a = make_float2(s, s * 0.999f), b = make_float2(1e-3f, 2e-3f);
  float m = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2) {
        x[i].x = fmaf(x[i].x, a.x, b.x);
        x[i].y = fmaf(x[i].y, a.y, b.y);
      } else {
        x[i] = fma2(x[i], a, b);
      }
      if (MODE >= 2 && (i & 1) == 0) m = fminf(fmaxf(m, x[i].x), 3.0f);   // one ALU-pipe pair per 4 FMAs, as in the rollout loop
    }
  }
  float acc = m;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += x[i].x + x[i].y;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, float* sink) {
  const int iters = 20000, grid = 148 * 8;
  probe<MODE><<<grid, 256>>>(100, sink);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    probe<MODE><<<grid, 256>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const double fmas = (double)iters * 16 * grid * 256;
  printf("%-28s %.3f ms  %.1f TFLOP/s (2 flops per FMA)\n", name, best, 2 * fmas / best * 1e-9);
}

int main() {
  float* sink;
  cudaMalloc(&sink, 148 * 8 * 256 * 4);
  run<0>("scalar FFMA", sink);
  run<1>("packed FFMA2", sink);
  run<2>("scalar FFMA + ALU mix", sink);
  run<3>("packed FFMA2 + ALU mix", sink);
  return cudaDeviceSynchronize() != cudaSuccess;
}

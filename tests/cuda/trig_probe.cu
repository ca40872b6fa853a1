// Accuracy + cost probe for the sin/cos candidates of the rollout kernels' fast path, on the wrapped range [-pi, pi)
// (and [0, 2 pi) for goals at pi).  Variants:
//   0  __sinf / __cosf                      (FMUL + MUFU.SIN, MUFU.COS)
//   1  MUFU + first-order correction of the argument's rounding  (x / 2 pi is rounded to fp32 before the MUFU)
//   2  quadrant reduction + minimax polynomials (no MUFU)
//   3  libdevice sincosf
//   4  shared-memory table (2^-7 spacing, built in double) + second-order Taylor step — what the rollout kernels use
// Prints max |err| of sin and cos against fp64 and ns per call in a dependent loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/build/trig_probe tests/cuda/trig_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../q_learning_with_hjb_b200/csrc/hjb_common.cuh"
using namespace hjb;

// Candidate 1 (rejected): MUFU.SIN / MUFU.COS take the angle in revolutions; sin.approx multiplies by fp32(1 / 2 pi)
// (FMUL.RZ) first.  The rounding error of that product is known exactly — e = fma(x, c_hi, -y) plus x * c_lo — so a
// first-order rotation by 2 pi (e + x c_lo) removes it.  Measured: 3.6e-7 remains — the error is the interpolator's own.
__device__ __forceinline__ void sincos_mufu_corrected(float x, float& s, float& c) {
  constexpr float kHi = 0.15915493667125702f;       // fp32(1 / 2 pi)
  constexpr float kLo2Pi = 4.0342060396188925e-08f; // 2 pi (1 / 2 pi - kHi)
  const float y = __fmul_rz(x, kHi);
  const float e = __fmaf_rn(x, kHi, -y);
  const float d = __fmaf_rn(x, kLo2Pi, e * 6.28318530717958648f);
  const float s0 = __sinf(x), c0 = __cosf(x);
  s = __fmaf_rn(d, c0, s0);
  c = __fmaf_rn(-d, s0, c0);
}

__shared__ float2 g_tab[kTrigSize];
__device__ void build_tab() {
  for (int i = threadIdx.x; i < kTrigSize; i += blockDim.x) {
    double sv, cv;
    sincos((double)(i - kTrigHalf) * (1.0 / (double)(1 << kTrigLog2)), &sv, &cv);
    g_tab[i] = make_float2((float)sv, (float)cv);
  }
  __syncthreads();
}

template <int V>
__device__ __forceinline__ void sc(float x, float& s, float& c) {
  if constexpr (V == 4) sincos_tab(g_tab, x, s, c);
  else if constexpr (V == 0) { s = __sinf(x); c = __cosf(x); }
  else if constexpr (V == 1) sincos_mufu_corrected(x, s, c);
  else if constexpr (V == 2) sincos_poly(x, s, c);
  else sincosf(x, &s, &c);
}

template <int V>
__global__ void err_kernel(float lo, float hi, long long npts, double* out) {
  if (V == 4) build_tab();
  double es = 0, ec = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += stride) {
    const float x = lo + (hi - lo) * (float)((double)i / (double)npts);
    float s, c;
    sc<V>(x, s, c);
    es = fmax(es, fabs((double)s - sin((double)x)));
    ec = fmax(ec, fabs((double)c - cos((double)x)));
  }
  for (int o = 16; o; o >>= 1) {
    es = fmax(es, __shfl_xor_sync(~0u, es, o));
    ec = fmax(ec, __shfl_xor_sync(~0u, ec, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(es));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(ec));
  }
}

template <int V>
__global__ void time_kernel(int reps, float* sink) {
  if (V == 4) build_tab();
  float x = 0.001f * threadIdx.x, acc = 0.f;
  for (int r = 0; r < reps; ++r) {
    float s, c;
    sc<V>(x, s, c);
    acc = fmaf(s, 0.5f, acc) + c;
    x = wrap_pi_<true>(x + 0.37f + 1e-3f * s);
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int V>
void run(const char* name) {
  double* d;
  cudaMalloc(&d, 16);
  const float ranges[3][2] = {{-3.14159274f, 3.14159274f}, {0.f, 6.2831855f}, {-0.01f, 0.01f}};
  for (auto& r : ranges) {
    if (V == 4 && r[1] > kTrigRange) continue;      // the table covers the wrapped range only
    cudaMemset(d, 0, 16);
    err_kernel<V><<<148 * 8, 256>>>(r[0], r[1], 1LL << 28, d);
    double h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-16s [%9.5f, %9.5f]  max|err| sin %.3e  cos %.3e\n", name, r[0], r[1], h[0], h[1]);
  }
  float* sink;
  cudaMalloc(&sink, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  time_kernel<V><<<148 * 8, 256>>>(1000, sink);
  cudaEventRecord(e0);
  time_kernel<V><<<148 * 8, 256>>>(20000, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-16s %.3f ms for 20000 iterations x %d threads: %.2f G sincos/s\n", name, ms, 148 * 8 * 256,
         20000.0 * 148 * 8 * 256 / ms * 1e-6);
  cudaFree(d); cudaFree(sink);
}

int main() {
  run<0>("mufu");
  run<1>("mufu+corr");
  run<2>("poly");
  run<3>("libdevice");
  run<4>("table+taylor2");
  return cudaDeviceSynchronize() != cudaSuccess;
}

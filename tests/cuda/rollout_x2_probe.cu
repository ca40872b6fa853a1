// Is fma.rn.f32x2 (FFMA2: two fp32 FMAs per lane and instruction) worth a two-environments-per-thread rollout loop?
// The C4 step (2-D quadrotor, hover LQR with clip, forward Euler, unit cost, table + Taylor sin/cos) written once over a
// generic scalar type T: T = float is the production arithmetic (one environment per thread), T = f2 carries two
// environments per thread through packed add / mul / fma; the table look-ups, the clamps and the index arithmetic stay
// scalar.  Both variants produce bit-identical states (packed ops are IEEE fma.rn / add.rn / mul.rn per half).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/build/rollout_x2_probe tests/cuda/rollout_x2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

struct f2 { float a, b; };

__device__ __forceinline__ uint64_t pk(f2 v) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.a), "f"(v.b)); return r; }
__device__ __forceinline__ f2 upk(uint64_t r) { f2 v; asm("mov.b64 {%0, %1}, %2;" : "=f"(v.a), "=f"(v.b) : "l"(r)); return v; }
__device__ __forceinline__ f2 fma_(f2 x, f2 y, f2 z) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(x)), "l"(pk(y)), "l"(pk(z))); return upk(d); }
__device__ __forceinline__ f2 mul_(f2 x, f2 y) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(x)), "l"(pk(y))); return upk(d); }
__device__ __forceinline__ f2 add_(f2 x, f2 y) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(x)), "l"(pk(y))); return upk(d); }
__device__ __forceinline__ float fma_(float x, float y, float z) { return __fmaf_rn(x, y, z); }
__device__ __forceinline__ float mul_(float x, float y) { return __fmul_rn(x, y); }
__device__ __forceinline__ float add_(float x, float y) { return __fadd_rn(x, y); }
template <class T> __device__ __forceinline__ T bc(float c);
template <> __device__ __forceinline__ float bc<float>(float c) { return c; }
template <> __device__ __forceinline__ f2 bc<f2>(float c) { return f2{c, c}; }
__device__ __forceinline__ float neg_(float x) { return -x; }
__device__ __forceinline__ f2 neg_(f2 x) { return f2{-x.a, -x.b}; }

__device__ __forceinline__ float clampf(float v, float lo, float hi) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "f"(lo));
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 clampf(f2 v, float lo, float hi) { return f2{clampf(v.a, lo, hi), clampf(v.b, lo, hi)}; }

constexpr int kLog2 = 5, kHalf = 8 << kLog2, kSize = 2 * kHalf + 1;
__device__ __forceinline__ unsigned tab_index(float t) { return min((unsigned)(__float_as_int(t) - (0x4B400000 - kHalf)), (unsigned)(kSize - 1)); }

template <class T> struct Tab;
template <> struct Tab<float> {
  static __device__ __forceinline__ void load(const float2* tab, float t, float& S, float& C) { const float2 e = tab[tab_index(t)]; S = e.x; C = e.y; }
};
template <> struct Tab<f2> {
  static __device__ __forceinline__ void load(const float2* tab, f2 t, f2& S, f2& C) {
    const float2 e0 = tab[tab_index(t.a)], e1 = tab[tab_index(t.b)];
    S = f2{e0.x, e1.x}; C = f2{e0.y, e1.y};
  }
};

template <class T>
__device__ __forceinline__ void sincos_tab(const float2* tab, T x, T& s, T& c) {
  const T t = fma_(x, bc<T>(32.f), bc<T>(12582912.f));
  const T k = add_(t, bc<T>(-12582912.f));
  const T r = fma_(k, bc<T>(-1.0f / 32.f), x);
  T S, C;
  Tab<T>::load(tab, t, S, C);
  const T r2 = mul_(r, r);
  const T a = mul_(bc<T>(-0.5f), r2);
  const T sr = fma_(mul_(r2, r), bc<T>(-0.16666667f), r);
  s = fma_(S, a, fma_(C, sr, S));
  c = fma_(C, a, fma_(neg_(S), sr, C));
}
template <class T>
__device__ __forceinline__ T wrap_pi(T a) {
  const T k = add_(fma_(a, bc<T>(0.15915494309189535f), bc<T>(12582912.f)), bc<T>(-12582912.f));
  return fma_(neg_(k), bc<T>(-1.7484555e-07f), fma_(neg_(k), bc<T>(6.2831855f), a));
}

struct P { float K[12], u0[2], umin[2], umax[2], c[3], dt, r0[2]; const float* x0; float* xf; float* cost; int T; long long N; };

template <class T> struct Env;
template <> struct Env<float> {
  static constexpr int PER = 1;
  static __device__ __forceinline__ float load(const float* p, long long e, int i, long long N) { return p[e * 6 + i]; }
  static __device__ __forceinline__ void store(float* p, long long e, int i, float v) { p[e * 6 + i] = v; }
  static __device__ __forceinline__ void store1(float* p, long long e, float v) { p[e] = v; }
};
template <> struct Env<f2> {
  static constexpr int PER = 2;
  static __device__ __forceinline__ f2 load(const float* p, long long e, int i, long long N) { return f2{p[e * 6 + i], p[(e + 1) * 6 + i]}; }
  static __device__ __forceinline__ void store(float* p, long long e, int i, f2 v) { p[e * 6 + i] = v.a; p[(e + 1) * 6 + i] = v.b; }
  static __device__ __forceinline__ void store1(float* p, long long e, f2 v) { p[e] = v.a; p[e + 1] = v.b; }
};

template <class T>
__global__ void __launch_bounds__(256) rollout(const __grid_constant__ P a) {
  __shared__ __align__(16) float2 tab[kSize];
  for (int i = threadIdx.x; i < kSize; i += blockDim.x) {
    double s, c;
    sincos((double)(i - kHalf) / 32.0, &s, &c);
    tab[i] = make_float2((float)s, (float)c);
  }
  __syncthreads();
  constexpr int PER = Env<T>::PER;
  for (long long blk = blockIdx.x; blk * 256 * PER < a.N; blk += gridDim.x) {
    const long long env = (blk * 256 + threadIdx.x) * PER;
    if (env + PER > a.N) break;
    T z[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) z[i] = Env<T>::load(a.x0, env, i, a.N);
    T J = bc<T>(0.f);
#pragma unroll 4
    for (int t = 0; t < a.T; ++t) {
      T s, c;
      sincos_tab<T>(tab, z[2], s, c);
      T u[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        T acc = bc<T>(a.u0[k]);
#pragma unroll
        for (int i = 0; i < 6; ++i) acc = fma_(bc<T>(-a.K[k * 6 + i]), z[i], acc);
        u[k] = clampf(acc, a.umin[k], a.umax[k]);
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) J = fma_(z[i], z[i], J);
#pragma unroll
      for (int k = 0; k < 2; ++k) { const T y = add_(u[k], bc<T>(a.r0[k])); J = fma_(y, y, J); }
      const T sm = mul_(add_(u[0], u[1]), bc<T>(a.c[1]));
      T d[6];
      d[0] = z[3]; d[1] = z[4]; d[2] = z[5];
      d[3] = mul_(neg_(s), sm);
      d[4] = fma_(c, sm, bc<T>(-a.c[0]));
      d[5] = mul_(add_(u[0], neg_(u[1])), bc<T>(a.c[2]));
#pragma unroll
      for (int i = 0; i < 6; ++i) z[i] = fma_(d[i], bc<T>(a.dt), z[i]);
      z[2] = wrap_pi<T>(z[2]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) Env<T>::store(a.xf, env, i, z[i]);
    Env<T>::store1(a.cost, env, mul_(J, bc<T>(a.dt)));
  }
}

// third variant: one environment per thread (the production layout), only the cost accumulation packed — the six z_i^2
// terms as three FFMA2 on register pairs (z0, z1), (z2, z3), (z4, z5) and the two input terms as one: 4 issue slots
// instead of 8 for the same FMA-pipe time
__global__ void __launch_bounds__(256) rollout_cost2(const __grid_constant__ P a) {
  __shared__ __align__(16) float2 tab[kSize];
  for (int i = threadIdx.x; i < kSize; i += blockDim.x) {
    double s, c;
    sincos((double)(i - kHalf) / 32.0, &s, &c);
    tab[i] = make_float2((float)s, (float)c);
  }
  __syncthreads();
  for (long long blk = blockIdx.x; blk * 256 < a.N; blk += gridDim.x) {
    const long long env = blk * 256 + threadIdx.x;
    if (env >= a.N) break;
    float z[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) z[i] = a.x0[env * 6 + i];
    f2 J2 = f2{0.f, 0.f};
#pragma unroll 4
    for (int t = 0; t < a.T; ++t) {
      float s, c;
      sincos_tab<float>(tab, z[2], s, c);
      float u[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float acc = a.u0[k];
#pragma unroll
        for (int i = 0; i < 6; ++i) acc = fma_(-a.K[k * 6 + i], z[i], acc);
        u[k] = clampf(acc, a.umin[k], a.umax[k]);
      }
#pragma unroll
      for (int i = 0; i < 6; i += 2) { const f2 zz = f2{z[i], z[i + 1]}; J2 = fma_(zz, zz, J2); }
      { const f2 y = add_(f2{u[0], u[1]}, f2{a.r0[0], a.r0[1]}); J2 = fma_(y, y, J2); }
      const float sm = (u[0] + u[1]) * a.c[1];
      float d[6];
      d[0] = z[3]; d[1] = z[4]; d[2] = z[5];
      d[3] = -s * sm;
      d[4] = fma_(c, sm, -a.c[0]);
      d[5] = (u[0] - u[1]) * a.c[2];
#pragma unroll
      for (int i = 0; i < 6; ++i) z[i] = fma_(d[i], a.dt, z[i]);
      z[2] = wrap_pi<float>(z[2]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) a.xf[env * 6 + i] = z[i];
    a.cost[env] = (J2.a + J2.b) * a.dt;
  }
}

int main(int argc, char** argv) {
  const long long N = argc > 1 ? atoll(argv[1]) : (1ll << 24);
  const int T = argc > 2 ? atoi(argv[2]) : 1000;
  P p{};
  const float K[12] = {0.7071f, -0.7071f, -2.9f, 1.0f, -1.3f, -0.55f, 0.7071f, 0.7071f, 2.9f, 1.0f, 1.3f, 0.55f};
  memcpy(p.K, K, sizeof K);
  p.u0[0] = p.u0[1] = 4.905f; p.umin[0] = p.umin[1] = 0.f; p.umax[0] = p.umax[1] = 20.f;
  p.c[0] = 9.81f; p.c[1] = 1.0f; p.c[2] = 1.0f / 0.0025f * 0.25f * 0.01f; p.dt = 0.02f; p.r0[0] = p.r0[1] = -4.905f;
  p.T = T; p.N = N;
  std::vector<float> h(N * 6);
  uint32_t s = 12345;
  for (auto& v : h) { s = s * 1664525u + 1013904223u; v = ((s >> 8) * (1.0f / 8388608.f) - 1.0f) * 0.5f; }
  float *x0, *xf[2], *cost[2];
  cudaMalloc(&x0, N * 24);
  cudaMemcpy(x0, h.data(), N * 24, cudaMemcpyHostToDevice);
  for (int v = 0; v < 2; ++v) { cudaMalloc(&xf[v], N * 24); cudaMalloc(&cost[v], N * 4); }
  p.x0 = x0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int v = 0; v < 2; ++v) {
    p.xf = xf[v]; p.cost = cost[v];
    for (int occ : {8, 6, 4, 3, 2}) {
      float best = 1e30f;
      for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        if (v == 0) rollout<float><<<148 * occ, 256>>>(p); else rollout<f2><<<148 * occ, 256>>>(p);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
      }
      printf("%s grid=148x%d  %.3f ms  %.3e env-steps/s  (%s)\n", v ? "packed f32x2" : "scalar      ", occ, best, (double)N * T / best * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
  }
  {
    p.xf = xf[1]; p.cost = cost[1];
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
      cudaEventRecord(e0);
      rollout_cost2<<<148 * 8, 256>>>(p);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r > 0 && ms < best) best = ms;
    }
    printf("scalar + packed cost (unroll 4)  %.3f ms  %.3e env-steps/s  (%s)\n", best, (double)N * T / best * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
  std::vector<float> a(N * 6), b(N * 6);
  cudaMemcpy(a.data(), xf[0], N * 24, cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), xf[1], N * 24, cudaMemcpyDeviceToHost);
  long long diff = 0, nan = 0;
  for (long long i = 0; i < N * 6; ++i) { if (memcmp(&a[i], &b[i], 4)) ++diff; if (std::isnan(a[i])) ++nan; }
  printf("final states differing bitwise: %lld of %lld (NaN: %lld)\n", diff, N * 6, nan);
  return 0;
}

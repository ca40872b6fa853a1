// Counter-based sampling of initial states / sampled states on the device ("generated on device with a
// counter-based generator, seed = 1234 + rank").
//
// The reference draws one state at a time from the global NumPy RNG: x0 = wrap(U(-std, std) + mean)
// (dynamics/dynamics_basic.py:28-29; the seed data set of controller/vhjb.py:136-151 has the same form).  That stream is
// inherently serial; for 16M environments per GPU the states are generated where they are consumed, with Philox4x32-10
// keyed by the seed and COUNTED by the global sample index — so sample i is the same number whichever GPU, launch or
// thread produces it, and a batch split over N GPUs is the same batch.
//
//   r[c]    = Philox4x32-10(key = (seed_lo, seed_hi), counter = (i_lo, i_hi, c / 4, 0))[c % 4]     c = component
//   t       = 2 * (r >> 8) * 2^-24 - 1                                                       in [-1, 1), exact in fp32
//   x[i][c] = fma(std[c], t, mean[c]), then states_wrap on the angle components
// oracle/x0_stream.py is the bit-exact NumPy twin (the oracle consumes the same numbers without a device).
#include <cstdint>

#include "systems.cuh"

namespace hjb {

__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t* out) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct SampleArgs {
  float mean[HJB_MAX_N], std[HJB_MAX_N];
  int ang[2];
  int nang;
  uint32_t k0, k1;
  int64_t first, count;
  float* x;
};

template <int N>
__global__ void __launch_bounds__(256) sample_states_kernel(const __grid_constant__ SampleArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  const uint64_t g = (uint64_t)(a.first + i);
  float x[N];
#pragma unroll
  for (int q = 0; q < (N + 3) / 4; ++q) {
    uint32_t r[4];
    philox4x32_10(a.k0, a.k1, (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)q, 0u, r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * q + j;
      if (c < N) {
        const float t = fmaf((float)(r[j] >> 8), 1.1920928955078125e-07f, -1.0f);   // 2 (r >> 8) 2^-24 - 1, exact
        x[c] = fmaf(a.std[c], t, a.mean[c]);
      }
    }
  }
  for (int k = 0; k < a.nang; ++k) {
#pragma unroll
    for (int c = 0; c < N; ++c)
      if (c == a.ang[k]) x[c] = wrap_pi(x[c]);
  }
  store_row<N>(a.x, i, x);
}

}  // namespace hjb

using namespace hjb;

extern "C" int hjb_sample_states(int32_t sys_kind, int32_t n, const float* mean, const float* std, uint64_t seed, int64_t first,
                                 int64_t count, float* x, void* stream) {
  if (!mean || !std || n <= 0 || n > HJB_MAX_N || first < 0 || count < 0) return HJB_ERR_BAD_ARG;
  if (count == 0) return HJB_OK;
  if (!x) return HJB_ERR_BAD_ARG;
  SampleArgs a;
  for (int i = 0; i < HJB_MAX_N; ++i) { a.mean[i] = i < n ? mean[i] : 0.f; a.std[i] = i < n ? std[i] : 0.f; }
  a.nang = 0;
  a.ang[0] = a.ang[1] = -1;
  switch (sys_kind) {   // states_wrap: cartpole.py:52-64, acrobot.py:72-81, quadrotors.py:48-70, 151-170
    case HJB_SYS_CARTPOLE: a.nang = 1; a.ang[0] = 1; break;
    case HJB_SYS_ACROBOT: a.nang = 2; a.ang[0] = 0; a.ang[1] = 1; break;
    case HJB_SYS_QUAD2D: a.nang = 1; a.ang[0] = 2; break;
    case HJB_SYS_QUAD10D: a.nang = 2; a.ang[0] = 3; a.ang[1] = 4; break;
    case HJB_SYS_LINEAR: break;
    default: return HJB_ERR_UNSUPPORTED;
  }
  for (int k = 0; k < a.nang; ++k)
    if (a.ang[k] >= n) return HJB_ERR_BAD_ARG;
  a.k0 = (uint32_t)seed;
  a.k1 = (uint32_t)(seed >> 32);
  a.first = first;
  a.count = count;
  a.x = x;
  const unsigned grid = (unsigned)((count + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (n) {
    case 2: sample_states_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 4: sample_states_kernel<4><<<grid, 256, 0, st>>>(a); break;
    case 6: sample_states_kernel<6><<<grid, 256, 0, st>>>(a); break;
    case 10: sample_states_kernel<10><<<grid, 256, 0, st>>>(a); break;
    default: return HJB_ERR_UNSUPPORTED;
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HJB_OK : (int)e;
}

// Time to the goal ball of recorded trajectories (examples/double_integrator_optimal_time.ipynb cell 20: the loop keeps
// `optimal_t = min(t_k, optimal_t)` whenever the state AFTER step k satisfies x^T x <= metric; T if it never does).
#include "mintime_ctl.cuh"

namespace hjb {

// xs: time-major [rows][N][n], row 0 = the initial states (hjb_rollout's layout with record_stride = 1).  One thread per
// environment; a warp reads 32 consecutive rows of n floats per time slice (coalesced).  The scan stops at the first hit.
__global__ void __launch_bounds__(256) first_hit_kernel(const float* __restrict__ xs, int64_t N, int n, int rows, float metric,
                                                        float dt, float t_max, float* __restrict__ t_hit) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  float t = t_max;
  for (int r = 1; r < rows; ++r) {
    const float* x = xs + ((int64_t)r * N + e) * n;
    float s = 0.f;
    for (int i = 0; i < n; ++i) s = fmaf(x[i], x[i], s);
    if (s <= metric) {
      t = fminf(t_max, (float)(r - 1) * dt);
      break;
    }
  }
  t_hit[e] = t;
}

cudaError_t first_hit(const float* xs, int64_t N, int32_t n, int32_t rows, float metric, float dt, float t_max, float* t_hit,
                      cudaStream_t st) {
  first_hit_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(xs, N, n, rows, metric, dt, t_max, t_hit);
  return cudaGetLastError();
}

}  // namespace hjb

// Device-resident replay buffer: the reference keeps (state, cost, done) samples in a
// deque(maxlen) wrapped by a torch Dataset / shuffling DataLoader (controller/vhjb.py:62-73, :153-154) and extends it
// trajectory by trajectory (:299-305).  Here the samples live in a ring in HBM; the records of a batched learned-policy
// rollout are appended by one launch and a shuffled minibatch is gathered by one launch.
#include <cstdint>

#include "hjb_common.cuh"

namespace hjb {

// one thread per (t, e); sample (t, e) has sequence number offsets[e] + t in deque order (trajectory by trajectory)
__global__ void __launch_bounds__(256) replay_append_kernel(const float* __restrict__ rec_x, const float* __restrict__ rec_cost,
                                                            const float* __restrict__ rec_done, const int64_t* __restrict__ offsets,
                                                            int64_t T1, int64_t N, int n, int64_t skip, int64_t tail, int64_t cap,
                                                            float* __restrict__ buf_x, float* __restrict__ buf_cost,
                                                            float* __restrict__ buf_done) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T1 * N) return;
  const float d = rec_done[i];
  if (d < 0.f) return;                       // the trajectory had ended: no sample
  const int64_t t = i / N, e = i - t * N;
  const int64_t seq = offsets[e] + t;
  if (seq < skip) return;                    // would be pushed out of the deque by this same extend
  const int64_t pos = (tail + seq) % cap;
  for (int k = 0; k < n; ++k) buf_x[pos * n + k] = rec_x[i * n + k];
  buf_cost[pos] = rec_cost[i];
  buf_done[pos] = d;
}

// out row b <- ring row index[b]; one thread per (b, component), components: n state entries, cost, done
__global__ void __launch_bounds__(256) replay_gather_kernel(const float* __restrict__ buf_x, const float* __restrict__ buf_cost,
                                                            const float* __restrict__ buf_done, const int64_t* __restrict__ index,
                                                            int64_t B, int n, float* __restrict__ xs, float* __restrict__ costs,
                                                            float* __restrict__ dones) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = n + 2;
  if (i >= B * w) return;
  const int64_t b = i / w;
  const int k = (int)(i - b * w);
  const int64_t src = index[b];
  if (k < n) xs[b * n + k] = buf_x[src * n + k];
  else if (k == n) costs[b] = buf_cost[src];
  else dones[b] = buf_done[src];
}

}  // namespace hjb

using namespace hjb;

extern "C" {

int hjb_replay_append(const float* rec_x, const float* rec_cost, const float* rec_done, const int64_t* offsets, int64_t T1,
                      int64_t N, int32_t n, int64_t skip, int64_t tail, int64_t capacity, float* buf_x, float* buf_cost,
                      float* buf_done, void* stream) {
  if (T1 < 0 || N < 0 || n <= 0 || n > HJB_MAX_N || capacity <= 0 || tail < 0 || tail >= capacity || skip < 0) return HJB_ERR_BAD_ARG;
  if (T1 * N == 0) return HJB_OK;
  if (!rec_x || !rec_cost || !rec_done || !offsets || !buf_x || !buf_cost || !buf_done) return HJB_ERR_BAD_ARG;
  const int64_t total = T1 * N;
  replay_append_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rec_x, rec_cost, rec_done, offsets, T1, N, n, skip, tail, capacity, buf_x, buf_cost, buf_done);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HJB_OK : (int)e;
}

int hjb_replay_gather(const float* buf_x, const float* buf_cost, const float* buf_done, const int64_t* index, int64_t B,
                      int32_t n, float* xs, float* costs, float* dones, void* stream) {
  if (B < 0 || n <= 0 || n > HJB_MAX_N) return HJB_ERR_BAD_ARG;
  if (B == 0) return HJB_OK;
  if (!buf_x || !buf_cost || !buf_done || !index || !xs || !costs || !dones) return HJB_ERR_BAD_ARG;
  const int64_t total = B * (n + 2);
  replay_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(buf_x, buf_cost, buf_done, index, B,
                                                                                         n, xs, costs, dones);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HJB_OK : (int)e;
}

}  // extern "C"

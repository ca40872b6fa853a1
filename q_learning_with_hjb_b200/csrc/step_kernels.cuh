// Argument blocks and host launchers of the per-step kernels (step_kernels.cu), shared with api.cu.
#pragma once
#include "hjb_common.cuh"

namespace hjb {

struct DynArgs {
  DevSys sys;
  const float* x;
  const float* u;
  float* f;
  float* g;
  float* xdot;
  float* x_next;
  int64_t B;
};

struct CtlArgs {
  DevSys sys;
  DevCtl ctl;
  const float* x;
  float* u;
  int64_t B;
};

// One step of the learned-policy rollout bookkeeping (controller/vhjb.py:171-193), see policy_step_kernel.
struct PolicyStepArgs {
  DevSys sys;
  float xf[HJB_MAX_N], lo[HJB_MAX_N], hi[HJB_MAX_N], uf[HJB_MAX_M];
  float Q[HJB_MAX_N * HJB_MAX_N], R[HJB_MAX_M * HJB_MAX_M], P[HJB_MAX_N * HJB_MAX_N];
  int terminal;
  float* x;            // [N, n] in/out
  const float* u;      // [N, m] the policy's control at x
  float* alive;        // [N] 1 while the trajectory runs
  float* total_cost;   // [N] running sum of the sample costs
  float* rec_x;        // [N, n] the sample's state            (nullable)
  float* rec_cost;     // [N]    the sample's cost             (nullable)
  float* rec_done;     // [N]    0 / 1, or -1: no sample       (nullable)
  int64_t N;
};
cudaError_t step_policy(int kind, const PolicyStepArgs& a, cudaStream_t st);

cudaError_t step_dynamics(int kind, const DynArgs& a, int integ, bool fast, cudaStream_t st);
cudaError_t step_control(int sys_kind, int ctl_kind, const CtlArgs& a, bool fast, cudaStream_t st);
cudaError_t step_wrap(int kind, int n, float* x, int64_t B, cudaStream_t st);
cudaError_t fma_probe(float* sink, int64_t sink_len, int iters, double* flops, cudaStream_t st);

}  // namespace hjb

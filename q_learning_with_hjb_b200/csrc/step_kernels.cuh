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

cudaError_t step_dynamics(int kind, const DynArgs& a, int integ, bool fast, cudaStream_t st);
cudaError_t step_control(int sys_kind, int ctl_kind, const CtlArgs& a, bool fast, cudaStream_t st);
cudaError_t step_wrap(int kind, int n, float* x, int64_t B, cudaStream_t st);
cudaError_t fma_probe(float* sink, int64_t sink_len, int iters, double* flops, cudaStream_t st);

}  // namespace hjb

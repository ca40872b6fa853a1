#!/bin/bash
# Every capture of a round, back to back (one B200, ~12 min):  gpurun --timeout 1500 -- 'ROUND=r02 bash profiles/capture_all.sh'
# then, here:  python profiles/summarize_ncu.py <the six tags>
R=${ROUND:-r02}
TAG=${R}_rollout_quad2d_euler KERNEL=rollout_kernel bash profiles/gpu_profile.sh
TAG=${R}_rollout_quad2d_rk4 KERNEL=rollout_kernel BENCH_ARGS="--integrator rk4" bash profiles/gpu_profile.sh
TAG=${R}_rollout_acrobot_euler KERNEL=rollout_kernel BENCH_ARGS="--workload acrobot_es" bash profiles/gpu_profile.sh
TAG=${R}_vhjb_quad10d_tc KERNEL=vhjb_tc_kernel BENCH_ARGS="--workload vhjb_quad10d" bash profiles/gpu_profile.sh
TAG=${R}_vhjb_quad10d_tc_residual KERNEL=vhjb_tc_residual2_kernel BENCH_ARGS="--workload vhjb_quad10d" bash profiles/gpu_profile.sh
TAG=${R}_vhjb_di_tc_sin KERNEL=vhjb_tc_kernel BENCH_ARGS="--workload vhjb_di" bash profiles/gpu_profile.sh

"""Launch latency of SMALL rollouts (C1 as BASELINE.json states it: 4096 cart-pole environments x 500 steps = 16 CTAs): what
a launch costs when the per-CTA set-up (the fast path's trig tables) is not amortised over many environment blocks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from q_learning_with_hjb_b200.workloads import make_controller, make_dynamics
from q_learning_with_hjb_b200.rollout import BatchedRollout, RunningCost

for envs, T in ((4096, 500), (4096, 1), (256, 500), (65536, 500)):
    for fast in (True, False):
        for integ in ("euler", "rk4"):
            dyn = make_dynamics("cartpole"); dyn.fast_trig = fast
            ctl = make_controller("cartpole_lqr", dyn)
            plan = BatchedRollout(dyn, ctl, envs, T, integrator=integ, record_stride=0,
                                  cost=RunningCost(np.eye(4), np.eye(1), ctl.xf, ctl.uf))
            x0 = dyn.sample_initial_states(envs)
            for _ in range(5):
                plan.launch(x0)
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); plan.launch(x0); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            print(f"envs {envs:6d} T {T:4d} fast {fast!s:5} {integ:5}: median {np.median(ts):8.1f} us  min {min(ts):8.1f} us")

"""Developer tool: throughput of learned-policy rollouts — VHJBController.
rollout_trajectories, i.e. one fused value-net launch (hjb_vhjb_residual -> u) and one hjb_policy_step launch per step for
ALL trajectories.  Run on a GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from q_learning_with_hjb_b200.configs import gin_compat as gin
from q_learning_with_hjb_b200.configs.controller.vhjb_controller_config import VHJBControllerConfig
from q_learning_with_hjb_b200.controller.vhjb import VHJBController
from q_learning_with_hjb_b200.workloads import PKG, make_dynamics

for kind, cfgfile in (("cartpole", "cartpole_vhjb_controller.gin"), ("quad2d", "quadrotors2DHovering_vhjb_controller.gin")):
    dyn = make_dynamics(kind)
    gin.parse_config_file(os.path.join(PKG, "configs", "controller", cfgfile))
    cfg = VHJBControllerConfig()
    cfg.num_of_interior_data = cfg.num_of_boundary_data = 4
    ctl = VHJBController(dyn, cfg)
    for N, T in ((20, 200), (4096, 200), (1 << 20, 50)):
        x0 = torch.as_tensor(dyn.get_initial_states(N).astype(np.float32)).cuda()
        ctl.rollout_trajectories(x0, max_steps=T)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            ctl.rollout_trajectories(x0, max_steps=T)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"{kind}: {N} trajectories x {T} steps: {ms:.2f} ms -> {N * T / ms * 1e3:.3e} env-steps/s ({ms / T * 1e3:.1f} us per step)")

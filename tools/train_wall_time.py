"""Developer tool: wall time of VHJBController.train() with the reference's own gin configs (run on a GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from q_learning_with_hjb_b200.configs import gin_compat as gin
from q_learning_with_hjb_b200.configs.controller.vhjb_controller_config import VHJBControllerConfig
from q_learning_with_hjb_b200.controller.vhjb import VHJBController
from q_learning_with_hjb_b200.workloads import PKG, make_dynamics

for kind, cfgfile in (("linear", "linear_vhjb_controller.gin"), ("cartpole", "cartpole_vhjb_controller.gin"),
                      ("quad2d", "quadrotors2DHovering_vhjb_controller.gin")):
    dyn = make_dynamics(kind)
    gin.parse_config_file(os.path.join(PKG, "configs", "controller", cfgfile))
    cfg = VHJBControllerConfig()
    ctl = VHJBController(dyn, cfg)
    torch.cuda.synchronize()
    t0 = time.time()
    lists = ctl.train()
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"{kind}: train() {cfg.epochs} epochs x {cfg.num_of_trajectories_per_epoch} trajectories x <= {cfg.maximum_step} steps, "
          f"batch {cfg.batch_size}: {dt:.1f} s, {ctl.update_counter} updates, buffer {len(ctl.replay_buffer)}, "
          f"final hjb loss {lists[4][-1]:.4f}, avg trajectory cost {lists[0][-1]:.2f}, length {lists[2][-1]:.1f}")

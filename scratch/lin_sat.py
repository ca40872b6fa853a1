import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from q_learning_with_hjb_b200.configs import gin_compat as gin
from q_learning_with_hjb_b200.configs.controller.vhjb_controller_config import VHJBControllerConfig
from q_learning_with_hjb_b200.controller import vhjb as V
from tests.helpers import PKG, make_dynamics
dyn = make_dynamics("linear")
gin.parse_config_file(os.path.join(PKG, "configs", "controller", "linear_vhjb_controller.gin"))
ctl = V.VHJBController(dyn, VHJBControllerConfig())
k = ctl.kernels
orig = k.saturated_total
calls = []
def spy(reset=True):
    v = orig(reset)
    calls.append(v)
    return v
k.saturated_total = spy
lists = ctl.train()
print("sat per epoch:", calls)
print("impl", k.impl, "final hjb", lists[4][-1], "first", lists[4][0])

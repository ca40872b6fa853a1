"""Developer tool: end-to-end time of VhjbKernels.train_step_host and BatchedRollout.run_host against the number of
pipeline pieces (run on a GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from oracle import vhjb_oracle as V
from tests.helpers_vhjb import flat_params, make_kernels, sample_batch
from q_learning_with_hjb_b200.controller.vhjb import AdamState


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


B = 1 << 20
k, p = make_kernels("quad10d")
params = torch.as_tensor(flat_params(V.init_weights(p.sys.n, seed=0))).cuda()
opt = AdamState(0, torch.zeros_like(params), torch.zeros_like(params))
xs, dones, costs = sample_batch("quad10d", B, seed=1)
host = [torch.as_tensor(a).pin_memory() for a in (xs, dones, costs)]
dev = [h.cuda() for h in host]
print("device batch train_step: %.3f ms" % timed(lambda: k.train_step(params, opt, *dev, 1e-5, 1e-3)))
for c in (None, 1, 4):
    def f():
        k.train_step_host(params, opt, host[0], host[1], host[2], 1e-5, 1e-3, chunks=c)
        torch.cuda.current_stream().synchronize()
    print("train_step_host chunks=%s: %.3f ms" % (c, timed(f)))

# ---- rollout run_host: number of environment ranges ----
from tests.helpers import make_controller, make_dynamics
from q_learning_with_hjb_b200.rollout import BatchedRollout, RunningCost
dyn = make_dynamics("quad2d"); dyn.fast_trig = True
ctl = make_controller("quad2d_hover", dyn)
N = 1 << 24
plan = BatchedRollout(dyn, ctl, N, 1000, cost=RunningCost(np.eye(6), np.eye(2), np.zeros(6), np.asarray(ctl.uf)))
pin = plan.pinned_x0()
pin.copy_(torch.rand((N, 6)) * 2 - 1)
x0d = pin.cuda()
print("rollout kernel alone: %.2f ms" % timed(lambda: plan.launch(x0d), 5))
for c in (1, 4, 8, 12, 16, 24, 32):
    print("run_host chunks=%2d: %.2f ms" % (c, timed(lambda: plan.run_host(pin, copy_in=False, chunks=c), 5)))

"""Developer tool (kept as the record of how the fp16 overflow of the adjoint chain was found): runs VHJBController.train()
with the reference's linear config, stops at the first update that leaves non-finite weights, saves that minibatch and
compares the tensor-core and fp32 gradients on it.  Run on a GPU box."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from q_learning_with_hjb_b200.configs import gin_compat as gin
from q_learning_with_hjb_b200.configs.controller.vhjb_controller_config import VHJBControllerConfig
from q_learning_with_hjb_b200.controller.vhjb import VHJBController
from tests.helpers import PKG, make_dynamics
dyn = make_dynamics("linear")
gin.parse_config_file(os.path.join(PKG, "configs", "controller", "linear_vhjb_controller.gin"))
cfg = VHJBControllerConfig()
ctl = VHJBController(dyn, cfg)
k = ctl.kernels
orig = ctl.params_update
state = {"n": 0, "bad": False}
def wrapped(params, states, opt, xs, dones, costs, reg):
    before = params.flat.clone(); mu = opt.mu.clone(); nu = opt.nu.clone(); cnt = opt.count
    out = orig(params, states, opt, xs, dones, costs, reg)
    state["n"] += 1
    if not state["bad"] and not bool(torch.isfinite(params.flat).all()):
        state["bad"] = True
        print("first non-finite params at update", state["n"], "reg", reg, "saturated", k.saturated())
        print("grad finite:", bool(torch.isfinite(k.grad).all()), "sums", k.sums.tolist(), "norm", k.norm.tolist())
        print("dones sum", float(dones.sum()), "min cost of done", float(costs[dones > 0].min()) if (dones > 0).any() else None)
        torch.save({"params": before.cpu(), "xs": xs.cpu(), "dones": dones.cpu(), "costs": costs.cpu(), "reg": reg}, "gpurun_out/bad_batch.pt")
        # same batch through the CUDA-core kernel
        os.environ["HJB_VHJB_IMPL"] = "simt"
        k.counts(dones, 0.0); k.norm.add_(cfg.epsilon)
        g2 = k.loss_grad(before, xs, dones, costs, reg)[0].clone()
        os.environ.pop("HJB_VHJB_IMPL")
        k.counts(dones, 0.0); k.norm.add_(cfg.epsilon)
        g1 = k.loss_grad(before, xs, dones, costs, reg)[0].clone()
        print("simt grad finite:", bool(torch.isfinite(g2).all()), "max", float(g2.abs().max()), "| tc grad finite:", bool(torch.isfinite(g1).all()),
              "nan count", int((~torch.isfinite(g1)).sum()), "sat", k.saturated())
        idx = (~torch.isfinite(g1)).nonzero().flatten()[:10].tolist(); print("bad idx", idx)
    return out
ctl.params_update = wrapped
ctl.epochs = 100
ctl.train()
print("updates", state["n"], "bad", state["bad"])

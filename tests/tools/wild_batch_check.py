"""Developer tool: gradient error of the tensor-core and fp32 kernels on tests/golden/vhjb_linear_overflow_batch.npz against
the float64 oracle, for the whole minibatch and for its near-goal states one by one (where the truncating accumulation of
tcgen05.mma meets a ~100-fold cancellation in dV/dx).  Run on a GPU box."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import vhjb_oracle as V
from tests.helpers_vhjb import make_kernels
d = np.load("tests/golden/vhjb_linear_overflow_batch.npz")
k, p = make_kernels("linear")
params = torch.as_tensor(d["params"]).cuda()
xs, dones, costs, reg = d["xs"], d["dones"], d["costs"], float(d["reg"])
dev = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda() for a in (xs, dones, costs)]
W = np.split(d["params"].astype(np.float64), [256, 256 + 128 * 128])
orc = V.VhjbOracle(p, [W[0].reshape(2, 128), W[1].reshape(128, 128), W[2].reshape(128, 64)])
_, _, _, grads, _ = orc.loss_and_grad(xs, dones, costs, reg)
go = np.concatenate([g.reshape(-1) for g in grads])
# oracle in float32-rounded inputs? (inputs are already float32)
for impl in ("tensor", "simt"):
    k.impl = impl
    k.counts(dev[1], p.eps)
    g = k.loss_grad(params, *dev, reg)[0].cpu().numpy().astype(np.float64)
    off = 0; errs = []
    for gi in grads:
        sl = slice(off, off + gi.size); off += gi.size
        errs.append(np.abs(g[sl] - go[sl]).max() / np.abs(go[sl]).max())
    print(impl, ["%.2e" % e for e in errs], "sat", k.saturated())
# without the two heavy states
keep = np.ones(len(xs), bool); keep[30] = False

def errs_for(xs_, dones_, costs_, norm=None, label=""):
    dev = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda() for a in (xs_, dones_, costs_)]
    out = {}
    for impl in ("tensor", "simt"):
        k.impl = impl
        if norm is None:
            k.counts(dev[1], p.eps)
        else:
            k.norm.copy_(torch.tensor(norm, dtype=torch.float32))
        out[impl] = k.loss_grad(params, *dev, reg)[0].cpu().numpy().astype(np.float64)
    k.impl = "tensor"
    d_ = np.abs(out["tensor"] - out["simt"])
    ref = out["simt"]
    print(label, "tensor vs simt: W1 %.2e W2 %.2e W3 %.2e" % (d_[:256].max() / np.abs(ref[:256]).max(), d_[256:256 + 16384].max() / np.abs(ref[256:256 + 16384]).max(),
                                                   d_[256 + 16384:].max() / np.abs(ref[256 + 16384:]).max()), "| |grad| max %.3g" % np.abs(ref).max())

full_norm = [float((1 - dones).sum() + 1e-10), float(dones.sum() + 1e-10)]
errs_for(xs, dones, costs, label="full batch          ")
x2 = xs.copy(); x2[30] = xs[31]
errs_for(x2, dones, costs, label="state 30 replaced   ")
errs_for(xs[30:31], dones[30:31], costs[30:31], norm=full_norm, label="state 30 alone      ")
heavy = [i for i in range(len(xs)) if np.abs(xs[i]).max() < 0.02]
print("near-goal states:", heavy, [xs[i].tolist() for i in heavy])
for i in heavy:
    errs_for(xs[i:i + 1], dones[i:i + 1], costs[i:i + 1], norm=full_norm, label=f"state {i} alone      ")

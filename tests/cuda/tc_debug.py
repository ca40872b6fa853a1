"""Developer check (not a pytest): tensor-core vhjb kernel vs the torch-fp64 oracle, printing error statistics."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import vhjb_oracle as V
from tests.helpers_vhjb import flat_params, make_kernels, sample_batch

def dev(*arrays):
    return [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda() for a in arrays]

def run(name, B, reg=0.37, do_grad=True, near=0.0):
    k, p = make_kernels(name)
    W32 = [w.astype(np.float32) for w in V.init_weights(p.sys.n, seed=2)]
    xs, dones, costs = sample_batch(name, B, seed=3)
    if near > 0:   # a few states very close to the goal: adjoint seeds ~ 1 / near^2 times the typical ones
        rng = np.random.default_rng(5)
        idx = rng.choice(B, size=min(7, B), replace=False)
        xs[idx] = (p.xf + near * rng.uniform(-1, 1, size=(len(idx), p.sys.n))).astype(np.float32)
        dones[idx] = 0
    orc = V.VhjbOracle(p, [w.astype(np.float64) for w in W32])
    params = torch.as_tensor(flat_params(W32)).cuda()
    xd, dd, cd = dev(xs, dones, costs)
    out, sums = k.residual(params, xd, dd, cd)
    torch.cuda.synchronize()
    q = orc.pieces(xs, running=costs if p.residual_form == "min_time" else None)
    for key in ("V", "p", "u", "r"):
        a = out[key].cpu().numpy().astype(np.float64).reshape(B, -1); b = q[key].detach().numpy().reshape(B, -1)
        scale = np.maximum(np.abs(b), np.abs(b).mean(axis=0, keepdims=True) + 1e-30)
        err = (np.abs(a - b) / scale).max(axis=1)
        print(f"  {name} B={B} {key}: frac>1e-4 {np.mean(err > 1e-4):.2e} median {np.median(err):.2e} p99 {np.quantile(err, .99):.2e} max {err.max():.2e} nan {np.isnan(a).sum()}")
    hjb, term, _ = orc.losses(xs, dones, costs)
    d = dones.astype(np.float64)
    s = sums.cpu().numpy().astype(np.float64)
    if p.residual_form == "normalized":
        print(f"  residual sums: hjb relerr {abs(s[0] / ((1 - d).sum() + p.eps) - float(hjb)) / float(hjb):.2e} term relerr {abs(s[1] / (d.sum() + p.eps) - float(term)) / float(term):.2e}")
    if not do_grad:
        return
    if p.residual_form == "min_time":
        k.norm.copy_(torch.tensor([float(B), 1.0]))
    else:
        k.counts(dd, p.eps)
    grad, sums = k.loss_grad(params, xd, dd, cd, reg)
    torch.cuda.synchronize()
    total, hjb, term, grads, _ = orc.loss_and_grad(xs, dones, costs, reg)
    g = grad.cpu().numpy().astype(np.float64)
    off = 0
    for nm, gi in zip(("dW1", "dW2", "dW3"), grads):
        sl = slice(off, off + gi.size); off += gi.size
        print(f"  {nm}: normwise err {np.abs(g[sl] - gi.reshape(-1)).max() / np.abs(gi).max():.2e} (|g|max {np.abs(gi).max():.3e}) nan {np.isnan(g[sl]).sum()}")
    norm = k.norm.cpu().numpy().astype(np.float64); s = sums.cpu().numpy().astype(np.float64)
    print(f"  grad-pass sums: hjb relerr {abs(s[0] / norm[0] - hjb) / hjb:.2e}  saturated states: {k.saturated()}")

if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["quad10d"]
    sizes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 4101]
    print("impl:", os.environ.get("HJB_VHJB_IMPL", "tensor (default)"))
    for n in names:
        for B in sizes:
            t0 = time.time(); run(n, B, near=float(os.environ.get("NEAR", "0"))); print(f"  [{time.time() - t0:.1f}s]")

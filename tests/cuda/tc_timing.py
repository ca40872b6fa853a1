"""Developer probe: per-step cycle breakdown of the tensor-core vhjb kernel (HJB_TC_DEBUG_TIMING=1) and wall times."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("HJB_TC_DEBUG_TIMING", "1")
import torch
from oracle import vhjb_oracle as V
from tests.helpers_vhjb import flat_params, make_kernels, sample_batch

name = sys.argv[1] if len(sys.argv) > 1 else "quad10d"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
k, p = make_kernels(name)
W32 = [w.astype(np.float32) for w in V.init_weights(p.sys.n, seed=2)]
xs, dones, costs = sample_batch(name, B, seed=3)
params = torch.as_tensor(flat_params(W32)).cuda()
xd, dd, cd = (torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda() for a in (xs, dones, costs))
k.counts(dd, p.eps)
for _ in range(2):
    k.residual(params, xd, dd, cd)
    k.loss_grad(params, xd, dd, cd, 0.3)
torch.cuda.synchronize()
os.environ.pop("HJB_TC_DEBUG_TIMING")
for label, fn in (("residual", lambda: k.residual(params, xd, dd, cd)), ("loss_grad", lambda: k.loss_grad(params, xd, dd, cd, 0.3))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{label}: {ms:.3f} ms per {B} states = {B / ms * 1e3:.3e} states/s")

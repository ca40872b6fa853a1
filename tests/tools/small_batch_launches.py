"""Developer tool: the kernels of ONE minibatch update at the reference's batch size (256 states) — run under
`ncu --metrics gpu__time_duration.sum` for the per-launch times, or plain for the CUDA-event time per update."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import vhjb_oracle as V
from q_learning_with_hjb_b200.controller.vhjb import AdamState
from tests.helpers_vhjb import flat_params, make_kernels, sample_batch

name = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
k, p = make_kernels(name)
xs, dones, costs = (torch.as_tensor(a).cuda() for a in sample_batch(name, B, seed=3))
w = torch.as_tensor(flat_params(V.init_weights(p.sys.n, seed=1))).cuda()
opt = AdamState(0, torch.zeros_like(w), torch.zeros_like(w))
for _ in range(10):
    k.train_step(w, opt, xs, dones, costs, 0.25, 1e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    k.train_step(w, opt, xs, dones, costs, 0.25, 1e-4)
e1.record()
torch.cuda.synchronize()
print(f"{name} B={B}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us per update (deferred states of the last launch: {k.deferred()})")

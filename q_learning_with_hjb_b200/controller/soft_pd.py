"""The notebooks' soft-PD baseline on the CUDA library.

``SoftPDValueApproximator`` (examples/cartpole_balancing.ipynb cell 6, examples/drone_hovering.ipynb cell 6) is the
unconstrained value net the reference compares its positive-definite one against: Dense layers WITH biases, a Dense(1)
head, tanh (cart-pole) or relu (drone) activations.  It is trained (cell 11 of both notebooks) on

    mean_i [ res_i + reg * max(0, V(xf) - V(x_i)) ],   res = |vdot + l| (cart-pole)  or  |vdot / (l + 1e-10) + 1| (drone)

after a warm-up on  |V - z^T P z|  (cart-pole)  or on the same residual under the LQR's control (drone).  Everything that
touches the net — V, dV/dx, u = clip(-R^-1 g^T dV/dx / 2 + uf), the losses and the full parameter gradient (weights and
biases) — is one fused fp32 kernel family (``csrc/softpd.cu``); Adam is ``hjb_adam``.  No CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.controller.controller_basic import Controller, lqr_gain
from q_learning_with_hjb_b200.controller.vhjb import AdamState, lecun_normal

FEATURES = (128, 128, 64, 1)
LOSS_FORMS = {"hjb": 0, "value_match": 1, "hjb_lqr": 2}


class SoftPDController(Controller):
    """The soft-PD value net with its control law, its three losses and their Adam updates.

    ``params`` is one flat fp32 device buffer [W1 | b1 | W2 | b2 | W3 | b3 | w4 | b4]; ``views()`` gives the Flax-style
    pytree ``{"Dense_k": {"kernel", "bias"}}`` onto it."""

    def __init__(self, dynamics, xf, uf, Q, R, activation: str = "tanh", normalized_residual: bool = False,
                 lr: float = 1e-3, epsilon: float = 1e-10, seed: int = 0, K=None, P=None) -> None:
        super().__init__()
        torch = L.require_cuda()
        self.torch = torch
        self.dynamics = dynamics
        self.n, self.m = dynamics.get_dimension()
        self.xf, self.uf = np.asarray(xf, dtype=np.float64), np.asarray(uf, dtype=np.float64)
        self.Q, self.R = np.asarray(Q, dtype=np.float64), np.asarray(R, dtype=np.float64).reshape(self.m, self.m)
        self.lr = float(lr)
        if K is None or P is None:      # the LQR about (xf, uf): warm-up target z^T P z / warm-up control -K z + uf
            A, B = dynamics.linearize(self.xf, self.uf)
            K, P = lqr_gain(A, B, self.Q, self.R)
        self.K, self.P = np.asarray(K, dtype=np.float64), np.asarray(P, dtype=np.float64)
        self.P_count = int(L.lib().hjb_softpd_param_count(self.n))
        rng = np.random.default_rng(seed)
        dims = [self.n, *FEATURES]
        parts = []
        for i in range(4):               # Flax Dense defaults: lecun-normal kernel, zero bias
            parts += [lecun_normal(rng, dims[i], dims[i + 1]).reshape(-1), np.zeros(dims[i + 1], dtype=np.float32)]
        self.params = torch.as_tensor(np.concatenate(parts).astype(np.float32)).cuda()
        assert self.params.numel() == self.P_count
        self.opt = AdamState(0, torch.zeros_like(self.params), torch.zeros_like(self.params))
        self.grad = torch.zeros_like(self.params)
        self.sums = torch.zeros(3, device="cuda", dtype=torch.float32)
        self.workspace = torch.zeros(int(L.lib().hjb_softpd_workspace_bytes(self.n)) // 4, device="cuda", dtype=torch.float32)
        self.net = L.HjbSoftPD()
        self.net.n, self.net.act = self.n, L.ACTIVATIONS[activation]
        self.net.normalized_residual = int(bool(normalized_residual))
        L.fill(self.net.xf, self.xf)
        L.fill(self.net.uf, self.uf)
        L.fill(self.net.Q, self.Q)
        L.fill(self.net.R, self.R)
        L.fill(self.net.Rinv, np.linalg.inv(self.R))
        L.fill(self.net.K, self.K)
        L.fill(self.net.P, self.P)
        self.net.eps = float(epsilon)

    def views(self):
        n, o, out = self.n, 0, {}
        dims = [n, *FEATURES]
        for i in range(4):
            k = self.params[o:o + dims[i] * dims[i + 1]].view(dims[i], dims[i + 1]); o += dims[i] * dims[i + 1]
            b = self.params[o:o + dims[i + 1]]; o += dims[i + 1]
            out[f"Dense_{i}"] = {"kernel": k, "bias": b}
        return out

    def _bind(self):
        self.net.params = self.params.data_ptr()

    def get_control_efforts_with_additional_term(self, x):
        """(u, V, dV/dx) for state(s) x — the notebooks' get_soft_pd_control_with_additional_term, batched, on device."""
        torch = self.torch
        single = np.ndim(x) == 1
        xd = L.dev_f32(x, (-1, self.n))
        B = xd.shape[0]
        V = torch.empty(B, device="cuda", dtype=torch.float32)
        p = torch.empty((B, self.n), device="cuda", dtype=torch.float32)
        u = torch.empty((B, self.m), device="cuda", dtype=torch.float32)
        self._bind()
        L.check(L.lib().hjb_softpd_policy(self.dynamics.system_spec(), self.net, L.ptr(xd), B, L.ptr(V), L.ptr(p), L.ptr(u),
                                          L.ptr(self.workspace), L.stream_ptr()), "hjb_softpd_policy")
        return (u[0], V[0], p[0]) if single else (u, V, p)

    def get_control_efforts(self, x):
        u, _, _ = self.get_control_efforts_with_additional_term(x)
        return u.cpu().numpy().astype(np.float64)

    def loss_grad(self, xs, form: str = "hjb", regularization: float = 1.0):
        """(loss, mean residual, mean hinge) as 0-d device tensors and the gradient in ``self.grad`` (mean over the batch)."""
        xd = L.dev_f32(xs, (-1, self.n))
        B = xd.shape[0]
        self._bind()
        L.check(L.lib().hjb_softpd_loss_grad(self.dynamics.system_spec(), self.net, L.ptr(xd), B, LOSS_FORMS[form],
                                             float(regularization), L.ptr(self.grad), L.ptr(self.sums), L.ptr(self.workspace),
                                             L.stream_ptr()), "hjb_softpd_loss_grad")
        res, hinge = self.sums[0] / B, self.sums[1] / B
        return res + float(regularization) * hinge, res, hinge

    def params_update(self, xs, form: str = "hjb", regularization: float = 1.0):
        """One Adam step on the chosen loss (the notebooks' soft_pd_params_update / soft_pd_params_warmup_update)."""
        out = self.loss_grad(xs, form, regularization)
        self.opt.count += 1
        L.check(L.lib().hjb_adam(L.ptr(self.params), L.ptr(self.opt.mu), L.ptr(self.opt.nu), L.ptr(self.grad), self.P_count,
                                 self.lr, 0.9, 0.999, 1e-8, int(self.opt.count), L.stream_ptr()), "hjb_adam")
        return out

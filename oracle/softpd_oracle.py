"""TEST INFRASTRUCTURE — CPU oracle for the notebooks' soft-PD baseline.

The reference's notebooks compare the positive-definite value net of controller/vhjb.py with an unconstrained one,
``SoftPDValueApproximator`` (examples/cartpole_balancing.ipynb cell 6, examples/drone_hovering.ipynb cell 6):

    z = wrap(x - xf);  V = Dense(1)(s(Dense(64)(s(Dense(128)(s(Dense(128)(z)))))))     — every Dense WITH bias, s = tanh / relu

trained on  mean_i [ res_i + reg * max(0, V(xf) - V(x_i)) ]  with

    res = |vdot + l(x, u)|                      (cartpole_balancing.ipynb cell 11: "unnormalized ... work better")
    res = |vdot / (l(x, u) + 1e-10) + 1|        (drone_hovering.ipynb cell 11)
    u   = clip(-R^-1 g^T dV/dx / 2 + uf)        or, in the drone notebook's warm-up, the LQR's  clip(-K z + uf)

and a warm-up on  mean_i |V(x_i) - z_i^T P z_i|  (cartpole_balancing.ipynb cell 11, ``soft_pd_warmup_hjb_loss``).

A PyTorch float64 autograd restatement (JAX / Flax are absent, as for oracle/vhjb_oracle.py: PARITY UNPINNED w.r.t. real JAX
numbers; pinned by autograd-vs-finite-difference checks in tests/test_softpd_oracle.py and, distributionally, by the
notebooks' printed costs).  Only tests/ and bench.py's checker legs import it.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from oracle.rollout_oracle import OracleSystem
from oracle.vhjb_oracle import lecun_normal

DT = torch.float64
FEATURES = (128, 128, 64, 1)


def init_params(n: int, seed: int = 0) -> List[np.ndarray]:
    """[W1, b1, W2, b2, W3, b3, W4, b4] — Flax Dense defaults: lecun-normal kernels (in, out), zero biases."""
    rng = np.random.default_rng(seed)
    dims = [n, *FEATURES]
    out = []
    for i in range(4):
        out += [lecun_normal(rng, dims[i], dims[i + 1]), np.zeros(dims[i + 1])]
    return out


def flat(params: Sequence[np.ndarray]) -> np.ndarray:
    return np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1) for p in params])


@dataclass
class SoftPDProblem:
    sys: OracleSystem
    Q: np.ndarray
    R: np.ndarray
    xf: np.ndarray
    uf: np.ndarray
    act: str = "tanh"            # cartpole notebook: tanh; drone notebook: relu
    residual: str = "plain"      # "plain": |vdot + l| ; "normalized": |vdot / (l + eps) + 1|
    eps: float = 1e-10


def _t(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=DT)


class SoftPDOracle:
    def __init__(self, prob: SoftPDProblem, params: Sequence[np.ndarray]):
        self.p = prob
        self.theta = [_t(w).clone().requires_grad_(True) for w in params]
        self.act = torch.tanh if prob.act == "tanh" else torch.relu
        self.Rinv = torch.linalg.inv(_t(prob.R))

    def _wrap(self, x):
        z = x - _t(self.p.xf)
        idx = self.p.sys.wrap_index()
        cols = [torch.remainder(z[:, i] + np.pi, 2 * np.pi) - np.pi if i in idx else z[:, i] for i in range(z.shape[1])]
        return torch.stack(cols, dim=1)

    def value(self, x: torch.Tensor):
        h = self._wrap(x)
        for k in range(4):
            h = h @ self.theta[2 * k] + self.theta[2 * k + 1]
            if k < 3:
                h = self.act(h)
        return h[:, 0]

    def pieces(self, xs: np.ndarray, create_graph: bool = False, K: Optional[np.ndarray] = None):
        """V, p = dV/dx, u (the net's control, or clip(-K z + uf) when K is given), xdot, vdot, l, z."""
        x = _t(xs).clone().requires_grad_(True)
        V = self.value(x)
        (p,) = torch.autograd.grad(V.sum(), x, create_graph=create_graph)
        f_np, g_np = self.p.sys.f_g(np.asarray(xs, dtype=np.float64))
        f, g = _t(f_np), _t(g_np)
        z = self._wrap(x.detach())
        umin, umax = _t(self.p.sys.umin), _t(self.p.sys.umax)
        if K is None:
            c = torch.einsum("bn,bnm->bm", p, g)
            u = torch.minimum(torch.maximum(-0.5 * c @ self.Rinv.T + _t(self.p.uf), umin), umax)
        else:
            u = torch.minimum(torch.maximum(-z @ _t(K).T + _t(self.p.uf), umin), umax)
        xdot = f + torch.einsum("bnm,bm->bn", g, u)
        vdot = (p * xdot).sum(dim=1)
        du = u - _t(self.p.uf)
        l = torch.einsum("bi,ij,bj->b", z, _t(self.p.Q), z) + torch.einsum("bi,ij,bj->b", du, _t(self.p.R), du)
        return {"V": V, "p": p, "u": u, "xdot": xdot, "vdot": vdot, "l": l, "z": z}

    def loss(self, xs: np.ndarray, form: str = "hjb", reg: float = 1.0, K=None, P=None):
        """form: "hjb" (residual + hinge, u from the net), "hjb_lqr" (the drone notebook's warm-up: the same with the LQR's u),
        "value_match" (the cart-pole notebook's warm-up: |V - z^T P z|).  Returns (loss, residual mean, hinge mean)."""
        q = self.pieces(xs, create_graph=True, K=K if form == "hjb_lqr" else None)
        if form == "value_match":
            tgt = torch.einsum("bi,ij,bj->b", q["z"], _t(P), q["z"])
            m = (q["V"] - tgt).abs().mean()
            return m, m, torch.zeros((), dtype=DT)
        if self.p.residual == "plain":
            res = (q["vdot"] + q["l"]).abs()
        else:
            res = (q["vdot"] / (q["l"] + self.p.eps) + 1.0).abs()
        v0 = self.value(_t(np.asarray(self.p.xf, dtype=np.float64)[None]))[0]
        hinge = torch.clamp(v0 - q["V"], min=0.0)
        return (res + reg * hinge).mean(), res.mean(), hinge.mean()

    def loss_and_grad(self, xs, form="hjb", reg=1.0, K=None, P=None):
        total, res, hinge = self.loss(xs, form, reg, K, P)
        grads = torch.autograd.grad(total, self.theta, allow_unused=True)
        grads = [torch.zeros_like(t) if g is None else g for g, t in zip(grads, self.theta)]
        return float(total), float(res), float(hinge), [g.detach().numpy() for g in grads]

"""TEST INFRASTRUCTURE — CPU oracle for the vhjb HJB-residual pass (hot path (b)).

A PyTorch float64 restatement of the arithmetic of the reference's ``controller/vhjb.py`` (which is
JAX/Flax/Optax and cannot run here: jax, flax and optax are absent and un-pinned — ``setup.py:6`` has
``install_requires=[]``).  Derivatives come from ``torch.autograd`` (``create_graph=True`` for the
input-gradient, then a second backward for the parameters), exactly the composition
``jax.value_and_grad(hjb_loss)`` of ``jax.grad(V)`` performs.

PARITY UNPINNED with respect to real JAX numbers: the reference holds no tests or golden vectors for this
path and its notebook loss printouts depend on JAX's PRNG (initial weights) and torch's shuffle order.  The
restatement is pinned instead by (i) a hand-derived closed-form reverse pass (``closed_form_grads``, SURVEY.md
§8a-V6) agreeing with autograd to ~1e-15 (tests/test_vhjb_oracle.py), (ii) the LQR fixed point: with
V = x^T P x the HJB residual vanishes (utils/debug_helper.py:78-102 ``check_hjb_condition_for_lqr``), and
(iii) the rollout half of ``get_control_efforts`` going through the reference-pinned dynamics oracle.  Bit-level
parity with JAX therefore stays unpinned; what real JAX output exists — the loss curves, rollout costs and closed-loop
costs printed in examples/double_integrator_optimal_time.ipynb (cells 11, 21) and examples/cartpole_balancing.ipynb
(cells 10, 16) — is reproduced statistically by the CUDA path, which this oracle checks sample by sample
(tests/test_training_quality_gpu.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / ``--impl reference``) may import
this module; the product package never does.  References are to /root/reference/controller/vhjb.py unless
another file is named.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from oracle.rollout_oracle import OracleSystem

DT = torch.float64

ACTS = {
    "relu": (torch.relu, ),
    "tanh": (torch.tanh, ),
    "sin": (torch.sin, ),
}


def lecun_normal(rng: np.random.Generator, fan_in: int, fan_out: int) -> np.ndarray:
    """Flax ``Dense`` default kernel init: truncated normal (+-2 sigma) with variance 1/fan_in after
    truncation (stddev = sqrt(1/fan_in) / 0.87962566...), kernel stored (in, out)."""
    std = np.sqrt(1.0 / fan_in) / 0.87962566103423978
    w = rng.normal(size=(fan_in, fan_out))
    bad = np.abs(w) > 2
    while bad.any():
        w[bad] = rng.normal(size=int(bad.sum()))
        bad = np.abs(w) > 2
    return w * std


def init_weights(n: int, features: Sequence[int] = (128, 128, 64), seed: int = 0) -> List[np.ndarray]:
    rng = np.random.default_rng(seed)
    dims = [n, *features]
    return [lecun_normal(rng, dims[i], dims[i + 1]) for i in range(len(features))]


@dataclass
class VhjbProblem:
    """Everything ``VHJBController`` reads from its dynamics and config (vhjb.py:77-128)."""
    sys: OracleSystem
    Q: np.ndarray
    R: np.ndarray
    xf: np.ndarray
    uf: np.ndarray
    mean: np.ndarray                     # normalization_mean (error coordinates)
    std: np.ndarray                      # normalization_std
    eps: float = 1e-10                   # config.epsilon
    eps_s: float = 1e-3                  # config.epsilon_scalar
    act: str = "relu"                    # vhjb.py: relu; notebooks: tanh (cartpole), sin (double integrator)
    control_form: str = "clip"           # "clip": :220   | "bang": u = -sign(p . g) (double_integrator nb cell 11:17)
    residual_form: str = "normalized"    # "normalized": :233,241 | "min_time": |vdot + l_i| with given l_i, plain mean
                                         # (double_integrator nb cell 11:20, l_i = 1[|x|^2 > 1e-4] from cell 7:4)


def set_dtype(dtype):
    """float64 (default, the checker) or float32 (CPU-baseline timing: the reference's JAX code is float32)."""
    global DT
    DT = dtype


def _t(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=DT)


class VhjbOracle:
    def __init__(self, prob: VhjbProblem, weights: Sequence[np.ndarray]):
        self.p = prob
        self.W = [_t(w).clone().requires_grad_(True) for w in weights]
        self.act = ACTS[prob.act][0]
        self.Rinv = torch.linalg.inv(_t(prob.R))

    # ---- value network (ValueFunctionApproximator.__call__, :29-60) --------------------------------
    def _wrap(self, x):
        z = x - _t(self.p.xf)
        cols = []
        idx = self.p.sys.wrap_index()
        for i in range(z.shape[1]):
            c = z[:, i]
            if i in idx:
                c = torch.remainder(c + np.pi, 2 * np.pi) - np.pi     # unit derivative w.r.t. x
            cols.append(c)
        return torch.stack(cols, dim=1)

    def value(self, x: torch.Tensor):
        z = self._wrap(x)                                    # :39
        e = self.p.eps_s * (z * z).sum(dim=1)                # :42
        h = (z - _t(self.p.mean)) / _t(self.p.std)           # :45
        for i, W in enumerate(self.W):                       # :47-55 (Dense, use_bias=False; y = x @ kernel)
            h = h @ W
            if i != len(self.W) - 1:
                h = self.act(h)
        return (h * h).sum(dim=1) + e                        # :58

    def value_and_gradient(self, x: torch.Tensor, create_graph: bool):
        x = x.clone().requires_grad_(True)
        V = self.value(x)
        (p,) = torch.autograd.grad(V.sum(), x, create_graph=create_graph)   # get_v_gradient, :201-202
        return V, p

    # ---- per-sample pieces --------------------------------------------------------------------------
    def pieces(self, xs: np.ndarray, create_graph: bool = False, running: Optional[np.ndarray] = None):
        """V, p = dV/dx, u*, xdot, vdot, l, residual argument — rows V1..V4 of SURVEY.md §8a."""
        x = _t(xs)
        f_np, g_np = self.p.sys.f_g(np.asarray(xs, dtype=np.float64))       # get_control_affine_matrix (:219, :229)
        f, g = _t(f_np), _t(g_np)
        V, p = self.value_and_gradient(x, create_graph)
        c = torch.einsum("bn,bnm->bm", p, g)                                 # g^T p
        if self.p.control_form == "clip":
            u_raw = -0.5 * c @ self.Rinv.T + _t(self.p.uf)                   # :220
            u = torch.minimum(torch.maximum(u_raw, _t(self.p.sys.umin)), _t(self.p.sys.umax))
        else:
            u = -torch.sign(c)
        xdot = f + torch.einsum("bnm,bm->bn", g, u)                          # :230
        vdot = (p * xdot).sum(dim=1)                                         # :231
        if self.p.residual_form == "normalized":
            z = self._wrap(x.detach())
            du = u - _t(self.p.uf)
            l = torch.einsum("bi,ij,bj->b", z, _t(self.p.Q), z) + torch.einsum("bi,ij,bj->b", du, _t(self.p.R), du)  # :162-165
            r = vdot / (l + self.p.eps) + 1.0                                # :232
        else:
            l = _t(running)
            r = vdot + l
        return dict(V=V, p=p, u=u, xdot=xdot, vdot=vdot, l=l, r=r)

    # ---- losses (:227-253) --------------------------------------------------------------------------
    def losses(self, xs, dones, costs, create_graph: bool = False):
        d = _t(dones)
        if self.p.residual_form == "normalized":
            q = self.pieces(xs, create_graph)
            hjb = (q["r"].abs() * (1 - d)).sum() / ((1 - d).sum() + self.p.eps)          # :233, :241
            term = ((q["V"] / (_t(costs) + self.p.eps) - 1).abs() * d).sum() / (d.sum() + self.p.eps)   # :247-253
        else:
            q = self.pieces(xs, create_graph, running=costs)
            hjb = q["r"].abs().mean()
            term = torch.zeros((), dtype=DT)
        return hjb, term, q

    def loss_and_grad(self, xs, dones, costs, reg: float):
        """params_update's value_and_grad part (:282-285): total = hjb + reg * term, grad likewise."""
        for W in self.W:
            W.grad = None
        hjb, term, q = self.losses(xs, dones, costs, create_graph=True)
        total = hjb + reg * term
        grads = torch.autograd.grad(total, self.W, allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(W) for g, W in zip(grads, self.W)]
        return float(total), float(hjb), float(term), [g.detach().numpy() for g in grads], q

    # ---- hand-derived reverse pass, used to cross-check autograd -------------------
    def closed_form_grads(self, xs, dones, costs, reg: float):
        p_ = self.p
        assert p_.residual_form == "normalized" and p_.control_form == "clip" and len(self.W) == 3
        x = _t(xs)
        W1, W2, W3 = [w.detach() for w in self.W]
        sig, mu = _t(p_.std), _t(p_.mean)
        act = p_.act
        d1 = {"relu": lambda a: (a > 0).to(DT), "tanh": lambda a: 1 - torch.tanh(a) ** 2, "sin": torch.cos}[act]
        d2 = {"relu": lambda a: torch.zeros_like(a), "tanh": lambda a: -2 * torch.tanh(a) * (1 - torch.tanh(a) ** 2),
              "sin": lambda a: -torch.sin(a)}[act]
        z = self._wrap(x)
        h0 = (z - mu) / sig
        a1 = h0 @ W1; h1 = self.act(a1)
        a2 = h1 @ W2; h2 = self.act(a2)
        y = h2 @ W3
        V = (y * y).sum(1) + p_.eps_s * (z * z).sum(1)
        gy = 2 * y
        b2 = gy @ W3.T; g2 = b2 * d1(a2)
        b1 = g2 @ W2.T; g1 = b1 * d1(a1)
        g0 = g1 @ W1.T
        pgrad = g0 / sig + 2 * p_.eps_s * z
        f_np, g_np = p_.sys.f_g(np.asarray(xs, dtype=np.float64))
        f, G = _t(f_np), _t(g_np)
        c = torch.einsum("bn,bnm->bm", pgrad, G)
        u_raw = -0.5 * c @ self.Rinv.T + _t(p_.uf)
        lo, hi = _t(p_.sys.umin), _t(p_.sys.umax)
        u = torch.minimum(torch.maximum(u_raw, lo), hi)
        inside = ((u_raw > lo) & (u_raw < hi)).to(DT)
        xdot = f + torch.einsum("bnm,bm->bn", G, u)
        vdot = (pgrad * xdot).sum(1)
        du = u - _t(p_.uf)
        Q, R = _t(p_.Q), _t(p_.R)
        l = torch.einsum("bi,ij,bj->b", z, Q, z) + torch.einsum("bi,ij,bj->b", du, R, du)
        r = vdot / (l + p_.eps) + 1
        dn = _t(dones)
        w = (1 - dn) / ((1 - dn).sum() + p_.eps)
        rbar = w * torch.sign(r)
        vbar = rbar / (l + p_.eps)
        lbar = -rbar * vdot / (l + p_.eps) ** 2
        pbar = vbar[:, None] * xdot
        ubar = vbar[:, None] * c + lbar[:, None] * (du @ (R + R.T))
        uraw_bar = ubar * inside
        pbar = pbar + torch.einsum("bm,bnm->bn", -0.5 * uraw_bar @ self.Rinv, G)
        cst = _t(costs)
        tq = V / (cst + p_.eps) - 1
        Vbar = reg * (dn / (dn.sum() + p_.eps)) * torch.sign(tq) / (cst + p_.eps)
        g0bar = pbar / sig
        dW1 = g0bar.T @ g1
        g1bar = g0bar @ W1; b1bar = g1bar * d1(a1)
        dW2 = b1bar.T @ g2
        g2bar = b1bar @ W2; b2bar = g2bar * d1(a2)
        dW3 = b2bar.T @ gy
        gybar = b2bar @ W3
        ybar = 2 * gybar + 2 * y * Vbar[:, None]
        dW3 = dW3 + h2.T @ ybar
        a2bar = (ybar @ W3.T) * d1(a2) + g2bar * b2 * d2(a2)
        dW2 = dW2 + h1.T @ a2bar
        a1bar = (a2bar @ W2.T) * d1(a1) + g1bar * b1 * d2(a1)
        dW1 = dW1 + h0.T @ a1bar
        return [dW1.numpy(), dW2.numpy(), dW3.numpy()]


# ----------------------------------------------------------------------------------------------------
# optimiser and schedule (optax defaults, vhjb.py:120-128, :286-287)
# ----------------------------------------------------------------------------------------------------
def adam_step(w, m, v, g, step: int, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam (scale_by_adam then scale by -lr): m,v EMA; bias-corrected; w -= lr * mhat / (sqrt(vhat) + eps).
    ``step`` is the 1-based count of this update."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mhat = m / (1 - b1 ** step)
    vhat = v / (1 - b2 ** step)
    return w - lr * mhat / (np.sqrt(vhat) + eps), m, v


def sgdr_schedule(step: int, init=0.0, peak=1e-5, end=0.0, cycles=10, warmup=1000, total=2000) -> float:
    """optax.sgdr_schedule of ``cycles`` identical warmup_cosine_decay_schedule cycles (vhjb.py:123-126):
    linear init -> peak over ``warmup`` steps, then cosine peak -> end until ``total``; after the last cycle the
    value stays at ``end``."""
    cycle = step // total
    if cycle >= cycles:
        # optax.join_schedules keeps evaluating the LAST schedule with a growing step -> it stays at its end value
        return float(end)
    s = step - cycle * total
    if s < warmup:
        return float(init + (peak - init) * s / warmup)
    frac = min((s - warmup) / max(1, total - warmup), 1.0)
    return float(end + (peak - end) * 0.5 * (1 + np.cos(np.pi * frac)))

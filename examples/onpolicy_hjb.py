"""On-policy HJB value learning on the GPU — the common loop of the reference's cart-pole, drone and 10-D quadcopter
notebooks ("ours" runs: examples/cartpole_balancing.ipynb cells 9-10, examples/10D_quadcopte.ipynb cells 9-10): every epoch
a batch of closed-loop trajectories under the CURRENT value-net policy is added to the data set (the states visited while
inside the observation box), then one shuffled pass of minibatches trains V on the normalised HJB residual
|dV/dx . (f + g u) / (l(x, u) + eps) + 1|.  All trajectories of an epoch advance together: one fused value-net launch
(hjb_vhjb_residual -> u) and one dynamics launch (hjb_dynamics) per step; the data set is a ring in HBM and every update is
one fused loss + gradient launch followed by Adam."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np


@dataclass
class Problem:
    dyn: object                      # q_learning_with_hjb_b200 Dynamics
    xf: np.ndarray
    uf: np.ndarray
    obs_min: np.ndarray
    obs_max: np.ndarray
    act: str = "tanh"
    rollout_steps: int = 200
    trajectories_per_epoch: int = 20
    batch: int = 256
    far_away: Optional[np.ndarray] = None     # stop a rollout once |wrap(x - xf)| exceeds this (10-D notebook cell 9)

    def kernels(self):
        from q_learning_with_hjb_b200.controller.vhjb import VhjbKernels
        n, m = self.dyn.get_dimension()
        return VhjbKernels(self.dyn, self.xf, self.uf, np.eye(n), np.eye(m), np.zeros(n), np.ones(n), 1e-10, 1e-3, act=self.act)


class Policy:
    """u(x) of the value net, batched on the device."""

    def __init__(self, k, params):
        import torch
        self.k, self.params, self.torch = k, params, torch

    def __call__(self, x):
        z = self.torch.zeros(x.shape[0], device="cuda")
        out, _ = self.k.residual(self.params, x, z, z, want=("u",))
        return out["u"]


def error_coords(p: Problem, x):
    """z = wrap(x - xf) on the device."""
    import torch
    z = x - torch.as_tensor(p.xf, dtype=torch.float32, device="cuda")
    for i in p.dyn.WRAP_INDEX:
        z[:, i] = torch.remainder(z[:, i] + np.pi, 2 * np.pi) - np.pi
    return z


def running_cost(p: Problem, x, u):
    """l(x, u) = z^T z + (u - uf)^T (u - uf)  (Q = I, R = I in every notebook)."""
    import torch
    z = error_coords(p, x)
    du = u - torch.as_tensor(p.uf, dtype=torch.float32, device="cuda")
    return (z * z).sum(1) + (du * du).sum(1)


def rollout(p: Problem, policy, x0):
    """The notebooks' rollout_trajectory for all rows of x0 at once: a state is collected while its trajectory has not
    left the observation box; cumulated_cost += dt l(x_next, u) every step (until the far-away stop, if any).
    Returns (states [K, n] trajectory by trajectory, costs [N], lengths [N])."""
    import torch
    lo = torch.as_tensor(p.obs_min, dtype=torch.float32, device="cuda")
    hi = torch.as_tensor(p.obs_max, dtype=torch.float32, device="cuda")
    far = None if p.far_away is None else torch.as_tensor(p.far_away, dtype=torch.float32, device="cuda")
    x = torch.as_tensor(np.asarray(x0, dtype=np.float32)).cuda()
    N = x.shape[0]
    within = torch.ones(N, dtype=torch.bool, device="cuda")
    running = torch.ones(N, dtype=torch.bool, device="cuda")
    cost = torch.zeros(N, device="cuda")
    kept, masks = [], []
    for _ in range(p.rollout_steps):
        z = error_coords(p, x)
        within = within & ~(((z > hi) | (z < lo)).any(dim=1))
        kept.append(x)
        masks.append(within & running)
        u = policy(x)
        xn = p.dyn.simulate(x, u)
        cost = cost + torch.where(running, p.dyn.dt * running_cost(p, xn, u), torch.zeros_like(cost))
        x = torch.where(running[:, None], xn, x)
        if far is not None:
            running = running & ~((error_coords(p, x).abs() > far).any(dim=1))
    states = torch.stack(kept, dim=1)
    mask = torch.stack(masks, dim=1)
    return states[mask], cost, mask.sum(1)


def closed_loop_cost(p: Problem, policy, x0, steps):
    """The notebooks' test_learned_policy: sum_t l(x_t, u_t) dt."""
    import torch
    x = torch.as_tensor(np.asarray(x0, dtype=np.float32)).cuda()
    cost = torch.zeros(x.shape[0], device="cuda")
    for _ in range(steps):
        u = policy(x)
        cost = cost + p.dyn.dt * running_cost(p, x, u)
        x = p.dyn.simulate(x, u)
    return cost.cpu().numpy().astype(np.float64)


def train(p: Problem, k, epochs, seed=0, log=print, log_every=10):
    """-> (flat device parameters, [(mean loss, mean rollout cost, mean collected length)] per epoch)."""
    import torch
    from q_learning_with_hjb_b200.controller.vhjb import AdamState, DeviceReplayBuffer, FEATURES, lecun_normal
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    n = p.dyn.get_dimension()[0]
    dims = [n, *FEATURES]
    params = torch.as_tensor(np.concatenate([lecun_normal(rng, dims[i], dims[i + 1]).reshape(-1) for i in range(3)])).cuda()
    opt = AdamState(0, torch.zeros_like(params), torch.zeros_like(params))
    policy = Policy(k, params)
    data = DeviceReplayBuffer(n, p.batch + epochs * p.trajectories_per_epoch * p.rollout_steps)
    data.extend(np.tile(np.asarray(p.xf, dtype=np.float32), (p.batch, 1)), np.ones(p.batch), np.zeros(p.batch))  # [xf] * 256
    history = []
    for epoch in range(epochs):
        x0 = np.stack([p.dyn.get_initial_state() for _ in range(p.trajectories_per_epoch)])
        states, costs, lengths = rollout(p, policy, x0)
        if states.shape[0]:
            data.extend(states, torch.ones(states.shape[0], device="cuda"), torch.zeros(states.shape[0], device="cuda"))
        total, nb = torch.zeros((), device="cuda"), 0
        for xs, cs, ds in data.batches(p.batch):
            sums, norm = k.train_step(params, opt, xs, ds, cs, 0.0, 1e-3)
            total += sums[0] / norm[0]
            nb += 1
        history.append((float(total) / nb, float(costs.mean()), float(lengths.float().mean())))
        if log and (epoch + 1) % log_every == 0:
            log(f"epoch:{epoch + 1} loss:{history[-1][0]:.5f}, cumulated cost:{history[-1][1]:.3f}, "
                f"avg trajectory length: {history[-1][2]:.2f}")
    return params, history


def soft_pd_policy(ctl):
    """u(x) of a SoftPDController, batched on the device."""
    return lambda x: ctl.get_control_efforts_with_additional_term(x)[0]


def train_soft_pd(p: Problem, ctl, epochs, warmup_epochs=20, warmup_form="value_match", regularization=1.0, seed=0,
                  log=print, log_every=10):
    """The notebooks' soft-PD runs (examples/cartpole_balancing.ipynb cell 11, examples/drone_hovering.ipynb cell 11; cell 12
    of both is the same with ``warmup_epochs=0``): the SAME on-policy loop as :func:`train` around the unconstrained value net
    of ``SoftPDController`` — ``warmup_epochs`` epochs on the warm-up loss (``"value_match"``: |V - z^T P z|, cart-pole;
    ``"hjb_lqr"``: the HJB residual under the LQR's control, drone), then the HJB residual + regularization * hinge.
    -> [(mean loss, mean rollout cost, mean collected length)] per epoch."""
    import torch
    from q_learning_with_hjb_b200.controller.vhjb import DeviceReplayBuffer
    torch.manual_seed(seed)
    n = p.dyn.get_dimension()[0]
    policy = soft_pd_policy(ctl)
    data = DeviceReplayBuffer(n, p.batch + epochs * p.trajectories_per_epoch * p.rollout_steps)
    data.extend(np.tile(np.asarray(p.xf, dtype=np.float32), (p.batch, 1)), np.ones(p.batch), np.zeros(p.batch))  # [xf] * 256
    history = []
    for epoch in range(epochs):
        x0 = np.stack([p.dyn.get_initial_state() for _ in range(p.trajectories_per_epoch)])
        states, costs, lengths = rollout(p, policy, x0)
        if states.shape[0]:
            data.extend(states, torch.ones(states.shape[0], device="cuda"), torch.zeros(states.shape[0], device="cuda"))
        form = warmup_form if epoch < warmup_epochs else "hjb"
        total, nb = torch.zeros((), device="cuda"), 0
        for xs, _, _ in data.batches(p.batch):
            loss, _, _ = ctl.params_update(xs, form, regularization)
            total += loss
            nb += 1
        history.append((float(total) / nb, float(costs.mean()), float(lengths.float().mean())))
        if log and (epoch + 1) % log_every == 0:
            log(f"{'warmup ' if epoch < warmup_epochs else ''}epoch:{epoch + 1} loss:{history[-1][0]:.5f}, "
                f"cumulated cost:{history[-1][1]:.3f}, avg trajectory length: {history[-1][2]:.2f}")
    return history


def lqr_policy(p: Problem, K, clip=False):
    """u = -K wrap(x - xf) + uf (optionally clipped; Dynamics.simulate clips anyway)."""
    import torch
    Kt = torch.as_tensor(np.asarray(K), dtype=torch.float32, device="cuda")
    uf = torch.as_tensor(p.uf, dtype=torch.float32, device="cuda")
    lo = torch.as_tensor(np.asarray(p.dyn.umin), dtype=torch.float32, device="cuda")
    hi = torch.as_tensor(np.asarray(p.dyn.umax), dtype=torch.float32, device="cuda")

    def policy(x):
        u = -(error_coords(p, x) @ Kt.T) + uf
        return torch.minimum(torch.maximum(u, lo), hi) if clip else u
    return policy

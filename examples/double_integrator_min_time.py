"""Minimum-time value learning for the double integrator, end to end on the GPU — the experiment of the reference's
examples/double_integrator_optimal_time.ipynb (cells 4-11, 21): a sin value net V(x) = |MLP(x)|^2 + 1e-3 |x|^2 trained on
the HJB residual |dV/dx . (A x + B u) + l(x)|, u = -sign(dV/dx . B), l = 1[|x|^2 > 1e-4], over 65,536 states sampled in
[-1, 1]^2 (Adam 1e-3, shuffled minibatches of 256, 100 epochs), then the time the learned bang-bang policy needs to reach
|x|^2 <= 1e-4 under the exact zero-order-hold step — against the analytic optimum and the saturated LQR.

The notebook reports (cell 21 output): learned 2.61 +- 0.89 s, LQR 4.10 +- 1.28 s, level-set solver's policy
1.62 +- 0.60 s, analytic 1.57 +- 0.54 s.  The three model-based policies run inside the rollout kernel
(controller/lqr.py, controller/min_time.py); the learned policy is evaluated per step through the residual kernel.

    python examples/double_integrator_min_time.py [--epochs 100] [--batch 256]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

METRIC = 1e-4
DT = 0.01


def make_problem():
    from q_learning_with_hjb_b200.configs.dynamics.dynamics_config import LinearDynamicsConfig
    from q_learning_with_hjb_b200.controller.vhjb import VhjbKernels
    from q_learning_with_hjb_b200.dynamics.linear import LinearDynamics
    cfg = LinearDynamicsConfig(seed=0, x0_mean=np.zeros(2), x0_std=np.ones(2), dt=DT, umin=np.array([-1.0]), umax=np.array([1.0]),
                               A=np.array([[0.0, 1.0], [0.0, 0.0]]), B=np.array([[0.0], [1.0]]))
    dyn = LinearDynamics(cfg)
    k = VhjbKernels(dyn, np.zeros(2), np.zeros(1), np.eye(2), np.eye(1), np.zeros(2), np.ones(2), 1e-10, 1e-3, act="sin",
                    control_form="bangbang", residual_form="min_time")
    return dyn, k


def analytic_control(x):
    """Time-optimal switching curve (notebook cell 18)."""
    p, v = x[:, 0], x[:, 1]
    plus = ((v < 0) & (p <= 0.5 * v * v)) | ((v >= 0) & (p < -0.5 * v * v))
    u = np.where(plus, 1.0, -1.0)
    return np.where((x * x).sum(1) <= METRIC, 0.0, u)[:, None]


def time_to_origin(control, x0, max_T=15.0):
    """Per-trajectory first time with |x|^2 <= METRIC under the exact ZOH step (notebook cells 4, 9); max_T if never."""
    Ad = np.array([[1.0, DT], [0.0, 1.0]])
    Bd = np.array([[0.5 * DT * DT], [DT]])
    x = np.array(x0, dtype=np.float64)
    t_hit = np.full(len(x), max_T)
    alive = np.ones(len(x), dtype=bool)
    for i in range(int(round(max_T / DT))):
        hit = alive & ((x * x).sum(1) <= METRIC)
        t_hit[hit] = i * DT
        alive &= ~hit
        if not alive.any():
            break
        u = control(x)
        x = x @ Ad.T + u @ Bd.T
    return t_hit


def train(k, epochs=100, batch=256, n_states=1 << 16, seed=0, log=print):
    import torch
    from q_learning_with_hjb_b200.controller.vhjb import AdamState, FEATURES, lecun_normal
    rng = np.random.default_rng(seed)
    dims = [2, *FEATURES]
    params = torch.as_tensor(np.concatenate([lecun_normal(rng, dims[i], dims[i + 1]).reshape(-1) for i in range(3)])).cuda()
    opt = AdamState(0, torch.zeros_like(params), torch.zeros_like(params))
    xs = torch.as_tensor(rng.uniform(-1, 1, size=(n_states, 2)).astype(np.float32)).cuda()
    run_cost = ((xs * xs).sum(1) > METRIC).float()
    dones = torch.zeros(n_states, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(seed)
    losses = []
    for epoch in range(epochs):
        perm = torch.randperm(n_states, device="cuda", generator=g)
        total = torch.zeros((), device="cuda")
        nb = n_states // batch
        for b in range(nb):
            idx = perm[b * batch:(b + 1) * batch]
            sums, norm = k.train_step(params, opt, xs[idx].contiguous(), dones[idx].contiguous(), run_cost[idx].contiguous(),
                                      0.0, 1e-3)
            total += sums[0] / norm[0]
        losses.append(float(total) / nb)
        if log and (epoch + 1) % 10 == 0:
            log(f"epoch:{epoch + 1}, loss:{losses[-1]:.5f}")
    return params, losses


def learned_control(k, params):
    import torch

    def control(x):
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()
        z = torch.zeros(len(x), device="cuda")
        out, _ = k.residual(params, xd, z, z, want=("u",))
        return out["u"].cpu().numpy().astype(np.float64)
    return control


def lqr_control():
    """Host-side saturated LQR with R = 1 (a cross-check for the tests; ``device_times`` runs the notebook's R = 0.01)."""
    import scipy.linalg
    A, B = np.array([[0.0, 1.0], [0.0, 0.0]]), np.array([[0.0], [1.0]])
    P = scipy.linalg.solve_continuous_are(A, B, np.eye(2), np.eye(1))
    K = B.T @ P
    return lambda x: np.clip(-x @ K.T, -1.0, 1.0)


def device_times(dyn, x0, T=15.0):
    """Saturated LQR, the analytic optimum and (when the level-set data is at hand) the level-set solver's policy, every
    trajectory stepped by the rollout kernel with the exact zero-order-hold update (notebook cells 18-21)."""
    import scipy.linalg
    from q_learning_with_hjb_b200.controller.lqr import StateFeedback
    from q_learning_with_hjb_b200.controller.min_time import GridPolicyController, SwitchingCurveController, time_to_goal
    A, B = np.array([[0.0, 1.0], [0.0, 0.0]]), np.array([[0.0], [1.0]])
    R = np.array([[0.01]])                                    # cell 4: the LQR the notebook compares with
    P = scipy.linalg.solve_continuous_are(A, B, np.eye(2), R)
    ctls = [("saturated LQR", StateFeedback(dyn, np.linalg.inv(R) @ B.T @ P, clip=True)),
            ("analytic optimum", SwitchingCurveController(dyn, METRIC))]
    fixture = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                           "double_integrator_level_set.npz")
    if os.path.exists(fixture):
        d = np.load(fixture)
        ctls.append(("level-set solver's policy", GridPolicyController.from_value_function(dyn, d["value_level_set"], d["pos"],
                                                                                            d["vel"])))
    steps = int(round(T / DT))
    out = []
    for name, ctl in ctls:
        res = dyn.rollout(ctl, np.asarray(x0, dtype=np.float32), steps, integrator="discrete", record_stride=1,
                          record_controls=False)
        out.append((name, time_to_goal(res, DT, METRIC, t_max=T)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--trajectories", type=int, default=200)
    args = ap.parse_args()
    dyn, k = make_problem()
    t0 = time.time()
    params, losses = train(k, args.epochs, args.batch)
    print(f"trained {args.epochs} epochs in {time.time() - t0:.1f} s, final loss {losses[-1]:.5f}")
    x0 = np.random.default_rng(1).uniform(-1, 1, size=(args.trajectories, 2))
    t = time_to_origin(learned_control(k, params), x0)
    print(f"time to origin, learned (sin net, tcgen05 kernels): {t.mean():.3f} +- {t.std():.3f} s")
    for name, t in device_times(dyn, x0):
        print(f"time to origin, {name}: {t.mean():.3f} +- {t.std():.3f} s")


if __name__ == "__main__":
    main()

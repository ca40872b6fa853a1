"""The named vhjb workloads of SURVEY.md section 8d as product-side objects: dynamics from the package's own gin files,
``VhjbKernels`` for the task, Flax-default random weights and a synthetic batch.  ``bench.py`` builds what it times from
here (no test infrastructure on the measured path); the tests build the same problems from the oracle's descriptions
(tests/helpers_vhjb.py) and `tests/test_host_logic.py` checks that the two agree.

    linear / cartpole / quad2d    the reference's gin configurations (configs/dynamics/*.gin + *_vhjb_controller.gin)
    quad10d                       C5: examples/10D_quadcopte.ipynb cells 4, 6, 9 (relu net, hover thrust as uf)
    di_mintime                    C2: examples/double_integrator_optimal_time.ipynb cells 5, 7, 11 (sin net, bang-bang control,
                                  minimum-time residual |vdot + l|)
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))


@dataclass
class VhjbWorkload:
    name: str
    dynamics: str                       # which Dynamics class / gin file
    xf: np.ndarray
    uf: np.ndarray
    obs: Sequence[float]                # half-widths of the sampling box about xf
    act: str = "relu"
    control_form: str = "clipped"       # "clipped": u = clip(uf - R^-1 g^T p / 2);  "bangbang": u = -sign(g^T p)
    residual_form: str = "normalized"   # "normalized": |vdot / (l + eps) + 1|;  "min_time": |vdot + l_i|
    eps: float = 1e-10
    eps_s: float = 1e-3


def make_dynamics(kind: str):
    """The package's Dynamics object of a kind, configured by the package's copy of the reference's gin file."""
    from q_learning_with_hjb_b200.configs import gin_compat as gin
    from q_learning_with_hjb_b200.configs.dynamics import dynamics_config as DC
    cfg = os.path.join(PKG, "configs", "dynamics")
    if kind == "linear":
        from q_learning_with_hjb_b200.dynamics.linear import LinearDynamics
        gin.parse_config_file(os.path.join(cfg, "linear.gin"))
        return LinearDynamics(DC.LinearDynamicsConfig())
    if kind == "cartpole":
        from q_learning_with_hjb_b200.dynamics.cartpole import Cartpole
        gin.parse_config_file(os.path.join(cfg, "cartpole.gin"))
        return Cartpole(DC.CartpoleDynamicsConfig())
    if kind == "acrobot":
        from q_learning_with_hjb_b200.dynamics.acrobot import Acrobot
        return Acrobot()
    if kind == "quad2d":
        from q_learning_with_hjb_b200.dynamics.quadrotors import Quadrotors2D
        gin.parse_config_file(os.path.join(cfg, "quadrotors2D.gin"))
        return Quadrotors2D(DC.Quadrotors2DConfig())
    if kind == "quad10d":
        from q_learning_with_hjb_b200.dynamics.quadrotors import NearHoverQuadcopter
        gin.parse_config_file(os.path.join(cfg, "near_hover_quadcopter.gin"))
        return NearHoverQuadcopter(DC.NearHoverQuadcopterConfig())
    raise ValueError(kind)


def make_controller(kind: str, dyn):
    """The reference's model-based controllers as the rollout workloads use them."""
    if kind == "lqr":
        from q_learning_with_hjb_b200.controller.lqr import LQR
        return LQR(dyn, np.eye(2), np.eye(1))
    if kind == "cartpole_lqr":   # the cart-pole notebook's inline LQR about xf = [0, 3.1415926, 0, 0], unclipped (cell 4)
        from q_learning_with_hjb_b200.controller.lqr import StateFeedback
        from q_learning_with_hjb_b200.controller.controller_basic import lqr_gain
        xf = np.array([0, 3.1415926, 0, 0])
        Minv = np.linalg.inv(dyn.get_M(xf))
        A = np.zeros((4, 4)); A[0, 2] = A[1, 3] = 1
        A[2:, :2] = -Minv @ np.array([[0, 0], [0, -dyn.mp * dyn.g * dyn.l]])
        B = np.concatenate([np.zeros(2), Minv @ dyn.get_B()]).reshape(4, 1)
        K, _ = lqr_gain(A, B, np.eye(4), np.eye(1))
        return StateFeedback(dyn, K, xf=xf, uf=np.zeros(1), clip=False)
    if kind == "cartpole_es":
        from q_learning_with_hjb_b200.controller.cartpole_energy_shaping import CartpoleEnergyShapingController
        return CartpoleEnergyShapingController(dyn)
    if kind == "acrobot_es":
        from q_learning_with_hjb_b200.controller.acrobot_energy_shaping import AcrobotEnergyShapingController
        return AcrobotEnergyShapingController(dyn)
    from q_learning_with_hjb_b200.controller import quadrotors_model_based_controller as QC
    if kind == "quad2d_hover":
        return QC.Quadrotors2DHoveringController(dyn, np.zeros(6), np.eye(6), np.eye(2))
    if kind == "quad10d_hover":
        return QC.NearHoverQuadcopterHoveringController(dyn, np.zeros(10), np.eye(10), np.eye(3))
    raise ValueError(kind)


def vhjb_workload(name: str) -> VhjbWorkload:
    if name == "linear":
        return VhjbWorkload(name, "linear", np.zeros(2), np.zeros(1), [2, 3])
    if name == "cartpole":
        return VhjbWorkload(name, "cartpole", np.array([0, 3.1415926, 0, 0]), np.zeros(1), [4.8, 0.418, 4, 4])
    if name == "quad2d":
        return VhjbWorkload(name, "quad2d", np.zeros(6), np.array([4.905, 4.905]), [2, 2, 1.5, 5, 5, 2])
    if name == "quad10d":
        dyn = make_dynamics("quad10d")
        uf = np.array([dyn.g * dyn.m / dyn.kT, 0.0, 0.0])
        return VhjbWorkload(name, "quad10d", np.zeros(10), uf, [2, 2, 2, .5, .5, 4, 4, 4, 2, 2])
    if name == "di_mintime":
        return VhjbWorkload(name, "linear", np.zeros(2), np.zeros(1), [1, 1], act="sin", control_form="bangbang",
                            residual_form="min_time")
    raise ValueError(name)


def make_vhjb_kernels(name: str):
    """(VhjbKernels, VhjbWorkload) of a named workload."""
    from q_learning_with_hjb_b200.controller.vhjb import VhjbKernels
    w = vhjb_workload(name)
    dyn = make_dynamics(w.dynamics)
    n, m = dyn.get_dimension()
    if name == "di_mintime":            # the notebook's double integrator: dt = 0.01, |u| <= 1
        dyn.dt = 0.01
        dyn.umin, dyn.umax = np.float32([-1]), np.float32([1])
    k = VhjbKernels(dyn, w.xf, w.uf, np.eye(n), np.eye(m), np.zeros(n), np.ones(n), w.eps, w.eps_s, act=w.act,
                    control_form=w.control_form, residual_form=w.residual_form)
    return k, w


def init_weights(n: int, features: Sequence[int] = (128, 128, 64), seed: int = 0) -> List[np.ndarray]:
    """Flax ``Dense`` default kernels (lecun-normal, (in, out)) of the bias-free value net (controller/vhjb.py:17-60)."""
    from q_learning_with_hjb_b200.controller.vhjb import lecun_normal
    rng = np.random.default_rng(seed)
    dims = [n, *features]
    return [lecun_normal(rng, dims[i], dims[i + 1]) for i in range(len(features))]


def flat_params(weights) -> np.ndarray:
    return np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in weights])


def sample_vhjb_batch(name: str, B: int, seed: int = 0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """States U(-obs, obs) about xf, dones ~ Bernoulli(0.1), terminal costs ~ U(0.1, 10); for the
    minimum-time form dones = 0 and the per-sample running cost l_i = 1[|x|^2 > 1e-4] (notebook cell 7:4)."""
    w = vhjb_workload(name)
    rng = np.random.default_rng(seed)
    xs = rng.uniform(-1, 1, size=(B, len(w.xf))) * np.asarray(w.obs) + w.xf
    if w.residual_form == "min_time":
        dones = np.zeros(B)
        costs = ((xs ** 2).sum(1) > 1e-4).astype(np.float64)
    else:
        dones = (rng.uniform(size=B) < 0.1).astype(np.float64)
        costs = rng.uniform(0.1, 10, size=B)
    return xs.astype(np.float32), dones.astype(np.float32), costs.astype(np.float32)

"""Shared test helpers: build the product objects (q_learning_with_hjb_b200) and the matching oracle objects
from the same configuration, and compare results with angle-aware relative errors."""
import os

import numpy as np

from oracle import rollout_oracle as O

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "q_learning_with_hjb_b200")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

WRAP_IDX = {"linear": (), "cartpole": (1,), "acrobot": (0, 1), "quad2d": (2,), "quad10d": (3, 4)}
PAIRS = [("linear", "lqr"), ("cartpole", "cartpole_lqr"), ("cartpole", "cartpole_es"), ("acrobot", "acrobot_es"),
         ("quad2d", "quad2d_hover"), ("quad10d", "quad10d_hover")]
SCALES = {"linear": [2, 2], "cartpole": [3, 4, 3, 5], "acrobot": [4, 4, 6, 6], "quad2d": [2, 2, 4, 3, 3, 3],
          "quad10d": [2, 2, 2, 1.2, 1.2, 2, 2, 2, 2, 2]}


def make_dynamics(kind):
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


def make_controller(kind, dyn):
    if kind == "lqr":
        from q_learning_with_hjb_b200.controller.lqr import LQR
        return LQR(dyn, np.eye(2), np.eye(1))
    if kind == "cartpole_lqr":   # the notebook's inline LQR about xf = [0, 3.1415926, 0, 0], unclipped
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
    if kind == "quad2d_track":
        planner = QC.Quadrotors2DWaypointsPlanner(O.TRACK_WAYPOINTS, dyn, avg_speed=O.TRACK_SPEED)
        return QC.Quadrotors2DTrackingController(dyn, planner, np.eye(6), np.eye(2))
    raise ValueError(kind)


def oracle_pair(skind, ckind):
    sys = O.std_system(skind)
    return sys, O.std_controller(ckind, sys)


def rand_states(skind, B, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, size=(B, len(SCALES[skind]))) * np.asarray(SCALES[skind], dtype=np.float64)


def angle_diff(a, b, wrap_idx):
    """a - b with the angle components compared modulo 2 pi."""
    d = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
    for i in wrap_idx:
        d[..., i] = np.remainder(d[..., i] + np.pi, 2 * np.pi) - np.pi
    return d


def rel_err(a, b, wrap_idx=()):
    """max |a - b| / max(1, |b|) per element — the fp32 relative error with an absolute floor of 1."""
    d = np.abs(angle_diff(a, b, wrap_idx))
    return float(np.max(d / np.maximum(1.0, np.abs(np.asarray(b, dtype=np.float64)))))

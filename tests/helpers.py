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
    """The package's Dynamics object of a kind (q_learning_with_hjb_b200/workloads.py: one definition for tests and bench)."""
    from q_learning_with_hjb_b200 import workloads
    return workloads.make_dynamics(kind)


def make_controller(kind, dyn):
    from q_learning_with_hjb_b200 import workloads
    if kind == "quad2d_track":
        from q_learning_with_hjb_b200.controller import quadrotors_model_based_controller as QC
        planner = QC.Quadrotors2DWaypointsPlanner(O.TRACK_WAYPOINTS, dyn, avg_speed=O.TRACK_SPEED)
        return QC.Quadrotors2DTrackingController(dyn, planner, np.eye(6), np.eye(2))
    return workloads.make_controller(kind, dyn)


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

"""Dynamics config dataclasses — same class names, field names and float32 casting as the reference's
configs/dynamics/dynamics_config.py:6-59, so its .gin files bind unchanged."""
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from q_learning_with_hjb_b200.configs import gin_compat as gin


def _f32(v):
    return np.array(v, dtype=np.float32)


@dataclass
class DynamicsConfig:
    seed: int
    dt: float
    umin: Sequence[float]
    umax: Sequence[float]
    x0_mean: Sequence[float]
    x0_std: Sequence[float]

    def __post_init__(self):
        # reference: dynamics_config.py:15-21 (everything becomes np.float32)
        self.x0_mean, self.x0_std = _f32(self.x0_mean), _f32(self.x0_std)
        self.umin, self.umax = _f32(self.umin), _f32(self.umax)
        self.state_dim = int(self.x0_mean.shape[0])
        self.control_dim = int(self.umin.shape[0])


@gin.configurable
@dataclass
class LinearDynamicsConfig(DynamicsConfig):
    A: Sequence[Sequence[float]]
    B: Sequence[Sequence[float]]

    def __post_init__(self):
        super().__post_init__()
        self.A, self.B = _f32(self.A), _f32(self.B)


@gin.configurable
@dataclass
class CartpoleDynamicsConfig(DynamicsConfig):
    mc: float
    mp: float
    g: float
    l: float


@gin.configurable
@dataclass
class Quadrotors2DConfig(DynamicsConfig):
    g: float
    m: float
    r: float
    I: float


@gin.configurable
@dataclass
class NearHoverQuadcopterConfig(DynamicsConfig):
    g: float
    m: float
    kT: float
    n0: float

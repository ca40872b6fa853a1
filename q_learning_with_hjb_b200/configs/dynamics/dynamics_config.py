"""Dynamics config classes — the class names, field names, field order and float32 casting of the reference's
configs/dynamics/dynamics_config.py:6-59, so that its .gin files bind unchanged (``Class.field = literal``).

The classes are generated from one field table: every system shares the integrator / limit / initial-state fields and adds
its physical parameters."""
from dataclasses import field, make_dataclass
from typing import Sequence

import numpy as np

from q_learning_with_hjb_b200.configs import gin_compat as gin

_VEC, _MAT = Sequence[float], Sequence[Sequence[float]]

#: shared by every system: RNG seed, step, input limits, initial-state box (mean +- std, uniform)
_SHARED = (("seed", int), ("dt", float), ("umin", _VEC), ("umax", _VEC), ("x0_mean", _VEC), ("x0_std", _VEC))

#: class name -> the system's own fields; array-valued ones are listed in _ARRAYS
_SYSTEMS = {
    "LinearDynamicsConfig": (("A", _MAT), ("B", _MAT)),                                         # x' = A x + B u
    "CartpoleDynamicsConfig": (("mc", float), ("mp", float), ("g", float), ("l", float)),       # cart / pole mass, gravity, length
    "Quadrotors2DConfig": (("g", float), ("m", float), ("r", float), ("I", float)),             # gravity, mass, arm, inertia
    "NearHoverQuadcopterConfig": (("g", float), ("m", float), ("kT", float), ("n0", float)),    # gravity, mass, thrust / torque gains
}
_ARRAYS = ("umin", "umax", "x0_mean", "x0_std", "A", "B")


def _f32(v):
    return np.array(v, dtype=np.float32)


def _cast(self):
    """Everything array-valued becomes np.float32 (reference: dynamics_config.py:15-21); dimensions are derived."""
    for name in _ARRAYS:
        if hasattr(self, name):
            setattr(self, name, _f32(getattr(self, name)))
    self.state_dim = int(self.x0_mean.shape[0])
    self.control_dim = int(self.umin.shape[0])


def _make(name, fields, bases=()):
    cls = make_dataclass(name, [(f, t, field()) for f, t in fields], bases=bases, namespace={"__post_init__": _cast})
    cls.__module__ = __name__
    return cls


DynamicsConfig = _make("DynamicsConfig", _SHARED)
for _name, _own in _SYSTEMS.items():
    globals()[_name] = gin.configurable(_make(_name, _own, bases=(DynamicsConfig,)))
del _name, _own

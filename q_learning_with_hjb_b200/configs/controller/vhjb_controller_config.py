"""``VHJBControllerConfig`` — the field names, order and float32 casting of the reference's
configs/controller/vhjb_controller_config.py:6-68, so that its *_vhjb_controller.gin files bind unchanged.

Generated from a grouped field table (the groups are the sections of the .gin files)."""
from dataclasses import field, make_dataclass
from typing import Sequence

import numpy as np

from q_learning_with_hjb_b200.configs import gin_compat as gin

_VEC, _MAT = Sequence[float], Sequence[Sequence[float]]

_GROUPS = {
    "general": (("seed", int), ("epsilon", float)),                       # epsilon: additive guard against division by zero
    "value network": (("features", Sequence[int]), ("normalization_mean", _VEC), ("normalization_std", _VEC),
                      ("epsilon_scalar", float), ("using_batch_norm", bool)),
    "optimisation": (("lr", float), ("epochs", int), ("batch_size", int)),
    "termination-loss weight (SGDR cycles)": tuple(
        ("regularization_" + k, t) for k, t in (("init_value", float), ("peak_value", float), ("end_value", float),
                                                ("num_of_cycles", int), ("warmup_steps_per_cycle", int),
                                                ("total_steps_per_cycle", int))),
    "seed dataset": (("num_of_interior_data", int), ("num_of_boundary_data", int), ("interior_states_mean", _VEC),
                     ("interior_states_std", _VEC), ("boundary_states_mean", _VEC), ("boundary_states_std", _VEC),
                     ("boundary_cost_clip", float)),
    "trajectory sampling": (("num_of_trajectories_per_epoch", int), ("maximum_step", int), ("maximum_buffer_size", int)),
    "task": (("Q", _MAT), ("R", _MAT), ("xf", _VEC), ("uf", _VEC), ("obs_min", _VEC), ("obs_max", _VEC)),   # obs_*: error coordinates
}
_FIELDS = [f for group in _GROUPS.values() for f in group]
_ARRAY_FIELDS = tuple(name for name, t in _FIELDS if t in (_VEC, _MAT))


def _cast(self):
    for name in _ARRAY_FIELDS:
        setattr(self, name, np.array(getattr(self, name), dtype=np.float32))


VHJBControllerConfig = make_dataclass("VHJBControllerConfig", [(n, t, field()) for n, t in _FIELDS],
                                      namespace={"__post_init__": _cast})
VHJBControllerConfig.__module__ = __name__
VHJBControllerConfig = gin.configurable(VHJBControllerConfig)

"""VHJBControllerConfig — same field names and float32 casting as the reference's
configs/controller/vhjb_controller_config.py:6-68, so its *_vhjb_controller.gin files bind unchanged."""
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from q_learning_with_hjb_b200.configs import gin_compat as gin

_ARRAY_FIELDS = ("normalization_mean", "normalization_std", "Q", "R", "xf", "uf",
                 "interior_states_mean", "interior_states_std", "boundary_states_mean", "boundary_states_std",
                 "obs_min", "obs_max")


@gin.configurable
@dataclass
class VHJBControllerConfig:
    # general
    seed: int
    epsilon: float                      # additive guard against division by zero
    # value network
    features: Sequence[int]
    normalization_mean: Sequence[float]
    normalization_std: Sequence[float]
    epsilon_scalar: float
    using_batch_norm: bool
    # optimisation
    lr: float
    epochs: int
    batch_size: int
    # termination-loss weight schedule (SGDR cycles)
    regularization_init_value: float
    regularization_peak_value: float
    regularization_end_value: float
    regularization_num_of_cycles: int
    regularization_warmup_steps_per_cycle: int
    regularization_total_steps_per_cycle: int
    # seed dataset
    num_of_interior_data: int
    num_of_boundary_data: int
    interior_states_mean: Sequence[float]
    interior_states_std: Sequence[float]
    boundary_states_mean: Sequence[float]
    boundary_states_std: Sequence[float]
    boundary_cost_clip: float
    # trajectory sampling
    num_of_trajectories_per_epoch: int
    maximum_step: int
    maximum_buffer_size: int
    # task
    Q: Sequence[Sequence[float]]
    R: Sequence[Sequence[float]]
    xf: Sequence[float]
    uf: Sequence[float]
    obs_min: Sequence[float]            # in error coordinates
    obs_max: Sequence[float]

    def __post_init__(self):
        for name in _ARRAY_FIELDS:
            setattr(self, name, np.array(getattr(self, name), dtype=np.float32))

"""``Quadrotors2D`` (6-D planar quadrotor) and ``NearHoverQuadcopter`` (10-D) —
reference: dynamics/quadrotors.py:9-70 and :102-170."""
import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.dynamics.dynamics_basic import Dynamics


class Quadrotors2D(Dynamics):
    """x = [x, y, theta, dx, dy, dtheta], u = the two rotor thrusts."""
    KIND = L.SYS_QUAD2D
    WRAP_INDEX = (2,)

    def __init__(self, config) -> None:
        super().__init__(config)
        self.g, self.m, self.r, self.I = config.g, config.m, config.r, config.I

    def system_params(self):
        return [self.g, self.m, self.r, self.I], np.zeros(0), np.zeros(0)


class NearHoverQuadcopter(Dynamics):
    """x = [p_x, p_y, p_z, theta_x, theta_y, v_x, v_y, v_z, omega_x, omega_y], u = [Tz, Sx, Sy]."""
    KIND = L.SYS_QUAD10D
    WRAP_INDEX = (3, 4)

    def __init__(self, config) -> None:
        super().__init__(config)
        self.g, self.kT, self.m, self.n0 = config.g, config.kT, config.m, config.n0

    def system_params(self):
        return [self.g, self.m, self.kT, self.n0], np.zeros(0), np.zeros(0)

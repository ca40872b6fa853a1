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

    def linearize(self, xf, uf):
        """(A, B) of xdot ~ A (x - xf) + B (u - uf) about (xf, uf), host side, once (at hover: the matrices of
        controller/quadrotors_model_based_controller.py:25-31)."""
        th, thrust = xf[2], (uf[0] + uf[1]) / self.m
        A = np.zeros((6, 6))
        A[:3, 3:] = np.eye(3)
        A[3, 2] = -np.cos(th) * thrust
        A[4, 2] = -np.sin(th) * thrust
        B = np.zeros((6, 2))
        B[3, :] = -np.sin(th) / self.m
        B[4, :] = np.cos(th) / self.m
        B[5, :] = [self.r / self.I, -self.r / self.I]
        return A, B

    def plot_trajectory(self, ts, xs):
        """Animation of a trajectory (reference: dynamics/quadrotors.py:72-104); needs matplotlib."""
        from q_learning_with_hjb_b200.utils import plotting as P
        xs = np.asarray(xs)
        return P.animate(ts, xs, lambda x: P.quad2d_frame(x, self.r), P.span(xs[:, 0], 2 * self.r), P.span(xs[:, 1], 2 * self.r),
                         trail=lambda x: x[:2])


class NearHoverQuadcopter(Dynamics):
    """x = [p_x, p_y, p_z, theta_x, theta_y, v_x, v_y, v_z, omega_x, omega_y], u = [Tz, Sx, Sy]."""
    KIND = L.SYS_QUAD10D
    WRAP_INDEX = (3, 4)

    def __init__(self, config) -> None:
        super().__init__(config)
        self.g, self.kT, self.m, self.n0 = config.g, config.kT, config.m, config.n0

    def system_params(self):
        return [self.g, self.m, self.kT, self.n0], np.zeros(0), np.zeros(0)

    def linearize(self, xf, uf):
        """(A, B) about (xf, uf) (at hover: controller/quadrotors_model_based_controller.py:58-68)."""
        A = np.zeros((10, 10))
        A[:5, 5:] = np.eye(5)
        A[5, 3] = self.g / np.cos(xf[3]) ** 2
        A[6, 4] = self.g / np.cos(xf[4]) ** 2
        B = np.zeros((10, 3))
        B[7, 0] = self.kT / self.m
        B[8, 1] = B[9, 2] = self.n0
        return A, B

    def plot_trajectory(self, ts, xs):
        """3-D animation of a trajectory (reference: dynamics/quadrotors.py:172-196); needs matplotlib."""
        from q_learning_with_hjb_b200.utils import plotting as P
        xs = np.asarray(xs)
        return P.animate(ts, xs, P.quad10d_frame, P.span(xs[:, 0], 0.5), P.span(xs[:, 1], 0.5), P.span(xs[:, 2], 0.5),
                         three_d=True, trail=lambda x: x[:3])

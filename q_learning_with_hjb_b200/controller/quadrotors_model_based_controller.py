"""Hover LQR controllers for the planar quadrotor and the 10-D near-hover quadcopter:
u = clip(-K wrap(x - xf) + uf, umin, umax)  (reference: controller/quadrotors_model_based_controller.py:7-75).

``Quadrotors2DWaypointsPlanner`` (the reference's min-snap planner, :77-233) is a one-off host-side linear
solve that no rollout uses; it is out of scope."""
import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.controller.controller_basic import DeviceController, lqr_gain
from q_learning_with_hjb_b200.dynamics.quadrotors import NearHoverQuadcopter, Quadrotors2D


class _HoverLQR(DeviceController):
    def _finish(self, dynamics, xf, Q, R, n_pos):
        self.dynamics = dynamics
        self.xf, self.Q, self.R = np.asarray(xf), np.asarray(Q), np.asarray(R)
        self.umin, self.umax = dynamics.get_control_limit()
        if np.linalg.norm(self.xf[n_pos:]) > 0:
            raise ValueError("Final Velocity or Angle is not zero")

    def control_spec(self):
        c = L.HjbControl()
        c.kind, c.clip = L.CTL_FEEDBACK, 1
        L.fill(c.K, self.K)
        L.fill(c.xf, self.xf)
        L.fill(c.uf, self.uf)
        return c


class Quadrotors2DHoveringController(_HoverLQR):
    def __init__(self, dynamics: Quadrotors2D, xf, Q, R) -> None:
        super().__init__()
        self._finish(dynamics, xf, Q, R, n_pos=2)
        d = dynamics
        self.uf = d.m * d.g / 2 * np.ones(2)
        # hover linearisation (:25-31): d(ddx)/dtheta = -g, thrust sum drives ddy, thrust difference dtheta
        self.A = np.zeros((6, 6))
        self.A[:3, 3:] = np.eye(3)
        self.A[3, 2] = -d.g
        self.B = np.zeros((6, 2))
        self.B[4, :] = 1.0 / d.m
        self.B[5, :] = [d.r / d.I, -d.r / d.I]
        self.K, self.P = lqr_gain(self.A, self.B, self.Q, self.R)


class NearHoverQuadcopterHoveringController(_HoverLQR):
    def __init__(self, dynamics: NearHoverQuadcopter, xf, Q, R) -> None:
        super().__init__()
        self._finish(dynamics, xf, Q, R, n_pos=3)
        d = dynamics
        self.uf = np.array([d.g * d.m / d.kT, 0.0, 0.0])
        # hover linearisation (:58-68): tan(theta) ~ theta
        self.A = np.zeros((10, 10))
        self.A[:5, 5:] = np.eye(5)
        self.A[5, 3] = self.A[6, 4] = d.g
        self.B = np.zeros((10, 3))
        self.B[7, 0] = d.kT / d.m
        self.B[8, 1] = self.B[9, 2] = d.n0
        self.K, self.P = lqr_gain(self.A, self.B, self.Q, self.R)

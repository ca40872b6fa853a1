"""``LQR``: u = clip(-K x, umin, umax) for a ``LinearDynamics`` (reference: controller/lqr.py:10-30)."""
import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.controller.controller_basic import DeviceController, lqr_gain
from q_learning_with_hjb_b200.dynamics.linear import LinearDynamics


class LQR(DeviceController):
    def __init__(self, dynamics: LinearDynamics, Q: np.ndarray, R: np.ndarray) -> None:
        super().__init__()
        Q, R = np.asarray(Q), np.asarray(R)
        assert Q.ndim == 2 and R.ndim == 2
        assert Q.shape[0] == Q.shape[1] and R.shape[0] == R.shape[1]
        assert dynamics.A.shape[1] == Q.shape[1] and dynamics.B.shape[1] == R.shape[1]
        self.dynamics, self.Q, self.R = dynamics, Q, R
        self.K, self.P = lqr_gain(dynamics.A, dynamics.B, Q, R)
        self.umin, self.umax = dynamics.get_control_limit()

    def control_spec(self):
        c = L.HjbControl()
        c.kind, c.clip = L.CTL_FEEDBACK, 1   # regulates the raw state to the origin: xf = 0, uf = 0
        L.fill(c.K, self.K)
        return c


class StateFeedback(DeviceController):
    """u = -K wrap(x - xf) + uf, optionally clipped — the LQR the reference's notebooks write inline
    (examples/cartpole_balancing.ipynb cell 4:24-25 [unclipped], drone_hovering.ipynb cell 15:1-2)."""

    def __init__(self, dynamics, K, xf=None, uf=None, clip: bool = False) -> None:
        super().__init__()
        self.dynamics = dynamics
        n, m = dynamics.get_dimension()
        self.K = np.asarray(K, dtype=np.float64).reshape(m, n)
        self.xf = np.zeros(n) if xf is None else np.asarray(xf, dtype=np.float64)
        self.uf = np.zeros(m) if uf is None else np.asarray(uf, dtype=np.float64)
        self.clip = bool(clip)

    def control_spec(self):
        c = L.HjbControl()
        c.kind, c.clip = L.CTL_FEEDBACK, int(self.clip)
        L.fill(c.K, self.K)
        L.fill(c.xf, self.xf)
        L.fill(c.uf, self.uf)
        return c

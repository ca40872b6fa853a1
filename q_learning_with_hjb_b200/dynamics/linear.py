"""``LinearDynamics``: x' = A x + B u (reference: dynamics/linear.py:7-22)."""
import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.dynamics.dynamics_basic import Dynamics


class LinearDynamics(Dynamics):
    KIND = L.SYS_LINEAR
    WRAP_INDEX = ()

    def __init__(self, config) -> None:
        super().__init__(config)
        A, B = np.asarray(config.A), np.asarray(config.B)
        assert A.ndim == 2 and B.ndim == 2
        assert A.shape[0] == A.shape[1] == B.shape[0]
        self.A, self.B = config.A, config.B

    def states_wrap(self, x):
        return x  # no angles (linear.py:17-18)

    def system_params(self):
        return np.zeros(0), self.A, self.B

    def discretized(self, dt=None):
        """Exact zero-order-hold discretisation (A_d, B_d) for ``integrator="discrete"``
        (examples/double_integrator_optimal_time.ipynb cell 4)."""
        import scipy.signal

        n, m = self.state_dim, self.control_dim
        Ad, Bd, *_ = scipy.signal.cont2discrete((np.asarray(self.A, np.float64), np.asarray(self.B, np.float64),
                                                  np.eye(n), np.zeros((n, m))), dt=self.dt if dt is None else dt)
        return Ad, Bd

    def linearize(self, xf, uf):
        """(A, B) of xdot ~ A (x - xf) + B (u - uf): the system matrices themselves."""
        return np.asarray(self.A, dtype=np.float64), np.asarray(self.B, dtype=np.float64)

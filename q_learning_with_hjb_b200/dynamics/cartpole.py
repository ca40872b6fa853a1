"""``Cartpole``: x = [p, theta, dp, dtheta], theta = pi upright (reference: dynamics/cartpole.py:10-64)."""
import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.dynamics.dynamics_basic import Dynamics


class Cartpole(Dynamics):
    KIND = L.SYS_CARTPOLE
    WRAP_INDEX = (1,)

    def __init__(self, config) -> None:
        super().__init__(config)
        self.mc, self.mp, self.l, self.g = config.mc, config.mp, config.l, config.g

    # manipulator matrices, host side: used once by the model-based controllers to linearise about xf
    def get_M(self, x):
        k = self.mp * self.l * np.cos(x[1])
        return np.array([[self.mc + self.mp, k], [k, self.mp * self.l ** 2]])

    def get_C(self, x):
        return np.array([[0.0, -self.mp * self.l * x[3] * np.sin(x[1])], [0.0, 0.0]])

    def get_G(self, x):
        return np.array([0.0, self.mp * self.g * self.l * np.sin(x[1])])

    def get_B(self):
        return np.array([1, 0])

    def system_params(self):
        return [self.mc, self.mp, self.l, self.g], np.zeros(0), np.zeros(0)

    def linearize(self, xf, uf):
        """(A, B) about an equilibrium (xf, uf), host side, once: A = [[0, I], [-M^-1 dG/dq, 0]], B = [0; M^-1 B]
        (the M^-1 derivative multiplies B uf - G(qf) = 0 at an equilibrium; C dq vanishes with dq = 0)."""
        Minv = np.linalg.inv(self.get_M(xf))
        dG_dq = np.array([[0.0, 0.0], [0.0, self.mp * self.g * self.l * np.cos(xf[1])]])
        A = np.zeros((4, 4))
        A[0, 2] = A[1, 3] = 1.0
        A[2:, :2] = -Minv @ dG_dq
        B = np.concatenate([np.zeros(2), Minv @ self.get_B()]).reshape(4, 1)
        return A, B

    def plot_trajectory(self, ts, xs, cart_width=0.4, cart_height=0.2, pole_radius=0.05, x_range=np.array([-2, 2]),
                        y_range=np.array([-1, 3])):
        """Animation of a trajectory (reference: dynamics/cartpole.py:66-103); needs matplotlib."""
        from q_learning_with_hjb_b200.utils import plotting as P
        return P.animate(ts, xs, lambda x: P.cartpole_frame(x, self.l, cart_width, cart_height), tuple(x_range), tuple(y_range))

"""``Acrobot``: x = [q1, q2, dq1, dq2] (reference: dynamics/acrobot.py:19-81, parameters :8-16).

The reference's constructor is broken at HEAD (acrobot.py:22 calls ``super().__init__()`` without the
config the base class requires); this one takes the same ``params`` dict and works.
"""
import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.dynamics.dynamics_basic import Dynamics

dt = 0.05
p = {"l1": 0.5, "l2": 1, "m1": 8, "m2": 8, "I1": 2, "I2": 8, "umax": 25, "g": 10, "dt": dt}


class Acrobot(Dynamics):
    KIND = L.SYS_ACROBOT
    WRAP_INDEX = (0, 1)

    def __init__(self, params=p, seed: int = 0) -> None:
        self.p = dict(params)
        self.dim = 2
        self.state_dim, self.control_dim = 4, 1
        self.m1, self.m2, self.l1, self.l2 = self.p["m1"], self.p["m2"], self.p["l1"], self.p["l2"]
        self.I1, self.I2, self.g, self.dt = self.p["I1"], self.p["I2"], self.p["g"], self.p["dt"]
        self.umax = np.array([self.p["umax"]], dtype=np.float32)
        self.umin = -self.umax
        # the reference starts its demo from x0 = [0.001, 0, 0, 0] (acrobot_energy_shaping.py:131)
        self.x0_mean = np.zeros(4, dtype=np.float32)
        self.x0_std = np.full(4, 0.1, dtype=np.float32)
        self.seed = seed
        np.random.seed(seed)
        self.fast_trig = True

    def get_dimension(self):
        return self.dim * 2, 1

    def get_control_limit(self):
        return self.umin, self.umax

    def get_M(self, x):
        a = self.m2 * self.l1 * self.l2 / 2
        c2 = np.cos(x[1])
        return np.array([[self.I1 + self.I2 + self.m2 * self.l1 ** 2 + 2 * a * c2, self.I2 + a * c2],
                         [self.I2 + a * c2, self.I2]])

    def get_C(self, x):
        a = self.m2 * self.l1 * self.l2 / 2
        s2 = np.sin(x[1])
        return np.array([[-2 * a * s2 * x[3], -a * s2 * x[3]], [a * s2 * x[2], 0.0]])

    def get_G(self, x):
        g12 = self.m2 * self.g * self.l2 / 2 * np.sin(x[0] + x[1])
        return np.array([(self.m1 * self.l1 / 2 + self.m2 * self.l1) * self.g * np.sin(x[0]) + g12, g12])

    def get_B(self):
        return np.array([0, 1])

    def energy(self, x):
        """Total mechanical energy (acrobot.py:60-70); E(upright) = 100, E(hanging) = -100 for ``p``."""
        a = self.m2 * self.l1 * self.l2 / 2
        c1, c2, c12 = np.cos(x[0]), np.cos(x[1]), np.cos(x[0] + x[1])
        kinetic = 0.5 * (self.I1 + self.m2 * self.l1 ** 2 + self.I2 + 2 * a * c2) * x[2] ** 2 \
            + 0.5 * self.I2 * x[3] ** 2 + (self.I2 + a * c2) * x[2] * x[3]
        potential = -(self.m1 * self.l1 / 2 + self.m2 * self.l1) * self.g * c1 - self.m2 * self.g * self.l2 / 2 * c12
        return kinetic + potential

    def system_params(self):
        return [self.l1, self.l2, self.m1, self.m2, self.I1, self.I2, self.g], np.zeros(0), np.zeros(0)

    def linearize(self, xf, uf):
        """(A, B) about an equilibrium (xf, uf); see Cartpole.linearize."""
        Minv = np.linalg.inv(self.get_M(xf))
        g1 = (self.m1 * self.l1 / 2 + self.m2 * self.l1) * self.g
        g12 = self.m2 * self.g * self.l2 / 2
        c1, c12 = np.cos(xf[0]), np.cos(xf[0] + xf[1])
        dG_dq = np.array([[g1 * c1 + g12 * c12, g12 * c12], [g12 * c12, g12 * c12]])
        A = np.zeros((4, 4))
        A[0, 2] = A[1, 3] = 1.0
        A[2:, :2] = -Minv @ dG_dq
        B = np.concatenate([np.zeros(2), Minv @ self.get_B()]).reshape(4, 1)
        return A, B

    def plot_trajectory(self, ts, xs, margin=0.5):
        """Animation of a trajectory (reference: dynamics/acrobot.py:82-110); needs matplotlib."""
        from q_learning_with_hjb_b200.utils import plotting as P
        reach = self.l1 + self.l2 + margin
        return P.animate(ts, xs, lambda x: P.acrobot_frame(x, self.l1, self.l2), (-reach, reach), (-reach, reach))

"""``AcrobotEnergyShapingController`` — Spong's collocated energy-based swing-up with an LQR catch
(reference: controller/acrobot_energy_shaping.py:9-121).  As for the cart-pole, the Riccati solution the
reference recomputes per step (:112) is solved once."""
import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.controller.controller_basic import DeviceController, closed_loop, lqr_gain, unclipped
from q_learning_with_hjb_b200.dynamics.acrobot import Acrobot


def wrap(q):
    return (q + np.pi) % (2 * np.pi) - np.pi


class AcrobotEnergyShapingController(DeviceController):
    def __init__(self, acrobot_system: Acrobot, Q=np.eye(4), R=np.eye(1), eps=1000, K=np.array([1, 2, 1])) -> None:
        super().__init__()
        self.acrobot = self.dynamics = acrobot_system
        self.xf = np.array([np.pi, 0, 0, 0])
        self.K, self.Q, self.R, self.eps = np.asarray(K), np.asarray(Q), np.asarray(R), eps
        self._lqr = None

    def get_linearized_dynamics(self):
        """xdot ~ Alin (x - xf) + Blin u about the upright equilibrium (:23-46)."""
        ac = self.acrobot
        Minv = np.linalg.inv(ac.get_M(self.xf))
        g12 = ac.m2 * ac.g * ac.l2 / 2
        g1 = (ac.m1 * ac.l1 / 2 + ac.m2 * ac.l1) * ac.g
        dG_dq = -np.array([[g1 + g12, g12], [g12, g12]])               # dG/dq at q = (pi, 0)
        Alin = np.zeros((4, 4))
        Alin[0, 2] = Alin[1, 3] = 1.0
        Alin[2:, :2] = -Minv @ dG_dq
        Blin = np.concatenate([np.zeros(2), Minv @ ac.get_B()]).reshape(4, 1)
        return Alin, Blin

    def get_lqr_term(self):
        if self._lqr is None:
            self._lqr = lqr_gain(*self.get_linearized_dynamics(), self.Q, self.R)
        return self._lqr

    def control_spec(self):
        K_lqr, P = self.get_lqr_term()
        c = L.HjbControl()
        c.kind, c.clip = L.CTL_ACROBOT_ES, 1
        L.fill(c.K, K_lqr)
        L.fill(c.P, P)
        L.fill(c.xf, self.xf)
        L.fill(c.aux, [self.K[0], self.K[1], self.K[2], self.eps])
        return c

    def get_swingup_input(self, x):
        """The collocated swing-up branch alone, un-clipped (:74-98): the device law with the LQR catch region emptied."""
        c = self.control_spec()
        c.aux[3] = -1.0                                   # dx^T P dx < eps never holds
        return self._efforts(unclipped(self.acrobot.system_spec()), c, x)


def test_acrobot(acrobot: Acrobot, acrobot_controller: AcrobotEnergyShapingController, tf=25.0, plot=True):
    """The reference's demo (:123-160): 25 s of closed loop from x0 = [0.001, 0, 0, 0] — one rollout launch.
    Returns (t, xs, us, energy)."""
    t = np.arange(0, tf, acrobot.dt)
    xs, us = closed_loop(acrobot, acrobot_controller, np.array([0.001, 0, 0, 0]), t)
    e = np.array([acrobot.energy(x) for x in xs])
    if plot:
        try:
            acrobot.plot_trajectory(t, xs)
        except ImportError:
            pass
    return t, xs, us, e

"""``CartpoleEnergyShapingController`` — energy pumping for swing-up, LQR to catch at the top
(reference: controller/cartpole_energy_shaping.py:7-110).

The reference re-solves the Riccati equation inside every ``get_control_efforts`` call (:77); K and P are
constants of (cartpole, Q, R), so they are solved once here and live in the kernel's constant bank.
"""
import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.controller.controller_basic import DeviceController, closed_loop, lqr_gain, unclipped
from q_learning_with_hjb_b200.dynamics.cartpole import Cartpole


class CartpoleEnergyShapingController(DeviceController):
    def __init__(self, cartpole: Cartpole, Q=np.eye(4), R=np.eye(1), K=np.array([4, 4, 10]), eps_energy=1,
                 eps_state=1) -> None:
        super().__init__()
        self.cartpole = self.dynamics = cartpole
        self.xf = np.array([0, np.pi, 0, 0])
        self.umin, self.umax = cartpole.get_control_limit()
        self.Q, self.R, self.K = np.asarray(Q), np.asarray(R), np.asarray(K)
        self.eps_energy, self.eps_state = eps_energy, eps_state
        self._lqr = None

    def get_linearized_dynamics(self):
        """xdot ~ Alin (x - xf) + Blin u about the upright equilibrium (:21-45)."""
        cp = self.cartpole
        Minv = np.linalg.inv(cp.get_M(self.xf))
        dG_dq = np.array([[0.0, 0.0], [0.0, -cp.mp * cp.g * cp.l]])   # dG/dq at theta = pi
        Alin = np.zeros((4, 4))
        Alin[0, 2] = Alin[1, 3] = 1.0
        Alin[2:, :2] = -Minv @ dG_dq
        Blin = np.concatenate([np.zeros(2), Minv @ cp.get_B()]).reshape(4, 1)
        return Alin, Blin

    def get_lqr_term(self):
        """(K [1x4], P [4x4]) of the linearised system; solved once and cached."""
        if self._lqr is None:
            self._lqr = lqr_gain(*self.get_linearized_dynamics(), self.Q, self.R)
        return self._lqr

    def energy(self, x):
        """Pole 'energy' 0.5 dtheta^2 - cos(theta) (:90-95)."""
        return 0.5 * x[3] ** 2 - np.cos(x[1])

    def control_spec(self):
        K_lqr, _ = self.get_lqr_term()
        c = L.HjbControl()
        c.kind, c.clip = L.CTL_CARTPOLE_ES, 1
        L.fill(c.K, K_lqr)
        L.fill(c.xf, self.xf)
        L.fill(c.aux, [self.K[0], self.K[1], self.K[2], self.eps_energy, self.eps_state])
        return c

    def get_energy_shaping_input(self, x):
        """The energy-pumping branch alone, un-clipped (:97-110): the device law with the LQR catch region emptied."""
        c = self.control_spec()
        c.aux[3] = -1.0                                   # |dE| < eps_energy never holds
        return self._efforts(unclipped(self.cartpole.system_spec()), c, x)


def test_cartpole(cartpole: Cartpole, cartpole_controller: CartpoleEnergyShapingController, tf=10.0, plot=True):
    """The reference's demo (:113-141): 10 s of closed loop from a random initial state, then the animation — the loop is
    one rollout launch.  Returns (t, xs, us)."""
    t = np.arange(0, tf, cartpole.dt)
    xs, us = closed_loop(cartpole, cartpole_controller, cartpole.get_initial_state(), t)
    if plot:
        try:
            cartpole.plot_trajectory(t, xs)
        except ImportError:
            pass
    return t, xs, us

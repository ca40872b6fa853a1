"""The restated oracle against the committed golden vectors (tests/golden/rollout_reference.npz, produced by
the UNMODIFIED reference through oracle/make_golden.py).  Runs anywhere — this is what pins the oracle on the
GPU box, where /root/reference does not exist."""
import os

import numpy as np
import pytest

from oracle import rollout_oracle as O
from tests.helpers import GOLDEN

G = np.load(os.path.join(GOLDEN, "rollout_reference.npz"))
CASES = [("linear", "lqr"), ("cartpole", "cartpole_es"), ("acrobot", "acrobot_es"), ("quad2d", "quad2d_hover"),
         ("quad10d", "quad10d_hover")]


@pytest.mark.parametrize("skind,ckind", CASES)
def test_per_state_quantities(skind, ckind):
    sys = O.std_system(skind)
    ctl = O.std_controller(ckind, sys)
    x, u = G[f"{skind}/x"], G[f"{skind}/u"]
    f, g = sys.f_g(x)
    np.testing.assert_allclose(f, G[f"{skind}/f"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(g, G[f"{skind}/g"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(sys.step(x, u, "euler"), G[f"{skind}/x_next"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ctl.control(sys, x), G[f"{skind}/{ckind}/u_ctl"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(ctl.K, G[f"{skind}/{ckind}/K"], rtol=1e-12)


@pytest.mark.parametrize("skind,ckind", CASES)
def test_closed_loop_trajectories(skind, ckind):
    sys = O.std_system(skind)
    ctl = O.std_controller(ckind, sys)
    tx, tu = G[f"{skind}/{ckind}/traj_x"], G[f"{skind}/{ckind}/traj_u"]
    xs, us, _, _ = O.rollout(sys, ctl, tx[0], tu.shape[0], "euler", record_stride=1)
    tol = 1e-6 if skind == "acrobot" else 1e-9   # chaotic: 1e-16 rounding differences are amplified
    np.testing.assert_allclose(xs, tx, rtol=tol, atol=tol)
    np.testing.assert_allclose(us, tu, rtol=tol * 100, atol=tol * 100)


def test_rk4_is_fourth_order_and_consistent_with_euler():
    # RK4 is the north-star extension (no reference counterpart): check its order on the cartpole
    sys = O.std_system("cartpole")
    x0 = np.array([[0.1, 2.9, -0.2, 0.3]])
    u = np.array([[1.5]])

    def integrate(dt, steps, method):
        s = O.OracleSystem(sys.kind, sys.n, sys.m, dt, sys.umin, sys.umax, sys.par)
        x = x0.copy()
        for _ in range(steps):
            x = s.step(x, u, method)
        return x
    ref = integrate(1e-4, 2000, "rk4")
    e1 = np.abs(integrate(0.02, 10, "rk4") - ref).max()
    e2 = np.abs(integrate(0.01, 20, "rk4") - ref).max()
    assert 10 < e1 / e2 < 24          # ~2^4
    ee = np.abs(integrate(0.01, 20, "euler") - ref).max()
    assert ee > 100 * e2

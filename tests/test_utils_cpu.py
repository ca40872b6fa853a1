"""Host-side helpers (q_learning_with_hjb_b200/utils/utils.py): the reference's utilities and their batched versions."""
import numpy as np
import scipy.linalg

from q_learning_with_hjb_b200.utils import utils as U


def test_np_collate_and_keep_first_element():
    batch = [(np.ones(3) * i, float(i), i % 2) for i in range(4)]
    xs, costs, dones = U.np_collate(batch)
    assert xs.shape == (4, 3) and costs.tolist() == [0.0, 1.0, 2.0, 3.0] and dones.tolist() == [0, 1, 0, 1]
    assert U.keep_first_element(lambda: (1, 2))() == 1 and U.keep_first_element(lambda: 5)() == 5


def test_are_matches_scipy_and_known_answer():
    # the reference's docstring example (utils/utils.py:38-50): A = B = Q = R = I_2 -> P = (1 + sqrt 2) I
    I2 = np.eye(2)
    np.testing.assert_allclose(U.solve_continuous_are(I2, I2, I2, I2), (1 + np.sqrt(2)) * I2, atol=1e-12)
    sols = U.solve_continuous_are(I2, I2, I2, I2, multiple_sol=True)
    assert any(np.allclose(P, (1 + np.sqrt(2)) * I2, atol=1e-7) for P in sols)
    assert any(np.allclose(P, (1 - np.sqrt(2)) * I2, atol=1e-7) for P in sols)
    for P in sols:                                       # every returned matrix solves the equation
        assert np.abs(I2.T @ P + P @ I2 - P @ P + I2).max() < 1e-6
    rng = np.random.default_rng(0)
    for n, m in ((2, 1), (4, 1), (6, 2), (10, 3)):
        A, B = rng.normal(size=(n, n)), rng.normal(size=(n, m))
        Q, R = np.eye(n) * rng.uniform(0.5, 2), np.eye(m) * rng.uniform(0.5, 2)
        np.testing.assert_allclose(U.solve_continuous_are(A, B, Q, R), scipy.linalg.solve_continuous_are(A, B, Q, R),
                                   rtol=1e-8, atol=1e-8)


def test_batched_are_matches_scipy_over_a_sweep():
    rng = np.random.default_rng(1)
    n, m, N = 6, 2, 64
    A, B = rng.normal(size=(n, n)), rng.normal(size=(n, m))
    q, r = rng.uniform(0.1, 10, size=N), rng.uniform(0.1, 10, size=N)
    Q = q[:, None, None] * np.eye(n)
    R = r[:, None, None] * np.eye(m)
    K, P = U.lqr_gains_batched(A, B, Q, R)
    assert P.shape == (N, n, n) and K.shape == (N, m, n)
    for i in range(0, N, 7):
        Pi = scipy.linalg.solve_continuous_are(A, B, Q[i], R[i])
        np.testing.assert_allclose(P[i].numpy(), Pi, rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(K[i].numpy(), np.linalg.solve(R[i], B.T @ Pi), rtol=1e-8, atol=1e-9)
    # the hover linearisations of the reference (integrator chains: eigenvalues of A on the imaginary axis are fine, the
    # Hamiltonian's are not on it)
    from oracle import rollout_oracle as O
    for AB in (O.quad2d_hover_AB(O.std_system("quad2d").par), O.quad10d_hover_AB(O.std_system("quad10d").par)):
        A, B = AB
        P = U.solve_continuous_are_batched(A, B, np.eye(A.shape[0]), np.eye(B.shape[1]))
        np.testing.assert_allclose(P.numpy(), scipy.linalg.solve_continuous_are(A, B, np.eye(A.shape[0]), np.eye(B.shape[1])),
                                   rtol=1e-8, atol=1e-8)


def test_linearize_batched_on_the_oracle_dynamics():
    from oracle import rollout_oracle as O
    sys = O.std_system("cartpole")
    xf = np.array([[0, np.pi, 0, 0], [0.3, 0.0, 0, 0]])
    uf = np.zeros((2, 1))
    A, B = U.linearize_batched(lambda x, u: sys.xdot(x, u), xf, uf)
    Aref, Bref = O.cartpole_linearisation(sys.par)
    np.testing.assert_allclose(A[0], Aref, atol=1e-6)
    np.testing.assert_allclose(B[0], Bref, atol=1e-6)
    assert A.shape == (2, 4, 4) and B.shape == (2, 4, 1) and not np.allclose(A[0], A[1])


def test_waypoints_planner_interpolates_and_is_a_model_trajectory():
    """Quadrotors2DWaypointsPlanner without the reference at hand: the minimum-snap polynomials pass through the way-points
    at the segment times, start and end at rest, are C^6 across interior points, hover after the last one, and — with the
    consistent theta'' — (x(t), u(t)) satisfies x' = f(x) + g(x) u of the planar quadrotor."""
    from oracle import rollout_oracle as O
    from q_learning_with_hjb_b200.controller.quadrotors_model_based_controller import Quadrotors2DWaypointsPlanner
    from tests.helpers import make_dynamics
    dyn = make_dynamics("quad2d")
    pts = np.array([[0.0, 0.0], [1.0, 0.5], [2.0, -0.3], [2.5, 1.0]])
    pl = Quadrotors2DWaypointsPlanner(pts, dyn, avg_speed=0.5, exact_theta_ddot=True)
    assert pl.coeff.shape == (2, 3, 8)
    for i, t in enumerate(pl.cumulated_t):
        x, u = pl.update(min(t, pl.cumulated_t[-1] - 1e-12) if i == len(pts) - 1 else t)
        np.testing.assert_allclose(x[:2], pts[i], atol=1e-8)
    x0, u0 = pl.update(0.0)
    np.testing.assert_allclose(x0[2:], 0.0, atol=1e-9)                              # at rest, level
    np.testing.assert_allclose(u0.sum(), dyn.m * dyn.g, rtol=1e-9)                  # hover thrust (the snap is free: a torque)
    xe, ue = pl.update(pl.cumulated_t[-1] + 3.0)                                    # past the end: hover at the last point
    np.testing.assert_allclose(xe, [2.5, 1.0, 0, 0, 0, 0], atol=1e-12)
    np.testing.assert_allclose(ue, [dyn.m * dyn.g / 2] * 2, rtol=1e-12)
    for n in range(0, 7):                                                            # continuity of derivatives 0..6
        for i in range(1, len(pts) - 1):
            left = pl.coeff[:, i - 1, :] @ pl.get_polynomial_term(pl.interval_t[i - 1], n)
            right = pl.coeff[:, i, :] @ pl.get_polynomial_term(0.0, n)
            np.testing.assert_allclose(left, right, atol=1e-6 * max(1.0, np.abs(left).max()))
    sys = O.std_system("quad2d")
    ts = np.linspace(0.05, 0.95, 7) * pl.cumulated_t[-1]
    xs, us = pl.plan(ts)
    h = 1e-5
    for t, x, u in zip(ts, xs, us):
        xm, _ = pl.update(t - h); xp, _ = pl.update(t + h)
        f, g = sys.f_g(x[None])
        np.testing.assert_allclose((xp - xm) / (2 * h), f[0] + g[0] @ u, rtol=1e-5, atol=1e-6)

"""Known-answer tests: numbers printed in the reference's saved notebook outputs, reproduced through the
oracle.  They depend only on NumPy's global RNG (seeded 0 by ``Dynamics.__init__``,
dynamics/dynamics_basic.py:26) and on how many ``get_initial_state()`` draws preceded the printed cell."""
import numpy as np
import scipy.linalg
import scipy.signal

from oracle import rollout_oracle as O


def _initial_states(std32, mean32, skip, count, wrap):
    """``Dynamics.get_initial_state`` (dynamics/dynamics_basic.py:28-29) after ``skip`` earlier draws."""
    np.random.seed(0)
    n = std32.shape[0]
    for _ in range(skip):
        np.random.uniform(size=(n,), low=-std32, high=std32)
    return np.stack([wrap((np.random.uniform(size=(n,), low=-std32, high=std32) + mean32)[None])[0]
                     for _ in range(count)])


def test_kat_cartpole_lqr_mean_cost():
    # examples/cartpole_balancing.ipynb cell 16: "mean lqr:  9.140986134043468"
    sys = O.std_system("cartpole")
    ctl = O.std_controller("cartpole_lqr", sys)
    np.testing.assert_allclose(ctl.K, [[-1, 34.38942857, -2.41062161, 10.70392355]], rtol=2e-8)
    x0 = _initial_states(np.float32([2.4, 0.05, 1, 0.05]), np.float32([0, 3.14, 0, 0]), 6001, 10, sys.wrap)
    cost = O.OracleCost(np.eye(4), np.eye(1), ctl.xf, ctl.uf)
    _, _, _, J = O.rollout(sys, ctl, x0, 500, "euler", record_stride=0, cost=cost)
    assert abs(J.mean() - 9.140986134043468) < 1e-12


def test_kat_quad2d_hover_cost():
    # examples/drone_hovering.ipynb cell 16: "lqr:  1.335421313313018", "mean lqr:  9.983921427754535"
    sys = O.std_system("quad2d")
    ctl = O.std_controller("quad2d_hover", sys)
    ctl.uf = np.array([4.905, 4.905])
    x0 = _initial_states(np.float32([1] * 6), np.float32([0] * 6), 12290, 10, sys.wrap)
    cost = O.OracleCost(np.eye(6), np.eye(2), ctl.xf, ctl.uf)
    _, _, _, J = O.rollout(sys, ctl, x0, 200, "euler", record_stride=0, cost=cost)
    assert abs(J[0] - 1.335421313313018) < 1e-13
    assert abs(J.mean() - 9.983921427754535) < 1e-12


def test_kat_quad10d_hover_cost():
    # examples/10D_quadcopte.ipynb cell 14: "lqr cost 9.085334056081662"
    sys = O.std_system("quad10d")
    ctl = O.std_controller("quad10d_hover", sys)
    x0 = _initial_states(np.float32([1, 1, 1, .5, .5, 1, 1, 1, .5, .5]), np.float32([0] * 10), 4000, 1, sys.wrap)
    cost = O.OracleCost(np.eye(10), np.eye(3), ctl.xf, ctl.uf)
    _, _, _, J = O.rollout(sys, ctl, x0, 400, "euler", record_stride=0, cost=cost)
    assert abs(J[0] - 9.085334056081662) < 1e-12


def test_kat_double_integrator_time_to_origin():
    # examples/double_integrator_optimal_time.ipynb cell 21: analytic 1.572 / 0.5365407719828942,
    # saturated LQR (R=0.01) 4.104 / 1.281602122345309 -- exact-ZOH stepping (cell 4), metric 1e-4.
    A = np.array([[0.0, 1.0], [0.0, 0.0]])
    B = np.array([[0.0], [1.0]])
    dt, T, metric = 0.01, 5, 1e-4
    Ad, Bd, *_ = scipy.signal.cont2discrete((A, B, np.eye(2), np.eye(1)), dt=dt)
    R = np.array([[0.01]])
    P = scipy.linalg.solve_continuous_are(A, B, np.eye(2), R)
    np.random.seed(0)
    np.random.uniform(low=-1, high=1, size=(2 ** 16, 2))
    for _ in range(6012):
        np.random.uniform(low=-1, high=1, size=(2,))
    x0s = [np.random.uniform(low=-1, high=1, size=(2,)) for _ in range(10)]
    ts = np.arange(0, T, dt)

    def analytic(x):
        if x @ x <= metric:
            return np.array([0.0])
        if (x[1] < 0 and x[0] <= 0.5 * x[1] ** 2) or (x[1] >= 0 and x[0] < -0.5 * x[1] ** 2):
            return np.array([1.0])
        return np.array([-1.0])

    def lqr(x):
        return np.clip(-np.linalg.inv(R) @ B.T @ P @ x, -1, 1)

    def time_to_origin(policy, x0):
        x, best = x0, T
        for t in ts:
            x = Ad @ x + Bd @ policy(x)
            if x @ x <= metric:
                best = min(t, best)
        return best

    ta = np.array([time_to_origin(analytic, x) for x in x0s])
    tl = np.array([time_to_origin(lqr, x) for x in x0s])
    assert abs(ta.mean() - 1.572) < 1e-12 and abs(ta.std() - 0.5365407719828942) < 1e-12
    assert abs(tl.mean() - 4.104) < 1e-12 and abs(tl.std() - 1.281602122345309) < 1e-12


def _notebook_cell21_states():
    """The ten initial states of double_integrator_optimal_time.ipynb cell 21 (NumPy's global stream as the notebook left it)."""
    np.random.seed(0)
    np.random.uniform(low=-1, high=1, size=(2 ** 16, 2))
    for _ in range(6012):
        np.random.uniform(low=-1, high=1, size=(2,))
    return np.stack([np.random.uniform(low=-1, high=1, size=(2,)) for _ in range(10)])


def _double_integrator():
    return O.OracleSystem("linear", 2, 1, 0.01, np.array([-1.0]), np.array([1.0]),
                          {"A": np.array([[0.0, 1.0], [0.0, 0.0]]), "B": np.array([[0.0], [1.0]])})


def _level_set_fixture():
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "double_integrator_level_set.npz"))
    V, vel = d["value_level_set"], d["vel"]
    return (V[2:, :] - V[:-2, :]) / (2 * float(d["dv"])), d["pos"], vel[1:-1], d


def _oracle_time_to_origin(ctl, x0, T=500, dt=0.01, metric=1e-4):
    xs, _, _, _ = O.rollout(_double_integrator(), ctl, x0, T, "discrete", record_stride=1)
    hit = (xs[1:] ** 2).sum(-1) <= metric                     # the state AFTER step k
    return np.where(hit.any(0), np.argmax(hit, 0) * dt, T * dt)


def test_kat_double_integrator_level_set_and_analytic_policies_of_the_oracle():
    # examples/double_integrator_optimal_time.ipynb cell 21 prints "mean level set 1.6170000000000002 / std level set
    # 0.6021802055863344" and "mean analytical set 1.572 / std 0.5365407719828942": the oracle's grid_sign / switch_curve
    # controllers and its exact-ZOH step reproduce both from the reference's own level-set data (tests/golden/
    # double_integrator_level_set.npz, oracle/make_golden_level_set.py).
    dVdvel, pos, vel, _ = _level_set_fixture()
    x0 = _notebook_cell21_states()
    tl = _oracle_time_to_origin(O.OracleController("grid_sign", grid=dVdvel, grid_axes=(pos, vel)), x0)
    ta = _oracle_time_to_origin(O.OracleController("switch_curve"), x0)
    assert abs(tl.mean() - 1.6170000000000002) < 1e-12 and abs(tl.std() - 0.6021802055863344) < 1e-12
    assert abs(ta.mean() - 1.572) < 1e-12 and abs(ta.std() - 0.5365407719828942) < 1e-12


def test_nearest_node_is_scipys_nearest_interpolation():
    # the notebook's policy goes through scipy's RegularGridInterpolator(method="nearest", bounds_error=False,
    # fill_value=None): the oracle's restatement must pick the same node everywhere, ties and extrapolation included
    from scipy.interpolate import RegularGridInterpolator
    rng = np.random.default_rng(0)
    pos, vel = np.linspace(-1, 1, 101), np.linspace(-1, 1, 101)[1:-1]
    T = rng.standard_normal((99, 101))
    f = RegularGridInterpolator((vel, pos), T, method="nearest", bounds_error=False, fill_value=None)
    x = rng.uniform(-1.3, 1.3, size=(100000, 2))
    x[:1000, 0] = pos[rng.integers(0, 100, 1000)] + 0.01       # (near-)ties
    x[1000:2000, 1] = vel[rng.integers(0, 98, 1000)] + 0.01
    x[2000:3000, 0] = pos[rng.integers(0, 101, 1000)]          # on the nodes
    mine = T[O.nearest_node(x[:, 1], vel), O.nearest_node(x[:, 0], pos)]
    assert np.array_equal(mine, f(x[:, ::-1]))


def test_level_set_fixture_agrees_with_the_analytic_minimum_time():
    # the fixture's analytic value function is the double integrator's closed-form minimum time (to the origin) within
    # one time step; the level-set solver's surface lies within its own (boundary-dominated) error of it
    _, pos, _, d = _level_set_fixture()
    P, V = np.meshgrid(pos, d["vel"])
    s = np.where(P > -0.5 * V * np.abs(V), 1.0, -1.0)
    Tstar = s * V + 2 * np.sqrt(np.maximum(0.5 * V * V + s * P, 0.0))
    assert np.abs(Tstar - d["value_analytic"]).max() < 0.011
    assert np.abs(d["value_level_set"] - d["value_analytic"]).mean() < 0.3


def test_kat_are_identity():
    # examples/nonpostive-definite-neural-structures.ipynb cell 4: A=B=Q=R=I2 -> P = (1+sqrt 2) I
    _, P = O.lqr_gain(np.eye(2), np.eye(2), np.eye(2), np.eye(2))
    np.testing.assert_allclose(P, (1 + np.sqrt(2)) * np.eye(2), atol=1e-12)


def test_gain_constants():
    # SURVEY.md §8a A9/A10 gains quoted from the reference's own ctor code paths
    sys = O.std_system("acrobot")
    ctl = O.std_controller("acrobot_es", sys)
    np.testing.assert_allclose(ctl.K, [[-644.90192792, -257.98492568, -248.79972986, -132.9351164]], rtol=1e-9)
    assert abs(sys.acrobot_energy(np.array([[np.pi, 0, 0, 0]]))[0] - 100.0) < 1e-12
    assert abs(sys.acrobot_energy(np.zeros((1, 4)))[0] + 100.0) < 1e-12
    s2 = O.std_system("quad2d")
    c2 = O.std_controller("quad2d_hover", s2)
    np.testing.assert_allclose(c2.K[0], [-.707107, .707107, 5.36864, -1.128692, 1.098684, 1.357262], atol=2e-6)

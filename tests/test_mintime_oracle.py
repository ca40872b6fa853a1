"""CPU checks of the oracle's minimum-time pieces (examples/double_integrator_optimal_time.ipynb cells 4, 18): structural
properties that do not depend on the reference's data — the exact zero-order-hold step against the double integrator's
closed form, the switching-curve law's odd symmetry and optimality, and the grid policy read from the ANALYTIC value
function agreeing with the switching-curve law away from the curve."""
import numpy as np

from oracle import rollout_oracle as O


def _di(dt=0.01):
    return O.OracleSystem("linear", 2, 1, dt, np.array([-1.0]), np.array([1.0]),
                          {"A": np.array([[0.0, 1.0], [0.0, 0.0]]), "B": np.array([[0.0], [1.0]])})


def _min_time(p, v):
    s = np.where(p > -0.5 * v * np.abs(v), 1.0, -1.0)
    return s * v + 2 * np.sqrt(np.maximum(0.5 * v * v + s * p, 0.0))


def test_discrete_step_is_the_closed_form_of_the_double_integrator():
    rng = np.random.default_rng(0)
    x = rng.uniform(-2, 2, size=(1000, 2))
    u = rng.uniform(-1, 1, size=(1000, 1))
    dt = 0.01
    xn = _di(dt).step(x, u, "discrete")
    np.testing.assert_allclose(xn[:, 1], x[:, 1] + dt * u[:, 0], atol=1e-15)
    np.testing.assert_allclose(xn[:, 0], x[:, 0] + dt * x[:, 1] + 0.5 * dt * dt * u[:, 0], atol=1e-15)
    # inputs beyond the limits are clipped first (Dynamics.simulate, dynamics_basic.py:118)
    big = _di(dt).step(x, 5.0 * np.ones((1000, 1)), "discrete")
    np.testing.assert_allclose(big[:, 1], x[:, 1] + dt, atol=1e-15)


def test_switching_curve_law_is_odd_and_bang_bang():
    rng = np.random.default_rng(1)
    x = rng.uniform(-1.5, 1.5, size=(20000, 2))
    ctl = O.OracleController("switch_curve", metric=1e-4)
    u, um = ctl.control(None, x), ctl.control(None, -x)
    assert set(np.unique(u)) <= {-1.0, 0.0, 1.0}
    away = np.abs(x[:, 0] + 0.5 * x[:, 1] * np.abs(x[:, 1])) > 1e-9     # the notebook's <= / < make the curve itself one-sided
    assert np.array_equal(u[away], -um[away])
    inside = (x ** 2).sum(1) <= 1e-4
    assert (u[inside] == 0).all() and (u[~inside] != 0).all()


def test_switching_curve_law_reaches_the_ball_in_minimum_time():
    rng = np.random.default_rng(2)
    x0 = rng.uniform(-1, 1, size=(2000, 2))
    xs, _, _, _ = O.rollout(_di(), O.OracleController("switch_curve"), x0, 500, "discrete", record_stride=1)
    hit = (xs[1:] ** 2).sum(-1) <= 1e-4
    assert hit.any(0).all()
    t = np.argmax(hit, 0) * 0.01
    d = t - _min_time(x0[:, 0], x0[:, 1])
    assert d.min() > -0.22 and d.max() < 0.45 and 0.04 < d.mean() < 0.12          # ball radius / chattering of the 0.01 s hold


def test_grid_policy_of_the_analytic_value_function_is_the_switching_curve_law():
    pos, vel = np.linspace(-1, 1, 201), np.linspace(-1, 1, 201)
    P, V = np.meshgrid(pos, vel)
    T = _min_time(P, V)                                            # [vel, pos]
    dv = vel[1] - vel[0]
    dTdv = (T[2:, :] - T[:-2, :]) / (2 * dv)
    grid = O.OracleController("grid_sign", grid=dTdv, grid_axes=(pos, vel[1:-1]))
    law = O.OracleController("switch_curve", metric=0.0)
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.95, 0.95, size=(50000, 2))
    far = np.abs(x[:, 0] + 0.5 * x[:, 1] * np.abs(x[:, 1])) > 0.03   # more than a cell and a half from the curve
    ug, ul = grid.control(None, x), law.control(None, x)
    assert far.mean() > 0.9 and (ug[far] == ul[far]).mean() > 0.999

"""Pins the vhjb oracle (torch float64 autograd restatement of controller/vhjb.py): closed-form reverse pass vs
autograd, the LQR fixed point of the HJB residual, Adam / SGDR against their published definitions."""
import numpy as np
import pytest
import scipy.linalg

from oracle import rollout_oracle as O
from oracle import vhjb_oracle as V
from tests.helpers_vhjb import exact_quadratic_weights, problem, sample_batch


@pytest.mark.parametrize("name", ["linear", "cartpole", "cartpole_tanh", "quad2d", "quad10d"])
def test_closed_form_reverse_pass_matches_autograd(name):
    p = problem(name)
    W = V.init_weights(p.sys.n, seed=1)
    if p.act == "tanh":
        pass
    orc = V.VhjbOracle(p, W)
    xs, dones, costs = sample_batch(name, 512, seed=2)
    total, hjb, term, grads, _ = orc.loss_and_grad(xs, dones, costs, reg=0.37)
    cf = orc.closed_form_grads(xs, dones, costs, reg=0.37)
    for g, c in zip(grads, cf):
        assert np.abs(g - c).max() <= 1e-12 * max(1.0, np.abs(g).max())
    assert np.isfinite(total) and hjb > 0 and term > 0


def test_sin_activation_second_derivative_term():
    p = problem("cartpole"); p.act = "sin"
    orc = V.VhjbOracle(p, V.init_weights(4, seed=3))
    xs, dones, costs = sample_batch("cartpole", 256, seed=4)
    _, _, _, grads, _ = orc.loss_and_grad(xs, dones, costs, reg=1.0)
    cf = orc.closed_form_grads(xs, dones, costs, reg=1.0)
    for g, c in zip(grads, cf):
        assert np.abs(g - c).max() <= 1e-12 * max(1.0, np.abs(g).max())


@pytest.mark.parametrize("name", ["linear", "quad10d_linearised"])
def test_lqr_value_function_zeroes_the_hjb_residual(name):
    """With V = z^T P z (P from the Riccati equation) and no input saturation the HJB residual vanishes
    (utils/debug_helper.py:78-102 checks the same identity)."""
    if name == "linear":
        p = problem("linear")
        A, B = p.sys.par["A"], p.sys.par["B"]
    else:
        s = O.std_system("quad10d")
        A, B = O.quad10d_hover_AB(s.par)
        p = V.VhjbProblem(O.OracleSystem("linear", 10, 3, 0.05, -1e9 * np.ones(3), 1e9 * np.ones(3), {"A": A, "B": B}),
                          np.eye(10), np.eye(3), np.zeros(10), np.zeros(3), np.zeros(10), np.ones(10))
    p.sys.umin, p.sys.umax = -1e9 * np.ones(p.sys.m), 1e9 * np.ones(p.sys.m)
    P = scipy.linalg.solve_continuous_are(A, B, p.Q, p.R)
    orc = V.VhjbOracle(p, exact_quadratic_weights(p.sys.n, P, p.eps_s))
    xs = np.random.default_rng(0).uniform(-1, 1, size=(256, p.sys.n))
    q = orc.pieces(xs)
    np.testing.assert_allclose(q["V"].detach().numpy(), np.einsum("bi,ij,bj->b", xs, P, xs), rtol=1e-10)
    np.testing.assert_allclose(q["p"].detach().numpy(), 2 * xs @ P, rtol=1e-9, atol=1e-12)
    assert np.abs(q["r"].detach().numpy()).max() < 1e-7


def test_min_time_form_matches_notebook_definition():
    p = problem("di_mintime")
    orc = V.VhjbOracle(p, V.init_weights(2, seed=5))
    xs, dones, costs = sample_batch("di_mintime", 300, seed=6)
    hjb, term, q = orc.losses(xs, dones, costs)
    u = q["u"].detach().numpy()
    assert set(np.unique(u)).issubset({-1.0, 0.0, 1.0})
    pB = q["p"].detach().numpy() @ p.sys.par["B"]
    np.testing.assert_array_equal(u, -np.sign(pB))
    vdot = (q["p"].detach().numpy() * (xs.astype(np.float64) @ p.sys.par["A"].T + u @ p.sys.par["B"].T)).sum(1)
    assert abs(float(hjb) - np.abs(vdot + costs).mean()) < 1e-12
    assert float(term) == 0.0


def test_adam_and_sgdr_definitions():
    rng = np.random.default_rng(0)
    w = rng.normal(size=50); m = np.zeros(50); v = np.zeros(50)
    import torch
    wt = torch.tensor(w.copy(), requires_grad=True)
    opt = torch.optim.Adam([wt], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for step in range(1, 6):
        g = rng.normal(size=50)
        w, m, v = V.adam_step(w, m, v, g, step)
        wt.grad = torch.tensor(g.copy()); opt.step()
        np.testing.assert_allclose(w, wt.detach().numpy(), rtol=1e-10, atol=1e-12)
    # SGDR (vhjb.py:123-128 with the *.gin values): 0 -> 1e-5 linearly over 1000 steps, cosine back to 0 by 2000
    assert V.sgdr_schedule(0) == 0.0
    assert abs(V.sgdr_schedule(500) - 5e-6) < 1e-18
    assert abs(V.sgdr_schedule(1000) - 1e-5) < 1e-18
    assert abs(V.sgdr_schedule(1500) - 5e-6) < 1e-12
    assert abs(V.sgdr_schedule(2000) - 0.0) < 1e-18 and abs(V.sgdr_schedule(2500) - 5e-6) < 1e-18
    assert V.sgdr_schedule(20000) == 0.0 and V.sgdr_schedule(10 ** 6) == 0.0

"""The soft-PD oracle (oracle/softpd_oracle.py, torch float64 autograd) pinned on the CPU: its parameter gradient against
central finite differences of its own loss, for the three loss forms of the notebooks (cartpole_balancing.ipynb cell 11,
drone_hovering.ipynb cell 11), and the structure of the net (biases, Dense(1) head: V(xf) = head(bias path))."""
import numpy as np
import pytest

from oracle import rollout_oracle as O
from oracle import softpd_oracle as SP


def _problem(kind):
    if kind == "cartpole":
        s = O.std_system("cartpole")
        return SP.SoftPDProblem(s, np.eye(4), np.eye(1), np.array([0, 3.1415926, 0, 0]), np.zeros(1), act="tanh", residual="plain")
    s = O.std_system("quad2d")
    return SP.SoftPDProblem(s, np.eye(6), np.eye(2), np.zeros(6), np.array([4.905, 4.905]), act="relu", residual="normalized")


def _batch(p, B, seed):
    rng = np.random.default_rng(seed)
    span = {4: [2.4, 0.3, 1.0, 1.0], 6: [1, 1, 0.5, 1, 1, 1]}[p.sys.n]
    return p.xf + rng.uniform(-1, 1, size=(B, p.sys.n)) * np.asarray(span)


@pytest.mark.parametrize("kind,form", [("cartpole", "hjb"), ("cartpole", "value_match"), ("quad2d", "hjb_lqr"), ("quad2d", "hjb")])
def test_gradient_matches_finite_differences(kind, form):
    p = _problem(kind)
    params = SP.init_params(p.sys.n, seed=1)
    rng = np.random.default_rng(2)
    params = [w + 0.05 * rng.normal(size=w.shape) for w in params]          # non-zero biases
    xs = _batch(p, 48, seed=3)
    ctl = O.std_controller("cartpole_lqr" if kind == "cartpole" else "quad2d_hover", p.sys)
    K, P = ctl.K, ctl.P
    orc = SP.SoftPDOracle(p, params)
    total, res, hinge, grads = orc.loss_and_grad(xs, form, reg=0.7, K=K, P=P)
    assert np.isfinite(total) and (form == "value_match" or hinge >= 0)
    flat = SP.flat(params)
    g = SP.flat(grads)
    sizes = np.cumsum([w.size for w in params])[:-1]
    picks = np.random.default_rng(4).choice(flat.size, size=12, replace=False)
    for j in list(picks) + [flat.size - 1, flat.size - 2]:               # ... and the head's bias / last weight
        h = 1e-6 * max(1.0, abs(flat[j]))
        vals = []
        for sgn in (+1, -1):
            f2 = flat.copy(); f2[j] += sgn * h
            pr = [a.reshape(w.shape) for a, w in zip(np.split(f2, sizes), params)]
            vals.append(SP.SoftPDOracle(p, pr).loss(xs, form, 0.7, K, P)[0].item())
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - g[j]) <= 2e-5 * max(1e-3, abs(g[j]), np.abs(g).max() * 1e-3), (j, fd, g[j])


def test_value_at_the_goal_is_the_bias_path():
    p = _problem("cartpole")
    params = SP.init_params(4, seed=0)
    params[1][:] = 0.1; params[3][:] = -0.2; params[5][:] = 0.05; params[7][:] = 0.3
    orc = SP.SoftPDOracle(p, params)
    v0 = float(orc.value(SP._t(p.xf[None]))[0])
    h = np.tanh(params[1]); h = np.tanh(h @ params[2] + params[3]); h = np.tanh(h @ params[4] + params[5])
    assert abs(v0 - float((h @ params[6] + params[7])[0])) < 1e-12

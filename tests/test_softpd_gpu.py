"""GPU parity tests for the notebooks' soft-PD baseline (csrc/softpd.cu through the C ABI) against
oracle/softpd_oracle.py (torch float64 autograd): value, input gradient, control; the three losses of
examples/cartpole_balancing.ipynb cell 11 and examples/drone_hovering.ipynb cell 11 and their parameter gradients (weights
AND biases) at the north star's 1e-4; the cart-pole notebook's warm-up reproducing the LQR's closed-loop cost; and the whole
soft-PD experiment of the cart-pole notebook (on-policy loop, 20 warm-up + 80 HJB epochs) against the curve it printed."""
import numpy as np
import pytest

from oracle import rollout_oracle as O
from oracle import softpd_oracle as SP
from tests.helpers import make_dynamics

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _make(kind, seed=1):
    import torch
    assert torch.cuda.is_available()
    from q_learning_with_hjb_b200.controller.soft_pd import SoftPDController
    if kind == "cartpole":
        osys = O.std_system("cartpole")
        prob = SP.SoftPDProblem(osys, np.eye(4), np.eye(1), np.array([0, 3.1415926, 0, 0]), np.zeros(1), act="tanh", residual="plain")
        octl = O.std_controller("cartpole_lqr", osys)
    else:
        osys = O.std_system("quad2d")
        prob = SP.SoftPDProblem(osys, np.eye(6), np.eye(2), np.zeros(6), np.array([4.905, 4.905]), act="relu", residual="normalized")
        octl = O.std_controller("quad2d_hover", osys)
    dyn = make_dynamics(kind)
    ctl = SoftPDController(dyn, prob.xf, prob.uf, prob.Q, prob.R, activation=prob.act,
                           normalized_residual=prob.residual == "normalized", K=octl.K, P=octl.P, seed=seed)
    # non-zero biases (Flax initialises them to zero: every bias path would be untested)
    rng = np.random.default_rng(seed + 10)
    ctl.params += torch.as_tensor(0.05 * rng.normal(size=ctl.params.numel()).astype(np.float32)).cuda()
    flat = ctl.params.cpu().numpy().astype(np.float64)
    shapes = [w.shape for w in SP.init_params(osys.n)]
    sizes = np.cumsum([int(np.prod(s)) for s in shapes])[:-1]
    params = [a.reshape(s) for a, s in zip(np.split(flat, sizes), shapes)]
    return ctl, prob, SP.SoftPDOracle(prob, params), octl


def _batch(prob, B, seed):
    rng = np.random.default_rng(seed)
    span = {4: [2.4, 0.3, 1.0, 1.0], 6: [1, 1, 0.5, 1, 1, 1]}[prob.sys.n]
    return (prob.xf + rng.uniform(-1, 1, size=(B, prob.sys.n)) * np.asarray(span)).astype(np.float32)


@pytest.mark.parametrize("kind", ["cartpole", "quad2d"])
def test_value_gradient_and_control_match_oracle(kind):
    ctl, prob, orc, _ = _make(kind)
    xs = _batch(prob, 3001, seed=2)
    u, V, p = ctl.get_control_efforts_with_additional_term(xs)
    q = orc.pieces(xs.astype(np.float64))
    for got, want in ((V, q["V"]), (p, q["p"]), (u, q["u"])):
        want = want.detach().numpy()
        assert np.abs(got.cpu().numpy() - want).max() <= TOL * max(1.0, np.abs(want).max())
    u1 = ctl.get_control_efforts(xs[5])
    assert u1.shape == (prob.sys.m,) and np.abs(u1 - q["u"][5].detach().numpy()).max() < 1e-4


@pytest.mark.parametrize("kind,form", [("cartpole", "hjb"), ("cartpole", "value_match"), ("quad2d", "hjb"), ("quad2d", "hjb_lqr")])
@pytest.mark.parametrize("B", [77, 5000])
def test_losses_and_parameter_gradients_match_oracle(kind, form, B):
    ctl, prob, orc, octl = _make(kind, seed=3)
    xs = _batch(prob, B, seed=4)
    total, res, hinge = ctl.loss_grad(xs, form, regularization=0.7)
    t_o, r_o, h_o, grads = orc.loss_and_grad(xs.astype(np.float64), form, reg=0.7, K=octl.K, P=octl.P)
    assert abs(float(total) - t_o) <= TOL * abs(t_o) and abs(float(res) - r_o) <= TOL * abs(r_o)
    assert abs(float(hinge) - h_o) <= TOL * max(abs(h_o), 1e-3)
    g = ctl.grad.cpu().numpy().astype(np.float64)
    off = 0
    for name, go in zip(("W1", "b1", "W2", "b2", "W3", "b3", "w4", "b4"), grads):
        sl = slice(off, off + go.size); off += go.size
        # (floor: b4-bar is an exact cancellation, -reg count / B from the states against +reg count / B from V(xf))
        scale = max(np.abs(go).max(), 1e-2 * np.abs(SP.flat(grads)).max())
        assert np.abs(g[sl] - go.reshape(-1)).max() <= TOL * scale, name
    assert off == g.size
    # deterministic: the same bits twice
    g1 = ctl.grad.clone()
    ctl.loss_grad(xs, form, regularization=0.7)
    import torch
    assert torch.equal(g1, ctl.grad)


def test_cartpole_warmup_reproduces_the_lqr_cost():
    """cartpole_balancing.ipynb cells 11 and 16: after the warm-up on |V - z^T P z| the soft-PD net's policy
    u = clip(-R^-1 g^T dV/dx / 2) balances the cart-pole at close to the LQR's cost — the notebook prints 9.1592 for the
    trained soft-PD policy against 9.1410 for the LQR on its ten evaluation states (same NumPy draw offset as the K-CP known
    answer); 6000 warm-up updates on uniformly sampled states get within 25 %."""
    import torch
    from q_learning_with_hjb_b200.controller.soft_pd import SoftPDController
    dyn = make_dynamics("cartpole")
    osys = O.std_system("cartpole")
    octl = O.std_controller("cartpole_lqr", osys)
    xf = np.array([0, 3.1415926, 0, 0])
    ctl = SoftPDController(dyn, xf, np.zeros(1), np.eye(4), np.eye(1), activation="tanh", K=octl.K, P=octl.P, seed=0)
    rng = np.random.default_rng(0)
    first = last = None
    for it in range(6000):
        xs = (xf + rng.uniform(-1, 1, size=(256, 4)) * [2.4, 0.25, 1.5, 1.0]).astype(np.float32)
        loss, _, _ = ctl.params_update(xs, "value_match")
        if it == 0:
            first = float(loss)
    last = float(loss)
    assert last < 0.1 * first, (first, last)
    # (the warm-up is what makes this baseline work at all: the notebooks print 82.4 / 221.4 for the soft-PD drone policy
    # without it, and its HJB phase is run on on-policy data; here only the warm-up is reproduced)
    # the notebook's ten evaluation states: np.random.seed(0), 6001 draws skipped (SURVEY.md section 4, K-CP)
    np.random.seed(0)
    for _ in range(6001):
        np.random.uniform(size=(4,), low=-dyn.x0_std, high=dyn.x0_std)
    x = torch.as_tensor(dyn.get_initial_states(10).astype(np.float32)).cuda()
    cost = torch.zeros(10, device="cuda")
    xf_d = torch.as_tensor(xf.astype(np.float32)).cuda()
    for _ in range(500):
        u, _, _ = ctl.get_control_efforts_with_additional_term(x)
        dx = x - xf_d
        dx[:, 1] = torch.remainder(dx[:, 1] + np.pi, 2 * np.pi) - np.pi
        cost += ((dx * dx).sum(1) + (u * u).sum(1)) * float(dyn.dt)
        x = dyn.simulate(x, u)
    mean_cost = float(cost.mean())
    assert abs(mean_cost - 9.140986134043468) < 0.25 * 9.14, mean_cost      # notebook: soft-PD 9.1592, LQR 9.1410
    before = ctl.params.clone()
    total, res, hinge = ctl.params_update(xs, "hjb", regularization=1.0)       # one step of the notebook's HJB loss + hinge
    assert torch.isfinite(ctl.params).all() and not torch.equal(before, ctl.params) and float(total) >= float(res) >= 0


# examples/cartpole_balancing.ipynb cell 11 output: (loss, cumulated cost, collected length) at epochs 10, 20, ..., 90
NOTEBOOK_SOFT_PD = [(18.924827575683594, 81.99, 69.75), (2.000885248184204, 5.657, 200.0), (0.43753620982170105, 5.310, 200.0),
                    (0.27048197388648987, 6.885, 200.0), (0.18633385002613068, 5.788, 200.0), (0.14706996083259583, 6.061, 200.0),
                    (0.12321638315916061, 5.440, 200.0), (0.10425637662410736, 5.721, 200.0), (0.09072544425725937, 7.442, 200.0)]


def test_cartpole_soft_pd_experiment_follows_the_notebook_curve():
    """The notebook-curve pin of SURVEY.md 8f row 3: examples/cartpole_balancing.py reruns cell 11 (20 on-policy
    trajectories per epoch into the data set, one shuffled pass of minibatches of 256; 20 epochs on |V - z^T P z|, then
    |vdot + l| + max(0, V(xf) - V)) on the CUDA soft-PD kernels.  JAX's PRNG is not reproducible here and the method is
    sensitive to the initialisation (it has no positive-definiteness guarantee: of the seeds 0..3, two keep every trajectory
    in the box, as the notebook's run does, two lose some after epoch 40), so the pin is the curve of a seed that trains:
    warm-up loss within 15 % at epoch 10 (18.92: the net has not moved yet, the loss is the data's z^T P z), balanced
    full-length trajectories from epoch 20 on with cumulated costs in the notebook's 5.3-7.5 band, HJB-phase losses that fall
    monotonically along the printed ones (the same curve about ten epochs behind: 1.5 x the printed value at epochs 30-50,
    1.2 x from epoch 60 on), and a closed-loop cost within 2 % of the LQR's on fresh initial states (notebook cell 16:
    9.159 against 9.141, 0.2 %)."""
    import os
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import cartpole_balancing as C
    import onpolicy_hjb as H
    p, k = C.make_problem()
    for _ in range(2001):            # the notebook reaches cell 11 after 1 + 100 x 20 draws of its initial-state stream
        p.dyn.get_initial_state()
    soft = C.make_soft_pd(p, seed=1)
    hist = np.array(H.train_soft_pd(p, soft, 100, warmup_epochs=20, warmup_form="value_match", regularization=1.0, seed=1, log=None))
    at = hist[9:90:10]
    ref = np.array(NOTEBOOK_SOFT_PD)
    assert abs(at[0, 0] - ref[0, 0]) < 0.15 * ref[0, 0], at[0]
    assert at[0, 2] < 120 and at[0, 1] > 30, at[0]                     # still falling over at warm-up epoch 10 (69.75 / 82.0)
    assert (at[1:, 2] == 200).all(), at[:, 2]                           # every trajectory stays in the box from epoch 20 on
    assert (at[1:, 1] > 4.0).all() and (at[1:, 1] < 9.0).all(), at[:, 1]
    ratio = at[2:, 0] / ref[2:, 0]                                      # measured: 1.47 1.47 1.56 1.27 1.21 1.23 1.17
    assert (ratio < 1.75).all() and (ratio > 1 / 1.75).all() and ratio[-1] < 1.35, (at[:, 0], ref[:, 0])
    assert (np.diff(at[1:, 0]) < 0).all(), at[:, 0]
    assert (hist[20:, 2] == 200).all() and hist[20:, 1].max() < 12.0
    x0 = np.stack([p.dyn.get_initial_state() for _ in range(10)])
    from q_learning_with_hjb_b200.controller.cartpole_energy_shaping import CartpoleEnergyShapingController
    K, _ = CartpoleEnergyShapingController(p.dyn).get_lqr_term()
    steps = int(round(10 / p.dyn.dt))
    c_soft = H.closed_loop_cost(p, H.soft_pd_policy(soft), x0, steps).mean()
    c_lqr = H.closed_loop_cost(p, H.lqr_policy(p, K), x0, steps).mean()
    assert abs(c_soft - c_lqr) < 0.02 * c_lqr, (c_soft, c_lqr)

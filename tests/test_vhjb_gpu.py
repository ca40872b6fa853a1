"""GPU parity tests for hot path (b): the fused vhjb kernels (through the ctypes C ABI) against the torch-float64
oracle (oracle/vhjb_oracle.py).  Tolerance (north_star): HJB residual, loss and gradient within 1e-4.

Per-sample quantities are discontinuous functions of the parameters (ReLU kinks move dV/dx by a finite jump, the
input clip and |.| have kinks): a state whose pre-activation lies within fp32 rounding of a kink may land on the
other side than float64 does.  Such samples are counted as outliers against a stated budget; losses and gradients
(sums over the batch) are compared without any budget.
"""
import numpy as np
import pytest

from oracle import vhjb_oracle as V
from tests.helpers_vhjb import exact_quadratic_weights, flat_params, make_kernels, problem, sample_batch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def _dev(torch, *arrays):
    return [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda() for a in arrays]


def _setup(name, B, seed=0, wseed=0):
    torch = _cuda()
    k, p = make_kernels(name)
    W32 = [w.astype(np.float32) for w in V.init_weights(p.sys.n, seed=wseed)]
    xs, dones, costs = sample_batch(name, B, seed=seed)
    orc = V.VhjbOracle(p, [w.astype(np.float64) for w in W32])
    params = torch.as_tensor(flat_params(W32)).cuda()
    return torch, k, p, orc, params, xs, dones, costs


NAMES = ["linear", "cartpole", "cartpole_tanh", "quad2d", "quad10d", "di_mintime", "linear_sin", "quad2d_tanh", "quad10d_sin"]


@pytest.mark.parametrize("name", NAMES)
def test_residual_pieces_match_oracle(name):
    B = 20000 + 17                                   # not a multiple of the 32-state tile
    torch, k, p, orc, params, xs, dones, costs = _setup(name, B)
    xd, dd, cd = _dev(torch, xs, dones, costs)
    out, sums = k.residual(params, xd, dd, cd)
    q = orc.pieces(xs, running=costs if p.residual_form == "min_time" else None)
    ref = {key: q[key].detach().numpy() for key in ("V", "p", "u", "r")}
    got = {key: out[key].cpu().numpy().astype(np.float64) for key in ("V", "p", "u", "r")}
    budget = {"V": 0.0, "p": 2e-3, "u": 2e-3, "r": 2e-3} if p.act == "relu" else {"V": 0, "p": 1e-4, "u": 1e-4, "r": 1e-4}
    for key in ("V", "p", "u", "r"):
        a, b = got[key].reshape(B, -1), ref[key].reshape(B, -1)
        scale = np.maximum(np.abs(b), np.abs(b).mean(axis=0, keepdims=True) + 1e-30)
        err = (np.abs(a - b) / scale).max(axis=1)
        frac = float(np.mean(err > TOL))
        assert frac <= budget[key], (key, frac, float(np.median(err)))
        assert np.median(err) < 1e-5
    # un-normalised loss sums
    hjb, term, _ = orc.losses(xs, dones, costs)
    d = dones.astype(np.float64)
    if p.residual_form == "normalized":
        assert abs(float(sums[0]) / ((1 - d).sum() + p.eps) - float(hjb)) < TOL * float(hjb)
        assert abs(float(sums[1]) / (d.sum() + p.eps) - float(term)) < TOL * float(term)
    else:
        assert abs(float(sums[0]) / B - float(hjb)) < TOL * float(hjb)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("B", [256, 4096 + 5])
def test_loss_and_gradient_match_oracle(name, B):
    torch, k, p, orc, params, xs, dones, costs = _setup(name, B, seed=3, wseed=2)
    xd, dd, cd = _dev(torch, xs, dones, costs)
    reg = 0.37
    if p.residual_form == "min_time":
        k.norm.copy_(torch.tensor([float(B), 1.0]))
    else:
        k.counts(dd, p.eps)
        d = dones.astype(np.float64)
        np.testing.assert_allclose(k.norm.cpu().numpy(), [(1 - d).sum() + p.eps, d.sum() + p.eps], rtol=1e-6)
    grad, sums = k.loss_grad(params, xd, dd, cd, reg)
    total, hjb, term, grads, _ = orc.loss_and_grad(xs, dones, costs, reg)
    g = grad.cpu().numpy().astype(np.float64)
    go = np.concatenate([x.reshape(-1) for x in grads])
    assert g.shape == go.shape
    off = 0
    for gi in grads:                                   # per-matrix normwise error
        sl = slice(off, off + gi.size); off += gi.size
        assert np.abs(g[sl] - go[sl]).max() <= TOL * np.abs(go[sl]).max(), (name, gi.shape)
    norm = k.norm.cpu().numpy().astype(np.float64)
    s = sums.cpu().numpy().astype(np.float64)
    assert abs(s[0] / norm[0] - hjb) <= TOL * hjb
    if p.residual_form == "normalized":
        assert abs(s[1] / norm[1] - term) <= TOL * term
        assert abs(s[0] / norm[0] + reg * s[1] / norm[1] - total) <= TOL * total


def test_lqr_fixed_point_through_cuda():
    """V = z^T P z represented exactly by the relu net => the HJB residual vanishes (no saturation)."""
    import scipy.linalg
    torch = _cuda()
    k, p = make_kernels("linear")
    A, Bm = p.sys.par["A"], p.sys.par["B"]
    P = scipy.linalg.solve_continuous_are(A, Bm, p.Q, p.R)
    params = torch.as_tensor(flat_params(exact_quadratic_weights(2, P, p.eps_s))).cuda()
    xs = np.random.default_rng(0).uniform(-1, 1, size=(4096, 2)).astype(np.float32)
    xs = xs[np.abs(xs @ P @ Bm).ravel() < 4.9]          # |u| = |B^T P x| < umax = 5: unsaturated
    xd, dd, cd = _dev(torch, xs, np.zeros(len(xs)), np.ones(len(xs)))
    out, _ = k.residual(params, xd, dd, cd)
    np.testing.assert_allclose(out["V"].cpu().numpy(), np.einsum("bi,ij,bj->b", xs, P, xs), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(out["p"].cpu().numpy(), 2 * xs @ P, rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(out["u"].cpu().numpy(), -(xs @ P @ Bm), rtol=2e-5, atol=1e-5)
    assert np.abs(out["r"].cpu().numpy()).max() < 2e-4


def test_adam_matches_optax_definition():
    torch = _cuda()
    k, p = make_kernels("linear")
    rng = np.random.default_rng(0)
    n = k.P
    w = rng.normal(size=n).astype(np.float32); m = np.zeros(n); v = np.zeros(n)
    wd = torch.as_tensor(w.copy()).cuda(); md = torch.zeros(n, device="cuda"); vd = torch.zeros(n, device="cuda")
    w64 = w.astype(np.float64)
    for step in range(1, 6):
        g = (rng.normal(size=n) * 10.0 ** rng.integers(-6, 2)).astype(np.float32)
        k.adam(wd, md, vd, torch.as_tensor(g).cuda(), step, 1e-3)
        w64, m, v = V.adam_step(w64, m, v, g.astype(np.float64), step)
        np.testing.assert_allclose(wd.cpu().numpy(), w64, rtol=1e-5, atol=1e-7)


def test_deterministic_and_shard_additive():
    """Bitwise reproducible; and the gradient of a batch equals the sum of the gradients of its two shards when both
    use the global normaliser — what the multi-GPU all-reduce relies on."""
    torch, k, p, orc, params, xs, dones, costs = _setup("quad10d", 8192, seed=5)
    xd, dd, cd = _dev(torch, xs, dones, costs)
    k.counts(dd, p.eps)
    g1 = k.loss_grad(params, xd, dd, cd, 0.2)[0].clone()
    s1 = k.sums.clone()
    g2 = k.loss_grad(params, xd, dd, cd, 0.2)[0].clone()
    assert torch.equal(g1, g2) and torch.equal(s1, k.sums)
    h = 4096 + 96
    ga = k.loss_grad(params, xd[:h].contiguous(), dd[:h].contiguous(), cd[:h].contiguous(), 0.2)[0].clone()
    sa = k.sums.clone()
    gb = k.loss_grad(params, xd[h:].contiguous(), dd[h:].contiguous(), cd[h:].contiguous(), 0.2)[0].clone()
    sb = k.sums.clone()
    assert (ga + gb - g1).abs().max() <= 2e-5 * g1.abs().max()
    assert ((sa + sb) - s1).abs().max() <= 1e-5 * s1.abs().max()


def test_controller_params_update_matches_oracle_and_trains():
    """VHJBController (reference interface, cartpole gin files): one params_update == oracle loss/grad + Adam; a few
    epochs of train() run and return the reference's six lists."""
    import os
    torch = _cuda()
    from q_learning_with_hjb_b200.configs import gin_compat as gin
    from q_learning_with_hjb_b200.configs.controller.vhjb_controller_config import VHJBControllerConfig
    from q_learning_with_hjb_b200.controller.vhjb import VHJBController
    from tests.helpers import PKG, make_dynamics
    dyn = make_dynamics("cartpole")
    gin.parse_config_file(os.path.join(PKG, "configs", "controller", "cartpole_vhjb_controller.gin"))
    cfg = VHJBControllerConfig()
    cfg.epochs, cfg.num_of_trajectories_per_epoch, cfg.maximum_step = 2, 3, 30
    ctl = VHJBController(dyn, cfg)
    assert len(ctl.replay_buffer) == 20 and ctl.P.shape == (4, 4)
    W32 = [kk.cpu().numpy().copy() for kk in ctl.model_params.kernels()]
    xs, dones, costs = sample_batch("cartpole", 256, seed=9)
    p = problem("cartpole")
    orc = V.VhjbOracle(p, [w.astype(np.float64) for w in W32])
    total, hjb, term, grads, _ = orc.loss_and_grad(xs, dones, costs, 1e-5)
    params, states, opt, t, h, tm = ctl.params_update(ctl.model_params, ctl.model_states, ctl.optimizer_states,
                                                      xs, dones, costs, 1e-5)
    assert abs(float(h) - hjb) <= TOL * hjb and abs(float(tm) - term) <= TOL * term and abs(float(t) - total) <= TOL * total
    assert opt.count == 1 and states == {}
    off = 0
    for w, g in zip(W32, grads):   # first Adam step moves each weight by -lr * sign(g) (up to eps)
        new = params.flat[off:off + w.size].cpu().numpy().reshape(w.shape); off += w.size
        wn, _, _ = V.adam_step(w.astype(np.float64), 0, 0, g, 1)
        # (elements whose gradient is below the 1e-4 gradient tolerance have an undetermined sign: not compared)
        big = np.abs(g) > TOL * np.abs(g).max()
        np.testing.assert_allclose(new[big], wn[big], rtol=0, atol=2e-6)
    # single-state interface
    x = dyn.get_initial_state()
    u = ctl.get_control_efforts(x)
    assert u.shape == (1,)
    uq = V.VhjbOracle(p, [kk.cpu().numpy().astype(np.float64) for kk in ctl.model_params.kernels()]).pieces(x[None])["u"]
    assert abs(u[0] - float(uq[0, 0])) < 1e-3 * max(1.0, abs(float(uq[0, 0])))
    lists = ctl.train()
    assert len(lists) == 6 and len(lists[0]) == 2 and all(np.isfinite(v) for v in lists[3])


def test_full_size_batch_is_consistent_with_its_chunks():
    """C5 per-GPU size (1,048,576 states, n = 10): loss sums and gradient equal the sum over 16 chunks."""
    torch, k, p, orc, params, xs, dones, costs = _setup("quad10d", 1 << 20, seed=11)
    xd, dd, cd = _dev(torch, xs, dones, costs)
    k.counts(dd, p.eps)
    g = k.loss_grad(params, xd, dd, cd, 0.1)[0].clone()
    s = k.sums.clone()
    acc_g, acc_s = torch.zeros_like(g), torch.zeros_like(s)
    for c in range(16):
        sl = slice(c << 16, (c + 1) << 16)
        acc_g += k.loss_grad(params, xd[sl].contiguous(), dd[sl].contiguous(), cd[sl].contiguous(), 0.1)[0]
        acc_s += k.sums
    assert torch.isfinite(g).all() and g.abs().max() > 0
    assert (acc_g - g).abs().max() <= 5e-5 * g.abs().max()
    assert (acc_s - s).abs().max() <= 5e-5 * s.abs().max()
    # and a 65,536-state chunk agrees with the float64 oracle
    sl = slice(0, 1 << 16)
    norm = k.norm.cpu().numpy().astype(np.float64)
    out, sums = k.residual(params, xd[sl].contiguous(), dd[sl].contiguous(), cd[sl].contiguous(), want=())
    hjb, term, _ = orc.losses(xs[sl], dones[sl], costs[sl])
    d = dones[sl].astype(np.float64)
    assert abs(float(sums[0]) / ((1 - d).sum() + p.eps) - float(hjb)) <= TOL * float(hjb)


# ---- the two kernels behind the same C ABI: tcgen05 (default) and CUDA-core fp32 (HJB_VHJB_IMPL=simt) ----
def _with_impl(impl, fn):
    import os
    old = os.environ.get("HJB_VHJB_IMPL")
    try:
        if impl is None:
            os.environ.pop("HJB_VHJB_IMPL", None)
        else:
            os.environ["HJB_VHJB_IMPL"] = impl
        return fn()
    finally:
        if old is None:
            os.environ.pop("HJB_VHJB_IMPL", None)
        else:
            os.environ["HJB_VHJB_IMPL"] = old


@pytest.mark.parametrize("name", ["linear", "quad10d", "cartpole_tanh", "di_mintime", "quad2d_sin"])
def test_tensor_core_and_cuda_core_kernels_agree(name):
    """Same inputs through vhjb_tc.cuh (fp16x3 on tcgen05) and vhjb_simt.cuh (fp32 FMAs): per-state outputs, loss sums
    and gradient agree within the parity tolerance (they are two independent implementations of SURVEY.md 8a-V1..V6)."""
    B = 12000 + 37
    torch, k, p, orc, params, xs, dones, costs = _setup(name, B, seed=21, wseed=4)
    xd, dd, cd = _dev(torch, xs, dones, costs)
    k.counts(dd, p.eps)

    def run():
        out, sums = k.residual(params, xd, dd, cd)
        res = {key: v.clone() for key, v in out.items()}
        rs = sums.clone()
        g = k.loss_grad(params, xd, dd, cd, 0.25)[0].clone()
        return res, rs, g, k.sums.clone()

    rt, st, gt, s2t = _with_impl(None, run)
    rc, sc, gc, s2c = _with_impl("simt", run)
    assert not torch.equal(gt, gc)                       # really two different kernels
    assert (gt - gc).abs().max() <= TOL * gc.abs().max()
    assert ((st - sc).abs() <= TOL * sc.abs() + 1e-30).all() and ((s2t - s2c).abs() <= TOL * s2c.abs() + 1e-30).all()
    for key in ("V", "p", "u", "r"):
        a, b = rt[key].reshape(B, -1).double(), rc[key].reshape(B, -1).double()
        scale = torch.maximum(b.abs(), b.abs().mean(dim=0, keepdim=True) + 1e-30)
        err = ((a - b).abs() / scale).max(dim=1).values
        assert float((err > TOL).double().mean()) <= 2e-3 and float(err.median()) < 1e-5, key


def test_near_goal_states_keep_their_weight():
    """States a distance 1e-3 from the goal carry adjoint seeds thousands of times the batch-typical ones (1/(l + eps)):
    the tensor-core kernel's per-state power-of-two scaling keeps them exact — gradient within tolerance, none counted
    as saturated."""
    B = 4096
    torch, k, p, orc, params, xs, dones, costs = _setup("linear", B, seed=8, wseed=1)
    rng = np.random.default_rng(0)
    idx = rng.choice(B, size=9, replace=False)
    xs[idx] = (p.xf + 1e-3 * rng.uniform(-1, 1, size=(9, p.sys.n))).astype(np.float32)
    dones[idx] = 0
    xd, dd, cd = _dev(torch, xs, dones, costs)
    k.counts(dd, p.eps)
    grad = k.loss_grad(params, xd, dd, cd, 0.37)[0]
    assert k.saturated() == 0
    _, _, _, grads, _ = orc.loss_and_grad(xs, dones, costs, 0.37)
    g = grad.cpu().numpy().astype(np.float64)
    off = 0
    for gi in grads:
        sl = slice(off, off + gi.size); off += gi.size
        assert np.abs(g[sl] - gi.reshape(-1)).max() <= TOL * np.abs(gi).max()


def test_out_of_range_seeds_take_the_fp32_pass():
    """A terminal sample whose stored cost is 0 has the weight 1/(0 + eps) = 1e10: beyond the fp16 range management of
    the tensor-core kernel.  Its index goes to the deferred list and the fp32 CUDA-core pass behind the tensor kernel
    computes it: the gradient of the batch is exact (round 1 clipped and counted such a state), nothing is saturated."""
    B = 2048
    torch, k, p, orc, params, xs, dones, costs = _setup("quad10d", B, seed=13, wseed=3)
    dones[:] = 0
    dones[5] = 1
    costs[5] = 0.0
    xd, dd, cd = _dev(torch, xs, dones, costs)
    k.counts(dd, p.eps)
    g_tc = _with_impl(None, lambda: k.loss_grad(params, xd, dd, cd, 0.5)[0].clone())
    assert torch.isfinite(g_tc).all()
    assert _with_impl(None, k.saturated) == 0 and _with_impl(None, k.deferred) >= 1
    g_cc = _with_impl("simt", lambda: k.loss_grad(params, xd, dd, cd, 0.5)[0].clone())
    assert _with_impl("simt", k.saturated) == 0
    _, _, _, grads, _ = orc.loss_and_grad(xs, dones, costs, 0.5)
    go = np.concatenate([x.reshape(-1) for x in grads])
    assert np.abs(g_cc.cpu().numpy() - go).max() <= TOL * np.abs(go).max()
    assert np.abs(g_tc.cpu().numpy() - go).max() <= TOL * np.abs(go).max()


@pytest.mark.parametrize("name,B", [("linear", 256), ("quad10d", 20000), ("linear_sin", 4099)])
def test_deferred_states_are_exact_deterministic_and_counted(name, B):
    """A third of the batch within 3e-5 of the goal: those states leave the tensor path for the fp32 pass — their adjoint
    seeds are > 2^6 x typical (an fp64 emulation of the seed exponents: 66 % / 100 % / 19 % of them in the three cases;
    in the double integrator the two terms of p-bar largely cancel), and their inputs are below 2^-10, where the fp16
    split of the activations is good to 3e-5 relative at best.  Loss sums and gradient against the
    oracle at the usual tolerance; two runs give the same bits (the deferred lists are filled in tile order per epilogue
    warp, and their partials are summed in a fixed order); the count says how many states went that way."""
    torch, k, p, orc, params, xs, dones, costs = _setup(name, B, seed=21, wseed=4)
    rng = np.random.default_rng(5)
    near = rng.choice(B, size=B // 3, replace=False)
    xs[near] = (p.xf + 3e-5 * rng.uniform(-1, 1, size=(len(near), p.sys.n))).astype(np.float32)
    dones[near] = 0
    xd, dd, cd = _dev(torch, xs, dones, costs)
    k.counts(dd, p.eps)
    g1, s1 = (t.clone() for t in k.loss_grad(params, xd, dd, cd, 0.3))
    nd = k.deferred()
    assert len(near) <= nd <= B and k.saturated() == 0
    g2, s2 = (t.clone() for t in k.loss_grad(params, xd, dd, cd, 0.3))
    assert torch.equal(g1, g2) and torch.equal(s1, s2)
    total, hjb, term, grads, _ = orc.loss_and_grad(xs, dones, costs, 0.3)
    go = np.concatenate([x.reshape(-1) for x in grads])
    assert np.abs(g1.cpu().numpy() - go).max() <= TOL * np.abs(go).max()
    assert abs(float(s1[0] / k.norm[0]) - hjb) <= TOL * abs(hjb)
    # a batch without such states defers nothing
    xs2, dones2, costs2 = sample_batch(name, B, seed=22)
    far = np.linalg.norm(xs2 - p.xf, axis=1) > 0.5
    x2, d2, c2 = _dev(torch, xs2[far], dones2[far] * 0, costs2[far])
    k.counts(d2, p.eps)
    k.loss_grad(params, x2, d2, c2, 0.0)
    assert k.deferred() == 0


# ---- learned-policy rollouts, all trajectories at once on device ----
def _cartpole_controller(max_step):
    import os
    from q_learning_with_hjb_b200.configs import gin_compat as gin
    from q_learning_with_hjb_b200.configs.controller.vhjb_controller_config import VHJBControllerConfig
    from q_learning_with_hjb_b200.controller.vhjb import VHJBController
    from tests.helpers import PKG, make_dynamics
    dyn = make_dynamics("cartpole")
    gin.parse_config_file(os.path.join(PKG, "configs", "controller", "cartpole_vhjb_controller.gin"))
    cfg = VHJBControllerConfig()
    cfg.epochs, cfg.num_of_trajectories_per_epoch, cfg.maximum_step = 1, 4, max_step
    return dyn, VHJBController(dyn, cfg)


def test_batched_policy_rollout_matches_the_per_trajectory_loop():
    """rollout_trajectories (one residual launch + one hjb_policy_step launch per step for ALL trajectories) produces
    the samples of rollout_trajectory (the reference's per-trajectory loop, vhjb.py:171-193) for the same initial states:
    same lengths and done flags, states / costs within 1e-5."""
    _cuda()
    T = 40
    dyn, ctl = _cartpole_controller(T)
    N = 7
    x0s = np.stack([dyn.get_initial_state() for _ in range(N)])
    x0s[3, 0] = 4.75                                   # leaves the observation box (|p| <= 4.8) within a few steps
    x0s[3, 2] = 3.0
    x0s[5, 0] = 5.5                                    # starts outside: a single terminal sample
    rx, rc, rd, total = ctl.rollout_trajectories(x0s)
    rx, rc, rd, total = rx.cpu().numpy(), rc.cpu().numpy(), rd.cpu().numpy(), total.cpu().numpy()
    it = iter(x0s)
    dyn.get_initial_state = lambda: next(it).copy()
    for e in range(N):
        traj = ctl.rollout_trajectory()
        L = len(traj)
        assert (rd[:L, e] >= 0).all() and (rd[L:, e] < 0).all(), (e, L)
        xs = np.stack([t[0] for t in traj]); cs = np.array([t[1] for t in traj]); ds = np.array([t[2] for t in traj])
        np.testing.assert_array_equal(rd[:L, e], ds)
        scale = np.abs(xs).max(axis=0) + 1e-6
        assert (np.abs(rx[:L, e] - xs) / scale).max() < 1e-5
        np.testing.assert_allclose(rc[:L, e], cs, rtol=2e-5, atol=1e-6)
        assert abs(total[e] - ctl.get_trajectory_cost(traj)) <= 2e-5 * abs(ctl.get_trajectory_cost(traj)) + 1e-6
    assert rd[0, 5] == 1 and rd[1, 5] < 0              # out of the box at the first check
    assert 1 < int((rd[:, 3] >= 0).sum()) < T + 1      # left the box on the way


def test_batched_policy_rollout_matches_oracle_steps():
    """The first steps of the batched learned-policy rollout against the oracles: u from the torch-fp64 value-net oracle,
    x' from the rollout oracle (the reference's Dynamics.simulate), sample cost l(x, u) dt."""
    _cuda()
    from oracle import rollout_oracle as O
    T = 6
    dyn, ctl = _cartpole_controller(T)
    p = problem("cartpole")
    W = [kk.cpu().numpy().astype(np.float64) for kk in ctl.model_params.kernels()]
    orc = V.VhjbOracle(p, W)
    osys = O.std_system("cartpole")
    x0s = np.stack([dyn.get_initial_state() for _ in range(5)])
    rx, rc, rd, _ = ctl.rollout_trajectories(x0s)
    rx, rc = rx.cpu().numpy().astype(np.float64), rc.cpu().numpy().astype(np.float64)
    x = x0s.astype(np.float32).astype(np.float64)
    for t in range(T):
        assert np.abs(rx[t] - x).max() < 2e-5 * (t + 1)
        u = orc.pieces(rx[t])["u"].detach().numpy()              # policy at the device trajectory's own state
        z = osys.wrap(rx[t] - p.xf)
        l = np.einsum("bi,ij,bj->b", z, p.Q, z) + np.einsum("bi,ij,bj->b", u - p.uf, p.R, u - p.uf)
        np.testing.assert_allclose(rc[t], l * osys.dt, rtol=1e-4, atol=1e-7)
        x = osys.step(rx[t], u)


@pytest.mark.parametrize("name", ["quad10d", "di_mintime"])
def test_host_batch_train_step_matches_device_batch(name):
    """VhjbKernels.train_step_host (pinned host batch, copies pipelined under the kernel, gradient accumulated piece by
    piece with the whole batch's normalisers) against train_step on the same batch already on the device: identical for
    one piece, and equal up to fp32 summation order for several."""
    from q_learning_with_hjb_b200.controller.vhjb import AdamState
    B = 4 * 32768 + 4321
    torch, k, p, orc, params, xs, dones, costs = _setup(name, B, seed=11, wseed=5)
    xd, dd, cd = _dev(torch, xs, dones, costs)
    host = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).pin_memory() for a in (xs, dones, costs)]

    def run(fn):
        w = params.clone()
        opt = AdamState(0, torch.zeros_like(w), torch.zeros_like(w))
        sums, norm = fn(w, opt)
        return w, k.grad.clone(), sums.clone(), norm.clone()

    w0, g0, s0, n0 = run(lambda w, opt: k.train_step(w, opt, xd, dd, cd, 0.3, 1e-3))
    w1, g1, s1, n1 = run(lambda w, opt: k.train_step_host(w, opt, host[0], host[1], host[2], 0.3, 1e-3, chunks=1))
    assert torch.equal(g0, g1) and torch.equal(w0, w1) and torch.equal(s0, s1) and torch.equal(n0, n1)
    for chunks in (2, 4):
        w2, g2, s2, n2 = run(lambda w, opt: k.train_step_host(w, opt, host[0], host[1], host[2], 0.3, 1e-3, chunks=chunks))
        assert torch.equal(n0, n2)
        assert (g2 - g0).abs().max() <= 2e-5 * g0.abs().max()
        assert ((s2 - s0).abs() <= 1e-5 * s0.abs() + 1e-30).all()
    # default: ONE streamed launch that waits for per-piece arrival flags — the same launch geometry as the device batch,
    # hence the same bits, and repeatable while the staging buffers are being overwritten step after step
    for _ in range(3):
        w3, g3, s3, n3 = run(lambda w, opt: k.train_step_host(w, opt, host[0], host[1], host[2], 0.3, 1e-3))
        assert torch.equal(g0, g3) and torch.equal(w0, w3) and torch.equal(s0, s3) and torch.equal(n0, n3)
    assert not torch.equal(w0, params)                  # the step really updated the weights


def test_streamed_batch_that_never_arrives_is_reported_not_applied(monkeypatch):
    """hjb_vhjb_loss_grad_streamed with arrival flags that are never set: the bounded poll gives up (HJB_STREAM_POLL_LIMIT
    shortens it for the test), the step's loss sums are NaN, the workspace's failure word is raised, the guarded Adam
    update leaves the weights and the optimiser state untouched, and the host side raises."""
    import ctypes as C
    from q_learning_with_hjb_b200 import _lib as L
    monkeypatch.setenv("HJB_STREAM_POLL_LIMIT", "200")
    B = 4 * 148 * 64
    torch, k, p, orc, params, xs, dones, costs = _setup("quad10d", B, seed=3)
    xd, dd, cd = _dev(torch, xs, dones, costs)
    flags = torch.zeros(8, device="cuda", dtype=torch.int32)          # never set
    k.counts(dd, p.eps)
    k._bind(params)
    L.check(L.lib().hjb_vhjb_loss_grad_streamed(k.sys_spec, k.net, k.task, L.ptr(xd), L.ptr(dd), L.ptr(cd), B, L.ptr(k.norm),
                                                0.5, L.ptr(k.grad), L.ptr(k.sums), L.ptr(k.workspace), L.ptr(flags), 148 * 64,
                                                L.stream_ptr()))
    w = params.clone()
    mu, nu = torch.zeros_like(w), torch.zeros_like(w)
    L.check(L.lib().hjb_vhjb_adam_guarded(L.ptr(w), L.ptr(mu), L.ptr(nu), L.ptr(k.grad), k.n, 1e-3, 0.9, 0.999, 1e-8, 1,
                                          L.ptr(k.workspace), L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.isnan(k.sums).all()
    assert torch.equal(w, params) and not mu.any() and not nu.any()
    assert k.stream_failures(reset=False) > 0
    with pytest.raises(RuntimeError, match="timed out"):
        k.check_streams()
    assert k.stream_failures() == 0                                   # check_streams reset the word
    # the same launch with the flags set computes the device-batch gradient, and the guarded update applies it
    flags.fill_(1)
    L.check(L.lib().hjb_vhjb_loss_grad_streamed(k.sys_spec, k.net, k.task, L.ptr(xd), L.ptr(dd), L.ptr(cd), B, L.ptr(k.norm),
                                                0.5, L.ptr(k.grad), L.ptr(k.sums), L.ptr(k.workspace), L.ptr(flags), 148 * 64,
                                                L.stream_ptr()))
    g_stream = k.grad.clone()
    L.check(L.lib().hjb_vhjb_adam_guarded(L.ptr(w), L.ptr(mu), L.ptr(nu), L.ptr(k.grad), k.n, 1e-3, 0.9, 0.999, 1e-8, 1,
                                          L.ptr(k.workspace), L.stream_ptr()))
    k.loss_grad(params, xd, dd, cd, 0.5)
    assert torch.equal(g_stream, k.grad) and torch.isfinite(k.sums).all()
    assert not torch.equal(w, params)


def test_zero_normaliser_gives_zero_weight_not_nan():
    """eps = 0 and a shard without boundary samples: norm[1] = 0.  The masked terms carry weight 0, not 0 * inf."""
    torch, k, p, orc, params, xs, dones, costs = _setup("quad2d", 2048, seed=5)
    dones[:] = 0
    xd, dd, cd = _dev(torch, xs, dones, costs)
    k.counts(dd, 0.0)
    assert float(k.norm[1]) == 0.0
    for impl in ("tensor", "simt"):
        k.impl = impl
        g, sums = k.loss_grad(params, xd, dd, cd, 0.5)
        assert torch.isfinite(g).all() and torch.isfinite(sums).all(), impl


def test_device_replay_buffer_has_deque_semantics():
    """DeviceReplayBuffer (ring in HBM, hjb_replay_append / hjb_replay_gather) against collections.deque(maxlen): plain
    extends, rollout records appended trajectory by trajectory, wrap-around, and an extend larger than the capacity."""
    from collections import deque
    torch = _cuda()
    from q_learning_with_hjb_b200.controller.vhjb import DeviceReplayBuffer
    rng = np.random.default_rng(0)
    n, cap = 3, 37
    buf, ref = DeviceReplayBuffer(n, cap), deque(maxlen=cap)

    def check():
        xs, cs, ds = buf.contents()
        assert len(buf) == len(ref)
        np.testing.assert_array_equal(xs, np.array([r[0] for r in ref], dtype=np.float32).reshape(-1, n))
        np.testing.assert_array_equal(cs, np.array([r[1] for r in ref], dtype=np.float32))
        np.testing.assert_array_equal(ds, np.array([r[2] for r in ref], dtype=np.float32))

    def plain(k):
        xs = rng.normal(size=(k, n)).astype(np.float32); cs = rng.uniform(size=k).astype(np.float32)
        ds = (rng.uniform(size=k) < 0.3).astype(np.float32)
        buf.extend(xs, cs, ds)
        ref.extend((xs[i], cs[i], ds[i]) for i in range(k))

    def rollout(T1, N):
        rx = rng.normal(size=(T1, N, n)).astype(np.float32); rc = rng.uniform(size=(T1, N)).astype(np.float32)
        lens = rng.integers(1, T1 + 1, size=N)
        rd = np.full((T1, N), -1.0, dtype=np.float32)
        for e in range(N):
            rd[:lens[e], e] = 0.0
            rd[lens[e] - 1, e] = 1.0
        got = buf.extend_rollout(torch.as_tensor(rx).cuda(), torch.as_tensor(rc).cuda(), torch.as_tensor(rd).cuda())
        assert got == int(lens.sum())
        for e in range(N):                                # the reference's order: trajectory by trajectory (vhjb.py:305)
            ref.extend((rx[t, e], rc[t, e], rd[t, e]) for t in range(lens[e]))

    plain(10); check()
    buf.append(np.ones(n), 2.0, 1.0); ref.append((np.ones(n, np.float32), np.float32(2.0), np.float32(1.0))); check()
    rollout(6, 4); check()                                # still below the capacity
    rollout(9, 5); check()                                # wraps around
    plain(5); check()
    rollout(30, 4); check()                               # may exceed the capacity in one extend
    plain(100); check()                                   # certainly does
    # one shuffled pass: full minibatches, every row drawn at most once, all rows come from the buffer
    rows = {tuple(np.concatenate([x, [c, d]]).tolist()) for x, c, d in zip(*buf.contents())}
    seen = []
    for xs, cs, ds in buf.batches(8):
        assert xs.shape == (8, n) and cs.shape == (8,) and ds.shape == (8,)
        seen += [tuple(np.concatenate([x, [c, d]]).tolist()) for x, c, d in zip(xs.cpu().numpy(), cs.cpu().numpy(), ds.cpu().numpy())]
    assert len(seen) == (cap // 8) * 8 and len(set(seen)) == len(seen) and set(seen) <= rows


@pytest.mark.parametrize("name", ["quad10d", "cartpole_tanh", "di_mintime"])
def test_tiny_ragged_and_empty_batches(name):
    """Edge cases of the batch dimension on the tensor path: one state, one short of / one over a 64-state tile, fewer
    tiles than SMs, and the empty batch (loss sums and gradient are zeros, nothing is read)."""
    torch, k, p, orc, params, xs, dones, costs = _setup(name, 200, seed=17, wseed=6)
    for B in (1, 63, 65, 129):
        xd, dd, cd = _dev(torch, xs[:B], dones[:B], costs[:B])
        out, sums = k.residual(params, xd, dd, cd)
        q = orc.pieces(xs[:B], running=costs[:B] if p.residual_form == "min_time" else None)
        np.testing.assert_allclose(out["V"].cpu().numpy(), q["V"].detach().numpy(), rtol=2e-5, atol=1e-6)
        k.norm.copy_(torch.tensor([float(B), 1.0]))
        k.loss_grad(params, xd, dd, cd, 0.0)
        assert torch.isfinite(k.grad).all() and torch.isfinite(k.sums).all()
    empty = torch.empty((0, p.sys.n), device="cuda"), torch.empty(0, device="cuda"), torch.empty(0, device="cuda")
    out, sums = k.residual(params, *empty)
    assert out["V"].shape == (0,) and float(sums.abs().sum()) == 0.0
    k.norm.copy_(torch.tensor([1.0, 1.0]))
    grad, sums = k.loss_grad(params, *empty, 0.0)
    assert float(grad.abs().sum()) == 0.0 and float(sums.abs().sum()) == 0.0


@pytest.mark.parametrize("name,B", [("quad10d", 256), ("di_mintime", 256), ("cartpole_tanh", 4096 + 7), ("quad10d", (1 << 18) + 100)])
def test_single_call_train_step_is_bit_identical_to_the_separate_entry_points(name, B):
    """hjb_vhjb_train_step (count + fused loss/gradient + reduce-and-Adam: one call, three launches — the small-batch path of
    VhjbKernels.train_step) against hjb_vhjb_count + hjb_vhjb_loss_grad + hjb_adam, three consecutive updates."""
    from q_learning_with_hjb_b200 import parallel
    from q_learning_with_hjb_b200.controller.vhjb import AdamState
    torch, k, p, orc, params, xs, dones, costs = _setup(name, B, seed=23, wseed=7)
    xd, dd, cd = _dev(torch, xs, dones, costs)

    def separate(w, opt, reg):
        k.counts(dd, 0.0)
        if p.residual_form == "min_time":
            k.norm[1] = 1.0
        else:
            k.norm.add_(p.eps)
        k.loss_grad(w, xd, dd, cd, reg)
        opt.count += 1
        k.adam(w, opt.mu, opt.nu, k.grad, opt.count, 1e-3)

    results = []
    for fused in (False, True):
        w = params.clone()
        opt = AdamState(0, torch.zeros_like(w), torch.zeros_like(w))
        for i in range(3):
            if fused:
                assert not parallel.is_distributed()
                k.train_step(w, opt, xd, dd, cd, 0.1 * i, 1e-3)
            else:
                separate(w, opt, 0.1 * i)
        results.append((w.clone(), opt.mu.clone(), opt.nu.clone(), k.grad.clone(), k.sums.clone(), k.norm.clone()))
    for a, b in zip(*results):
        assert torch.equal(a, b)
    assert not torch.equal(results[0][0], params)


def test_adjoint_chain_overflow_sends_the_launch_through_the_fp32_pass():
    """A net with a large backward gain (weights of a trained net scaled up) and states next to the goal, whose seeds
    enter the fp16 chain at its cap: the chain passes fp16's 65504 behind in-range seeds.  The tensor kernel's conversions
    saturate and it counts the event; the fp32 pass behind it sees the count and runs the WHOLE launch, the reductions
    leave the tensor launch's partials out: the gradient is exact, nothing is reported as saturated, every state is
    reported as deferred — decided on the device, launch by launch, the kernels object stays on the tensor path.
    (Found in the wild: the reference's linear_vhjb_controller.gin run turned NaN at update 13,606 in round 1; round 2's
    first version clipped and counted such a launch and left the switch to fp32 to the host's per-epoch poll.)"""
    B = 1024
    torch, k, p, orc, params, xs, dones, costs = _setup("linear", B, seed=31, wseed=9)
    scale = torch.ones_like(params)
    n1 = 2 * 128
    scale[:n1] = 1e-4                                    # W1 x 1e-4, W2 x 100, W3 x 100: forward values stay moderate
    scale[n1:] = 100.0                                   # (|y| ~ |z|), the backward chain gains 1e-4 x 100^4 = 1e4
    big = (params * scale).contiguous()
    rng = np.random.default_rng(0)
    xs[:8] = (p.xf + 2e-3 * rng.uniform(-1, 1, size=(8, p.sys.n))).astype(np.float32)
    dones[:] = 0
    xd, dd, cd = _dev(torch, xs, dones, costs)
    k.counts(dd, p.eps)
    k.saturated_total(reset=True)
    g_tc, sums_tc = k.loss_grad(big, xd, dd, cd, 0.0)
    g_tc, sums_tc = g_tc.clone(), sums_tc.clone()
    assert torch.isfinite(g_tc).all()
    assert k.saturated() == 0 and k.deferred() == B and k.impl == "tensor"
    # a launch without such states, right after: back on the tensor cores
    k.loss_grad(params, xd, dd, cd, 0.0)
    assert k.saturated() == 0 and k.deferred() < B // 4 and k.saturated_total() == 0
    k.impl = "simt"
    try:
        g_cc, sums_cc = k.loss_grad(big, xd, dd, cd, 0.0)
        g_cc, sums_cc = g_cc.clone(), sums_cc.clone()
        assert k.saturated() == 0
    finally:
        k.impl = "tensor"
    # the same fp32 arithmetic in another tile order
    assert (g_tc - g_cc).abs().max() <= 1e-5 * g_cc.abs().max()
    assert torch.allclose(sums_tc[:2], sums_cc[:2], rtol=1e-5)
    W64 = [w.astype(np.float64) for w in np.split(big.cpu().numpy(), [n1, n1 + 128 * 128])]
    orc2 = V.VhjbOracle(p, [W64[0].reshape(2, 128), W64[1].reshape(128, 128), W64[2].reshape(128, 64)])
    _, _, _, grads, _ = orc2.loss_and_grad(xs, dones, costs, 0.0)
    go = np.concatenate([x.reshape(-1) for x in grads])
    assert np.abs(g_cc.cpu().numpy() - go).max() <= TOL * np.abs(go).max()


def test_wild_batch_that_overflowed_the_first_range_management():
    """tests/golden/vhjb_linear_overflow_batch.npz: weights and minibatch of update 13,606 of the reference's
    linear_vhjb_controller.gin run, where a state 1.6e-3 from the goal (seeds 2^8 x typical) met a backward gain of 267:
    68,398 in fp16.  A third of this minibatch lies within 0.02 of the goal, where the trained net's dV/dx is what is left
    after a ~100-fold cancellation in g1 W1^T; tcgen05.mma accumulates with truncation, so on the tensor chain the seeds
    of such a state are good to ~2.5e-4 only (round 1 asserted 3e-4 here).  Those states now take the fp32 pass (their seeds
    are > 2^6 x the batch-typical weight): the reference's own minibatch meets the north star's 1e-4, nothing saturates."""
    import os
    torch = _cuda()
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "vhjb_linear_overflow_batch.npz"))
    k, p = make_kernels("linear")
    params = torch.as_tensor(d["params"]).cuda()
    xs, dones, costs, reg = d["xs"], d["dones"], d["costs"], float(d["reg"])
    xd, dd, cd = _dev(torch, xs, dones, costs)
    k.counts(dd, p.eps)
    grad = k.loss_grad(params, xd, dd, cd, reg)[0]
    assert torch.isfinite(grad).all() and k.saturated() == 0
    W = np.split(d["params"].astype(np.float64), [256, 256 + 128 * 128])
    orc = V.VhjbOracle(p, [W[0].reshape(2, 128), W[1].reshape(128, 128), W[2].reshape(128, 64)])
    _, _, _, grads, _ = orc.loss_and_grad(xs, dones, costs, reg)
    g = grad.cpu().numpy().astype(np.float64)
    off = 0
    for gi in grads:
        sl = slice(off, off + gi.size); off += gi.size
        assert np.abs(g[sl] - gi.reshape(-1)).max() <= TOL * np.abs(gi).max()
    assert k.deferred() > 0
    k.impl = "simt"
    try:
        g32 = k.loss_grad(params, xd, dd, cd, reg)[0].cpu().numpy().astype(np.float64)
    finally:
        k.impl = "tensor"
    go = np.concatenate([gi.reshape(-1) for gi in grads])
    assert np.abs(g32 - go).max() <= 1e-5 * np.abs(go).max()          # the fp32 kernels: 2.7e-7


def test_train_loop_loss_accumulator_matches_params_update():
    """VHJBController.train()'s fast path (fused step + device accumulator of the step losses) returns the same epoch
    averages as the loop over params_update (taken when the method is overridden), update for update."""
    import torch
    results = []
    for override in (False, True):
        np.random.seed(0)
        dyn, ctl = _cartpole_controller(12)
        ctl.epochs, ctl.batch_size = 3, 16                                # (the seed data set holds 20 samples)
        if override:
            orig = ctl.params_update
            ctl.params_update = lambda *a, **kw: orig(*a, **kw)          # an instance override: the generic loop
        lists = ctl.train()
        results.append((lists, ctl.model_params.flat.clone(), ctl.update_counter))
    (la, wa, ca), (lb, wb, cb) = results
    assert ca == cb and ca > 0 and torch.equal(wa, wb)
    for a, b in zip(la, lb):
        np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize("name", ["quad10d", "di_mintime", "cartpole"])
def test_bench_workload_kernels_are_the_tested_kernels(name):
    """bench.py builds the kernels it times from q_learning_with_hjb_b200/workloads.py; the parity tests above build them
    from the oracle's problem descriptions (tests/helpers_vhjb.py).  Same task structure, the same bits out."""
    import torch
    from q_learning_with_hjb_b200 import workloads as WL
    k1, w = WL.make_vhjb_kernels(name)
    k2, p = make_kernels(name)
    params = torch.as_tensor(WL.flat_params(WL.init_weights(len(w.xf), seed=2))).cuda()
    xd, dd, cd = (torch.as_tensor(a).cuda() for a in WL.sample_vhjb_batch(name, 4096, seed=9))
    outs = []
    for k in (k1, k2):
        k.counts(dd, w.eps)
        g, s = k.loss_grad(params, xd, dd, cd, 0.5)
        outs.append((g.clone(), s.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_peer_exchange_step_with_one_rank_equals_the_plain_step():
    """hjb_vhjb_train_step_peer (reduce -> exchange over peer memory -> Adam in one kernel; the N-GPU path, verified against
    the single-GPU run at N = 2, 4, 8 by bench.py's `verify` block) driven here with world = 1: the rank stores into its own
    exchange buffer, raises and awaits its own flag, sums one slot and updates — three steps must leave the same bits in
    the weights, the Adam moments and the loss sums as hjb_vhjb_train_step, and return the next batch's normalisers."""
    import torch
    from q_learning_with_hjb_b200 import _lib as L
    from q_learning_with_hjb_b200.controller.vhjb import AdamState
    B = 20000
    name = "quad10d"
    ka, p = make_kernels(name)
    kb, _ = make_kernels(name)
    xs, dones, costs = (torch.as_tensor(a).cuda() for a in sample_batch(name, B, seed=3))
    W0 = torch.as_tensor(flat_params(V.init_weights(p.sys.n, seed=1))).cuda()
    # plain single-process steps
    wa = W0.clone()
    oa = AdamState(0, torch.zeros_like(wa), torch.zeros_like(wa))
    sums_a = []
    for _ in range(3):
        s, _n = ka.train_step(wa, oa, xs, dones, costs, 0.25, 1e-3, local=True)
        sums_a.append(s.clone())
    # the peer entry point, world = 1
    lib = L.lib()
    nf, ng = int(lib.hjb_vhjb_peer_exchange_floats(p.sys.n, 1)), int(lib.hjb_vhjb_peer_exchange_flags(p.sys.n, 1))
    buf = torch.zeros(nf, device="cuda", dtype=torch.float32)
    flg = torch.zeros(ng, device="cuda", dtype=torch.int32)
    bufs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device="cuda")
    flags = torch.tensor([flg.data_ptr()], dtype=torch.int64, device="cuda")
    wb = W0.clone()
    ob = AdamState(0, torch.zeros_like(wb), torch.zeros_like(wb))
    kb.counts(dones, p.eps)                                   # this batch's normalisers (one rank: local = global)
    norm = [kb.norm.clone(), torch.zeros(2, device="cuda")]
    local_next = torch.zeros(2, device="cuda")
    loss_acc = torch.zeros(3, device="cuda")
    for i in range(3):
        L.check(lib.hjb_vhjb_count(L.ptr(dones), B, 0.0, L.ptr(local_next), L.ptr(kb.workspace), L.stream_ptr()), "count")
        kb._bind(wb)
        ob.count += 1
        L.check(lib.hjb_vhjb_train_step_peer(kb.sys_spec, kb.net, kb.task, L.ptr(xs), L.ptr(dones), L.ptr(costs), B, 0.25, 1e-3,
                                             0.9, 0.999, 1e-8, int(ob.count), L.ptr(ob.mu), L.ptr(ob.nu), L.ptr(norm[i & 1]),
                                             L.ptr(kb.grad), L.ptr(kb.sums), L.ptr(loss_acc), L.ptr(local_next),
                                             L.ptr(norm[(i + 1) & 1]), L.ptr(bufs), L.ptr(flags), 0, 1, L.ptr(kb.workspace),
                                             L.stream_ptr()), "hjb_vhjb_train_step_peer")
        assert torch.equal(kb.sums[:2], sums_a[i][:2]), i
        assert torch.equal(norm[(i + 1) & 1], norm[i & 1])     # same batch again: the delivered normalisers are this step's
    assert torch.equal(wa, wb) and torch.equal(oa.mu, ob.mu) and torch.equal(oa.nu, ob.nu)
    assert kb.stream_failures() == 0 and not torch.equal(wb, W0)

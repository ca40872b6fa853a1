"""Helpers for the vhjb tests: standard problems (SURVEY.md §8d C2 / C5 and the reference's own gin configs)."""
import numpy as np

from oracle import rollout_oracle as O
from oracle import vhjb_oracle as V


def problem(name: str) -> V.VhjbProblem:
    if name == "linear":        # linear.gin + linear_vhjb_controller.gin
        s = O.std_system("linear")
        return V.VhjbProblem(s, np.eye(2), np.eye(1), np.zeros(2), np.zeros(1), np.zeros(2), np.ones(2))
    if name == "cartpole":      # cartpole.gin + cartpole_vhjb_controller.gin
        s = O.std_system("cartpole")
        return V.VhjbProblem(s, np.eye(4), np.eye(1), np.array([0, 3.1415926, 0, 0]), np.zeros(1), np.zeros(4), np.ones(4))
    for suffix in ("_tanh", "_sin"):   # same task, smooth value net (cartpole_balancing.ipynb cell 6 uses tanh)
        if name.endswith(suffix):
            p = problem(name[:-len(suffix)]); p.act = suffix[1:]; return p
    if name == "quad2d":        # quadrotors2D.gin + quadrotors2DHovering_vhjb_controller.gin
        s = O.std_system("quad2d")
        return V.VhjbProblem(s, np.eye(6), np.eye(2), np.zeros(6), np.array([4.905, 4.905]), np.zeros(6), np.ones(6))
    if name == "quad10d":       # C5: 10D_quadcopte.ipynb cells 4, 6, 9
        s = O.std_system("quad10d")
        uf = np.array([s.par["g"] * s.par["m"] / s.par["kT"], 0, 0])
        return V.VhjbProblem(s, np.eye(10), np.eye(3), np.zeros(10), uf, np.zeros(10), np.ones(10))
    if name == "di_mintime":    # C2: double_integrator_optimal_time.ipynb cells 5, 7, 11 (sin net, bang-bang, |vdot + l|)
        s = O.OracleSystem("linear", 2, 1, 0.01, np.array([-1.0]), np.array([1.0]),
                           {"A": np.array([[0.0, 1.0], [0.0, 0.0]]), "B": np.array([[0.0], [1.0]])})
        return V.VhjbProblem(s, np.eye(2), np.eye(1), np.zeros(2), np.zeros(1), np.zeros(2), np.ones(2),
                             act="sin", control_form="bang", residual_form="min_time")
    raise ValueError(name)


OBS = {"linear": [2, 3], "cartpole": [4.8, 0.418, 4, 4],
       "quad2d": [2, 2, 1.5, 5, 5, 2], "quad10d": [2, 2, 2, .5, .5, 4, 4, 4, 2, 2], "di_mintime": [1, 1]}


def base_name(name: str) -> str:
    for suffix in ("_tanh", "_sin"):
        if name.endswith(suffix):
            return name[:-len(suffix)]
    return name


def sample_batch(name: str, B: int, seed: int = 0):
    """States U(obs_min, obs_max) about xf, dones ~ Bernoulli(0.1), costs ~ U(0.1, 10)."""
    p = problem(name)
    rng = np.random.default_rng(seed)
    xs = rng.uniform(-1, 1, size=(B, p.sys.n)) * np.asarray(OBS[base_name(name)]) + p.xf
    if p.residual_form == "min_time":
        dones = np.zeros(B)
        costs = ((xs ** 2).sum(1) > 1e-4).astype(np.float64)     # running cost l_i (nb cell 7:4)
    else:
        dones = (rng.uniform(size=B) < 0.1).astype(np.float64)
        costs = rng.uniform(0.1, 10, size=B)
    return xs.astype(np.float32), dones.astype(np.float32), costs.astype(np.float32)


def exact_quadratic_weights(n, P, eps_s=1e-3):
    """Weights for which the relu net represents V(z) = z^T P z EXACTLY: layer 1 splits z into (z+, z-), layer 2
    passes them through, layer 3 applies the Cholesky factor of P - eps_s I to z+ - z-."""
    W1 = np.zeros((n, 128)); W2 = np.zeros((128, 128)); W3 = np.zeros((128, 64))
    W1[np.arange(n), np.arange(n)] = 1.0
    W1[np.arange(n), n + np.arange(n)] = -1.0
    W2[np.arange(2 * n), np.arange(2 * n)] = 1.0
    Lc = np.linalg.cholesky(P - eps_s * np.eye(n)).T      # Lc^T Lc = P - eps_s I
    W3[:n, :n] = Lc.T
    W3[n:2 * n, :n] = -Lc.T
    return [W1, W2, W3]


# ---- product-side construction (q_learning_with_hjb_b200) for the same problems ----
def make_kernels(name: str):
    """VhjbKernels + the matching oracle problem for a named problem."""
    from q_learning_with_hjb_b200.controller.vhjb import VhjbKernels
    from tests.helpers import make_dynamics
    p = problem(name)
    kind = {"linear": "linear", "cartpole": "cartpole", "quad2d": "quad2d", "quad10d": "quad10d",
            "di_mintime": "linear"}[base_name(name)]
    dyn = make_dynamics(kind)
    if name == "di_mintime":
        dyn.dt = 0.01
        dyn.umin, dyn.umax = np.float32([-1]), np.float32([1])
    k = VhjbKernels(dyn, p.xf, p.uf, p.Q, p.R, p.mean, p.std, p.eps, p.eps_s, act=p.act,
                    control_form={"clip": "clipped", "bang": "bangbang"}[p.control_form],
                    residual_form=p.residual_form)
    return k, p


def flat_params(weights):
    return np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in weights])


def smoke():
    """One small fused vhjb loss+gradient step on cuda:0 against the torch-fp64 oracle (used by __graft_entry__)."""
    import torch
    k, p = make_kernels("quad10d")
    W = V.init_weights(p.sys.n, seed=0)
    W32 = [w.astype(np.float32) for w in W]
    xs, dones, costs = sample_batch("quad10d", 2048, seed=1)
    params = torch.as_tensor(flat_params(W32)).cuda()
    xd, dd, cd = (torch.as_tensor(a).cuda() for a in (xs, dones, costs))
    k.counts(dd, p.eps)
    grad, sums = k.loss_grad(params, xd, dd, cd, 0.5)
    orc = V.VhjbOracle(p, [w.astype(np.float64) for w in W32])
    total, hjb, term, grads, _ = orc.loss_and_grad(xs, dones, costs, 0.5)
    g = grad.cpu().numpy()
    go = np.concatenate([x.reshape(-1) for x in grads])
    gerr = np.abs(g - go).max() / np.abs(go).max()
    herr = abs(float(sums[0] / k.norm[0]) - hjb) / hjb
    assert gerr < 1e-4 and herr < 1e-4, (gerr, herr)
    print(f"smoke: vhjb quad10d loss+grad 2048 states: grad err {gerr:.2e}, hjb loss err {herr:.2e}")

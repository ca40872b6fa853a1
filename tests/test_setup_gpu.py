"""SURVEY.md 8f row 2 on the device: the one-off set-up work of the reference's constructors — linearisation about (xf, uf)
and the Riccati solve (controller/vhjb.py:156-160, utils/utils.py:30-80, every model-based constructor) and the sampling of
the seed data set (controller/vhjb.py:136-151) — for MANY problems / samples at once, with the arrays on the GPU."""
import numpy as np
import pytest

from tests.helpers import make_controller, make_dynamics

pytestmark = pytest.mark.gpu


def test_batched_riccati_on_the_device_matches_scipy():
    import scipy.linalg
    import torch
    from q_learning_with_hjb_b200.utils import utils as U
    assert torch.cuda.is_available()
    dyn = make_dynamics("quad2d")
    hover = make_controller("quad2d_hover", dyn)
    A = torch.as_tensor(hover.A, device="cuda")
    B = torch.as_tensor(hover.B, device="cuda")
    rng = np.random.default_rng(0)
    qs = np.exp(rng.uniform(-2, 2, size=(2000, 6)))                       # 2000 cost weightings at once
    rs = np.exp(rng.uniform(-1, 1, size=(2000, 2)))
    Q = torch.diag_embed(torch.as_tensor(qs, device="cuda"))
    R = torch.diag_embed(torch.as_tensor(rs, device="cuda"))
    K, P = U.lqr_gains_batched(A, B, Q, R)
    assert P.is_cuda and P.shape == (2000, 6, 6) and K.shape == (2000, 2, 6)
    for i in (0, 7, 1999):
        Pi = scipy.linalg.solve_continuous_are(hover.A, hover.B, np.diag(qs[i]), np.diag(rs[i]))
        np.testing.assert_allclose(P[i].cpu().numpy(), Pi, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(K[i].cpu().numpy(), np.linalg.solve(np.diag(rs[i]), hover.B.T @ Pi), rtol=1e-8, atol=1e-10)
    # residual of the Riccati equation for all of them
    Bt = B.transpose(-1, -2)
    res = A.T @ P + P @ A - P @ B @ torch.linalg.solve(R, Bt @ P) + Q
    assert float(res.abs().max()) < 1e-8 * float(P.abs().max())


def test_batched_linearisation_through_the_device_dynamics():
    """linearize_batched drives Dynamics.dynamics_step (one hjb_dynamics launch for all perturbed points) — the hover
    linearisations the reference hard-codes (quadrotors_model_based_controller.py:25-31, :58-68) come out of the device
    dynamics for many hover points at once."""
    from q_learning_with_hjb_b200.utils import utils as U
    for kind, ckind in (("quad2d", "quad2d_hover"), ("quad10d", "quad10d_hover")):
        dyn = make_dynamics(kind)
        dyn.fast_trig = False
        hover = make_controller(ckind, dyn)
        n, m = dyn.get_dimension()
        rng = np.random.default_rng(1)
        xf = np.zeros((64, n))
        xf[:, : (2 if kind == "quad2d" else 3)] = rng.uniform(-2, 2, size=(64, 2 if kind == "quad2d" else 3))   # hover anywhere
        uf = np.tile(hover.uf, (64, 1))
        A, B = U.linearize_batched(dyn.dynamics_step, xf, uf, eps=1e-2)
        assert A.shape == (64, n, n) and B.shape == (64, n, m)
        np.testing.assert_allclose(A, np.tile(hover.A, (64, 1, 1)), atol=2e-3)
        np.testing.assert_allclose(B, np.tile(hover.B, (64, 1, 1)), atol=2e-3)


def test_seed_data_set_sampled_on_the_device():
    """The interior / boundary samples of the seed data set (controller/vhjb.py:136-151: wrap(U(-1, 1) std + mean)) drawn
    on the device by the counter-based generator instead of one np.random call per sample."""
    import torch
    dyn = make_dynamics("cartpole")
    mean, std = np.float32([0, 3.1415926, 0, 0]), np.float32([4.8, 0.418, 4, 4])
    x = dyn.sample_initial_states(1 << 20, seed=9, mean=mean, std=std)
    assert x.is_cuda and x.shape == (1 << 20, 4)
    z = x.clone()
    z[:, 1] = torch.remainder(z[:, 1] - 3.1415926 + np.pi, 2 * np.pi) - np.pi     # error angle
    assert float(z.abs().max(dim=0).values[0]) <= 4.8 and float(z[:, 1].abs().max()) <= 0.418 + 1e-5
    assert abs(float(z[:, 0].std()) - 4.8 / np.sqrt(3)) < 0.02 and abs(float(z[:, 2].mean())) < 0.02
    assert float(x[:, 1].min()) >= -np.pi - 1e-6 and float(x[:, 1].max()) <= np.pi + 1e-6   # states_wrap applied

"""``Controller`` — the reference's interface (controller/controller_basic.py:1-5): ``get_control_efforts(x)``.

Model-based controllers additionally describe themselves to the CUDA library through ``control_spec()``
(an ``hjb_control`` struct: gains, goal, clip flag), which is what lets ``Dynamics.rollout`` fuse the control
law into the rollout kernel; ``DeviceController.get_control_efforts`` evaluates the same device function
per state (or per batch of states) through ``hjb_control_efforts``.
"""
from __future__ import annotations

import numpy as np

from q_learning_with_hjb_b200 import _lib as L


class Controller:
    def __init__(self) -> None:
        pass

    def get_control_efforts(self, x):
        raise NotImplementedError


class DeviceController(Controller):
    """Base of the controllers whose law is compiled into libhjb_b200."""

    #: the Dynamics object the controller acts on (set by subclasses)
    dynamics = None

    def control_spec(self) -> "L.HjbControl":
        raise NotImplementedError

    def get_control_efforts(self, x):
        """x (n,) -> u (m,), or batched (B, n) -> (B, m).  NumPy in -> NumPy (float64) out; CUDA tensor in ->
        CUDA tensor out."""
        return self._efforts(self.dynamics.system_spec(), self.control_spec(), x)

    def _efforts(self, sys_spec, ctl_spec, x):
        """``hjb_control_efforts`` with explicit parameter structs (subclasses evaluate single branches of their law
        by editing a copy of the spec)."""
        torch = L.require_cuda()
        dyn = self.dynamics
        n, m = dyn.state_dim, dyn.control_dim
        single = np.ndim(x) == 1
        xd = L.dev_f32(x, (-1, n))
        u = torch.empty((xd.shape[0], m), device="cuda", dtype=torch.float32)
        L.check(L.lib().hjb_control_efforts(sys_spec, ctl_spec, int(dyn.fast_trig), L.ptr(xd),
                                            xd.shape[0], L.ptr(u), L.stream_ptr()), "hjb_control_efforts")
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return u[0] if single else u
        out = u.cpu().numpy().astype(np.float64)
        return out[0] if single else out


def unclipped(sys_spec):
    """The system's parameter struct with the input limits opened (for the un-clipped branch evaluations)."""
    for k in range(len(sys_spec.umin)):
        sys_spec.umin[k], sys_spec.umax[k] = -3.0e38, 3.0e38
    return sys_spec


def closed_loop(dynamics, controller, x0, t):
    """xs [len(t), n], us [len(t) - 1, m] of the reference's per-step loop ``u = get_control_efforts(x); x = simulate(x,
    u)`` over the time grid ``t`` — ONE rollout launch."""
    res = dynamics.rollout(controller, np.asarray(x0, dtype=np.float32).reshape(1, -1), len(t) - 1, record_stride=1)
    return res.xs_env[0].astype(np.float64), res.us_env[0].astype(np.float64)


def lqr_gain(A, B, Q, R):
    """P from the continuous-time algebraic Riccati equation, K = R^-1 B^T P — host-side, SciPy, once per
    controller (the reference does the same in every model-based constructor, e.g. controller/lqr.py:25-26)."""
    import scipy.linalg

    A, B, Q, R = (np.asarray(v, dtype=np.float64) for v in (A, B, Q, R))
    P = scipy.linalg.solve_continuous_are(A, B, Q, R)
    K = scipy.linalg.solve(R, B.T @ P)
    return K, P

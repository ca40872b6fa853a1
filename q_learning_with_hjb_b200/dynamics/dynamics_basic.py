"""``Dynamics`` — the reference's base-class interface (dynamics/dynamics_basic.py:7-122) on top of the CUDA
library.

Interface kept: ``get_initial_state / get_dimension / get_control_limit / get_M / get_C / get_G / get_B /
states_wrap / get_control_affine_matrix / dynamics_step / simulate`` with the reference's shapes.  The
arithmetic of ``get_control_affine_matrix``, ``dynamics_step`` and ``simulate`` runs in the sm_100a kernels
(``hjb_dynamics``); every one of them also accepts a leading batch axis.  New: ``rollout`` — the whole
closed loop for a batch of environments in one launch (``hjb_rollout``).

Host-side setup stays on the host exactly as in the reference: RNG sampling of initial states
(``np.random``, seeded in ``__init__`` — dynamics_basic.py:26), and the manipulator matrices ``get_M/C/G/B``
that the model-based controllers use once, at construction, to linearise about the goal.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from q_learning_with_hjb_b200 import _lib as L


class Dynamics:
    #: hjb_system_kind of the subclass
    KIND: int = -1
    #: indices of the angle components that ``states_wrap`` maps into [-pi, pi)
    WRAP_INDEX: Tuple[int, ...] = ()

    def __init__(self, config) -> None:
        self.state_dim = int(config.state_dim)
        self.control_dim = int(config.control_dim)
        self.dt = config.dt
        self.umin = config.umin
        self.umax = config.umax
        self.x0_mean = config.x0_mean
        self.x0_std = config.x0_std
        self.seed = config.seed
        np.random.seed(config.seed)
        #: True (default): the kernels' in-line trigonometry (quadrant reduction + minimax polynomials, <= 7e-8 absolute;
        #: MUFU.RCP for reciprocals) — the instantiation bench.py times; False: libdevice sincosf / tanf, IEEE division.
        #: Both meet the 1e-5 per-step / trajectory bounds (tests/test_rollout_gpu.py runs every parity test in both).
        self.fast_trig = True

    # -- reference interface: host-side pieces -------------------------------------------------------
    def get_initial_state(self) -> np.ndarray:
        """x0 = wrap(U(-x0_std, x0_std) + x0_mean), float64, from the global NumPy RNG
        (dynamics_basic.py:28-29)."""
        return self.states_wrap(np.random.uniform(size=(self.state_dim,), low=-self.x0_std, high=self.x0_std)
                                + self.x0_mean)

    def get_initial_states(self, count: int) -> np.ndarray:
        """``count`` consecutive ``get_initial_state()`` draws, stacked [count, n]: ONE vectorised draw from the global NumPy
        RNG — the legacy generator fills a (count, n) request in C order with broadcast bounds, so the numbers are those of
        ``count`` successive calls (tests/test_host_logic.py)."""
        x = np.random.uniform(size=(int(count), self.state_dim), low=-self.x0_std, high=self.x0_std) + self.x0_mean
        return self.states_wrap(x) if count else x

    def sample_initial_states(self, count: int, seed: int = 1234, first: int = 0, mean=None, std=None, out=None):
        """``count`` initial states generated ON THE DEVICE by the counter-based generator (``hjb_sample_states``; fp32 CUDA
        tensor [count, n]): the distribution of ``get_initial_state`` — wrap(U(-x0_std, x0_std) + x0_mean), or the given
        ``mean`` / ``std`` — from a Philox stream keyed by ``seed`` and counted by ``first + i``.  What 16M-environment
        rollouts start from instead of a host array crossing PCIe; ``first`` = a shard's offset into a global
        batch."""
        torch = L.require_cuda()
        x = torch.empty((int(count), self.state_dim), device="cuda", dtype=torch.float32) if out is None else out
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (int(count), self.state_dim)
        L.check(L.lib().hjb_sample_states(int(self.KIND), self.state_dim,
                                          L.c_floats(self.x0_mean if mean is None else mean, self.state_dim),
                                          L.c_floats(self.x0_std if std is None else std, self.state_dim),
                                          int(seed), int(first), int(count), L.ptr(x), L.stream_ptr()), "hjb_sample_states")
        return x

    def get_dimension(self) -> Tuple[int, int]:
        return self.state_dim, self.control_dim

    def get_control_limit(self) -> Tuple[np.ndarray, np.ndarray]:
        return self.umin, self.umax

    def get_M(self, x):
        raise NotImplementedError

    def get_C(self, x):
        raise NotImplementedError

    def get_G(self, x):
        raise NotImplementedError

    def get_B(self):
        raise NotImplementedError

    def states_wrap(self, x):
        """Wrap the angle components into [-pi, pi).  NumPy inputs ((n,) or (B, n)) are wrapped IN PLACE and
        returned, like the reference (cartpole.py:60-64, quadrotors.py:66-70); CUDA tensors are wrapped in
        place by ``hjb_states_wrap``."""
        if _is_cuda_tensor(x):
            t = x if x.dim() == 2 else x.reshape(1, -1)
            if t.dtype.is_floating_point and str(t.dtype) == "torch.float32" and t.is_contiguous():
                L.check(L.lib().hjb_states_wrap(self.system_spec(), L.ptr(t), t.shape[0], L.stream_ptr()),
                        "hjb_states_wrap")
                return x
            raise TypeError("states_wrap on device needs a contiguous float32 tensor")
        assert x.shape[-1] == self.state_dim and x.ndim in (1, 2)
        for i in self.WRAP_INDEX:
            x[..., i] = np.remainder(x[..., i] + np.pi, 2 * np.pi) - np.pi
        return x

    # -- packing for the C ABI -----------------------------------------------------------------------
    def system_params(self):
        """(par[<=8], A, B) in the order include/hjb_b200.h documents for this kind."""
        raise NotImplementedError

    def system_spec(self) -> "L.HjbSystem":
        s = L.HjbSystem()
        s.kind, s.n, s.m, s.dt = self.KIND, self.state_dim, self.control_dim, float(self.dt)
        L.fill(s.umin, np.broadcast_to(np.asarray(self.umin, dtype=np.float32), (self.control_dim,)))
        L.fill(s.umax, np.broadcast_to(np.asarray(self.umax, dtype=np.float32), (self.control_dim,)))
        par, A, B = self.system_params()
        L.fill(s.par, par)
        L.fill(s.A, A)
        L.fill(s.B, B)
        return s

    # -- reference interface: device-side pieces -----------------------------------------------------
    def _dyn(self, x, u, want, integrator="euler"):
        torch = L.require_cuda()
        single = np.ndim(x) == 1
        xd = L.dev_f32(x, (-1, self.state_dim))
        B = xd.shape[0]
        ud = None
        if u is not None:
            if _is_tensor(u):
                ud = L.dev_f32(u, (-1, self.control_dim)).expand(B, self.control_dim).contiguous()
            else:  # scalars are accepted like the reference does (cartpole.py:114-115 passes 0)
                a = np.asarray(u, dtype=np.float32)
                a = np.full((1, self.control_dim), a) if a.ndim == 0 else a.reshape(-1, self.control_dim)
                ud = L.dev_f32(np.broadcast_to(a, (B, self.control_dim)))
        n, m = self.state_dim, self.control_dim
        out = {k: None for k in ("f", "g", "xdot", "x_next")}
        shapes = {"f": (B, n), "g": (B, n, m), "xdot": (B, n), "x_next": (B, n)}
        for k in want:
            out[k] = torch.empty(shapes[k], device="cuda", dtype=torch.float32)
        L.check(L.lib().hjb_dynamics(self.system_spec(), L.INTEGRATORS[integrator], int(self.fast_trig),
                                     L.ptr(xd), L.ptr(ud), B, L.ptr(out["f"]), L.ptr(out["g"]),
                                     L.ptr(out["xdot"]), L.ptr(out["x_next"]), L.stream_ptr()), "hjb_dynamics")
        res = []
        for k in want:
            t = out[k]
            if _is_cuda_tensor(x):
                res.append(t[0] if single else t)
            else:
                a = t.cpu().numpy().astype(np.float64)
                res.append(a[0] if single else a)
        return res

    def get_control_affine_matrix(self, x):
        """x' = f(x) + g(x) u:  x (n,) -> f (n,), g (n, m); or batched (B, n) -> (B, n), (B, n, m)."""
        f, g = self._dyn(x, None, ("f", "g"))
        return f, g

    def dynamics_step(self, x, u):
        """xdot = f(x) + g(x) u, no clipping (dynamics_basic.py:96-105)."""
        return self._dyn(x, u, ("xdot",))[0]

    def simulate(self, x, u, integrator: str = "euler"):
        """One step: clip u to [umin, umax], integrate dt (forward Euler like the reference, or RK4), wrap
        (dynamics_basic.py:107-122)."""
        return self._dyn(x, u, ("x_next",), integrator)[0]

    # -- new: the whole closed loop on device --------------------------------------------------------
    def rollout(self, controller, x0, steps: int, **kwargs):
        """Batched closed-loop rollout; see :func:`q_learning_with_hjb_b200.rollout.rollout`."""
        from q_learning_with_hjb_b200.rollout import rollout

        return rollout(self, controller, x0, steps, **kwargs)


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _is_cuda_tensor(x) -> bool:
    return _is_tensor(x) and x.is_cuda

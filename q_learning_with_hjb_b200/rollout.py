"""Batched closed-loop rollout — the reference's per-environment Python loop

    for t in range(T): u = controller.get_control_efforts(x); x = dynamics.simulate(x, u)

(scripts/test_vhjb_policy.py:146-151 and the notebooks' ``test_learned_policy``) for N environments in one
``hjb_rollout`` launch.

Layout: trajectories are time-major on the device, ``xs[t, env, i]`` (coalesced stores from one thread per
environment); ``RolloutResult.xs_env`` is the reference's per-environment view ``[env, t, i]`` as a
zero-copy permute.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from q_learning_with_hjb_b200 import _lib as L


@dataclass
class RunningCost:
    """l(x, u) = dx^T Q dx + (u - uf)^T R (u - uf), dx = wrap(x - xf) (controller/vhjb.py:162-165); the rollout
    accumulates sum_t l(x_t, u_t) dt with u_t the controller output before ``simulate``'s clip."""
    Q: np.ndarray
    R: np.ndarray
    xf: np.ndarray
    uf: np.ndarray

    def spec(self, n: int, m: int) -> "L.HjbCost":
        c = L.HjbCost()
        L.fill(c.Q, np.asarray(self.Q, dtype=np.float64).reshape(n, n))
        L.fill(c.R, np.asarray(self.R, dtype=np.float64).reshape(m, m))
        L.fill(c.xf, np.asarray(self.xf).reshape(n))
        L.fill(c.uf, np.asarray(self.uf).reshape(m))
        return c


@dataclass
class Box:
    """Freeze an environment once wrap(x - xf) leaves [lo, hi] (controller/vhjb.py:176-181)."""
    xf: np.ndarray
    lo: np.ndarray
    hi: np.ndarray


@dataclass
class RolloutResult:
    xs: Optional[object]        # [T/s + 1, N, n] time-major, or None
    us: Optional[object]        # [T/s, N, m], or None
    x_final: object             # [N, n]
    cost: Optional[object]      # [N], or None
    steps: Optional[object]     # [N] int32, or None

    @property
    def xs_env(self):
        """Per-environment view [N, T/s + 1, n] (the reference's ``xs`` for one environment is ``xs_env[e]``)."""
        return None if self.xs is None else (self.xs.transpose(1, 0, 2) if isinstance(self.xs, np.ndarray)
                                             else self.xs.permute(1, 0, 2))

    @property
    def us_env(self):
        return None if self.us is None else (self.us.transpose(1, 0, 2) if isinstance(self.us, np.ndarray)
                                             else self.us.permute(1, 0, 2))


class BatchedRollout:
    """A rollout plan: parameter structs and device buffers built once, launched many times.

    ``launch(x0_dev)`` is asynchronous on the current CUDA stream and returns device tensors;
    ``run_host(x0_host)`` is the end-to-end call with HOST buffers (pinned staging, H2D, launch, D2H)."""

    def __init__(self, dynamics, controller, n_envs: int, steps: int, integrator: str = "euler",
                 record_stride: int = 0, record_controls: bool = True, cost: Optional[RunningCost] = None,
                 box: Optional[Box] = None, fast_trig: Optional[bool] = None, want_steps: bool = False,
                 want_final: bool = True):
        torch = L.require_cuda()
        self.torch = torch
        self.dyn, self.ctl = dynamics, controller
        self.N, self.T = int(n_envs), int(steps)
        n, m = dynamics.get_dimension()
        self.n, self.m = n, m
        if integrator not in L.INTEGRATORS:
            raise ValueError(f"integrator must be one of {sorted(L.INTEGRATORS)}")
        self.sys_spec = dynamics.system_spec()
        if integrator == "discrete":
            if dynamics.KIND != L.SYS_LINEAR:
                raise ValueError("integrator='discrete' (exact zero-order hold) needs a LinearDynamics")
            Ad, Bd = dynamics.discretized()
            L.fill(self.sys_spec.A, Ad)
            L.fill(self.sys_spec.B, Bd)
        self.ctl_spec = controller.control_spec()
        self.cost_spec = cost.spec(n, m) if cost is not None else None
        self.opts = L.HjbRolloutOpts()
        self.opts.integrator = L.INTEGRATORS[integrator]
        self.opts.record_stride = int(record_stride)
        self.opts.fast_trig = int(dynamics.fast_trig if fast_trig is None else fast_trig)
        if box is not None:
            self.opts.box_enabled = 1
            L.fill(self.opts.box_xf, box.xf)
            L.fill(self.opts.box_lo, box.lo)
            L.fill(self.opts.box_hi, box.hi)
        f32 = dict(device="cuda", dtype=torch.float32)
        self.n_rec = self.T // record_stride if record_stride > 0 else 0
        self.xs = torch.empty((self.n_rec + 1, self.N, n), **f32) if record_stride > 0 else None
        self.us = torch.empty((self.n_rec, self.N, m), **f32) if (record_stride > 0 and record_controls) else None
        # want_final = False: a cost-only plan (the per-environment cost is the result; no final-state write-back, and
        # run_host brings back N floats instead of N (n + 1))
        self.x_final = torch.empty((self.N, n), **f32) if want_final else None
        self.cost = torch.empty((self.N,), **f32) if cost is not None else None
        self.steps = torch.empty((self.N,), device="cuda", dtype=torch.int32) if (want_steps or box is not None) else None
        self._x0_dev = None
        self._pinned = None

    def launch(self, x0_dev) -> RolloutResult:
        if x0_dev.shape != (self.N, self.n) or not x0_dev.is_cuda or x0_dev.dtype != self.torch.float32 \
                or not x0_dev.is_contiguous():
            raise ValueError(f"x0 must be a contiguous float32 CUDA tensor of shape {(self.N, self.n)}")
        import ctypes as C

        L.check(L.lib().hjb_rollout(self.sys_spec, self.ctl_spec,
                                    self.cost_spec if self.cost_spec is not None else C.POINTER(L.HjbCost)(),
                                    self.opts, L.ptr(x0_dev), self.N, self.T, L.ptr(self.xs), L.ptr(self.us),
                                    L.ptr(self.x_final), L.ptr(self.cost), L.ptr(self.steps), L.stream_ptr()),
                "hjb_rollout")
        return RolloutResult(self.xs, self.us, self.x_final, self.cost, self.steps)

    def kernel_variant(self) -> dict:
        """The kernel instantiation this plan dispatches to (``hjb_rollout_variant``)."""
        import ctypes as C

        out = (C.c_int32 * 6)()
        L.check(L.lib().hjb_rollout_variant(self.sys_spec, self.ctl_spec,
                                            self.cost_spec if self.cost_spec is not None else C.POINTER(L.HjbCost)(),
                                            self.opts, int(self.xs is not None or self.us is not None), out),
                "hjb_rollout_variant")
        return {"integrator": {v: k for k, v in L.INTEGRATORS.items()}[out[0]], "recorded": bool(out[1]),
                "cost_mode": ("none", "diagonal", "dense", "unit")[out[2]], "box": bool(out[3]),
                "fast_trig": bool(out[4]), "controller_clips": bool(out[5])}

    # -- end to end with host buffers ------------------------------------------------------------------
    def _staging(self):
        if self._pinned is None:
            t = self.torch
            self._x0_dev = t.empty((self.N, self.n), device="cuda", dtype=t.float32)
            self._pinned = {
                "x0": t.empty((self.N, self.n), dtype=t.float32, pin_memory=True),
                "x_final": t.empty((self.N, self.n), dtype=t.float32, pin_memory=True) if self.x_final is not None else None,
                "cost": t.empty((self.N,), dtype=t.float32, pin_memory=True) if self.cost is not None else None,
            }
        return self._pinned

    def h2d_bytes(self) -> int:
        return self.N * self.n * 4

    def d2h_bytes(self) -> int:
        return (self.N * self.n * 4 if self.x_final is not None else 0) + (self.N * 4 if self.cost is not None else 0)

    def _launch_range(self, x0_dev, lo: int, hi: int):
        """Envs [lo, hi) of a final-state(+cost) plan on the current stream (the outputs are env-major: a range of
        environments is a contiguous slice of every buffer)."""
        import ctypes as C

        sl = slice(lo, hi)
        L.check(L.lib().hjb_rollout(self.sys_spec, self.ctl_spec,
                                    self.cost_spec if self.cost_spec is not None else C.POINTER(L.HjbCost)(),
                                    self.opts, L.ptr(x0_dev[sl]), hi - lo, self.T, None, None,
                                    L.ptr(self.x_final[sl]) if self.x_final is not None else None,
                                    L.ptr(self.cost[sl]) if self.cost is not None else None,
                                    L.ptr(self.steps[sl]) if self.steps is not None else None, L.stream_ptr()),
                "hjb_rollout")

    def _pipeline(self):
        if getattr(self, "_pipe", None) is None:
            t = self.torch
            self._pipe = {"h2d": t.cuda.Stream(), "d2h": t.cuda.Stream(), "run": [t.cuda.Stream(), t.cuda.Stream()]}
        return self._pipe

    def run_host(self, x0_host, copy_in: bool = True, chunks: int = 32):
        """x0 on the host -> (x_final, cost) on the host.  Copies x0 into pinned staging unless it already IS
        the staging buffer (``copy_in=False`` after writing into ``self.pinned_x0()``), then H2D, launch, D2H of
        the final states and per-environment costs, and a synchronise.

        Final-state plans (record_stride = 0) are software-pipelined over ``chunks`` ranges of environments (measured on
        C4: 1 range 42.5 ms, 8 ranges 28.7 ms, 32 ranges 27.2 ms against 26.7 ms for the kernel alone): the
        H2D copy of range c + 1 and the D2H copy of range c - 1 run on their own streams under the kernel of range
        c (PCIe is full duplex), and consecutive ranges alternate between two launch streams so that the tail wave
        of one overlaps the head of the next.  Recorded trajectories are time-major, so those plans run as one
        launch."""
        t = self.torch
        pin = self._staging()
        if copy_in:
            pin["x0"].copy_(t.as_tensor(np.asarray(x0_host, dtype=np.float32)) if not isinstance(x0_host, t.Tensor)
                            else x0_host)
        chunks = max(1, min(int(chunks), self.N // 65536)) if self.xs is None else 1
        if chunks <= 1:
            self._x0_dev.copy_(pin["x0"], non_blocking=True)
            res = self.launch(self._x0_dev)
            if res.x_final is not None:
                pin["x_final"].copy_(res.x_final, non_blocking=True)
            if res.cost is not None:
                pin["cost"].copy_(res.cost, non_blocking=True)
            t.cuda.current_stream().synchronize()
            return pin["x_final"], pin["cost"]
        pipe = self._pipeline()
        cur = t.cuda.current_stream()
        start = t.cuda.Event()
        start.record(cur)
        for s in (pipe["h2d"], pipe["d2h"], *pipe["run"]):
            s.wait_event(start)                              # earlier work on the caller's stream comes first
        step = -(-self.N // chunks)
        step = -(-step // 256) * 256                         # whole CTAs per range
        for c, lo in enumerate(range(0, self.N, step)):
            hi = min(self.N, lo + step)
            sl = slice(lo, hi)
            run = pipe["run"][c & 1]
            with t.cuda.stream(pipe["h2d"]):
                self._x0_dev[sl].copy_(pin["x0"][sl], non_blocking=True)
                up = t.cuda.Event()
                up.record(pipe["h2d"])
            with t.cuda.stream(run):
                run.wait_event(up)
                self._launch_range(self._x0_dev, lo, hi)
                done = t.cuda.Event()
                done.record(run)
            with t.cuda.stream(pipe["d2h"]):
                pipe["d2h"].wait_event(done)
                if self.x_final is not None:
                    pin["x_final"][sl].copy_(self.x_final[sl], non_blocking=True)
                if self.cost is not None:
                    pin["cost"][sl].copy_(self.cost[sl], non_blocking=True)
        pipe["d2h"].synchronize()                            # the last D2H copy follows every launch and every H2D copy
        return pin["x_final"], pin["cost"]

    def run_seeded(self, seed: int, first: int = 0, chunks: int = 32):
        """The end-to-end call whose INPUT is a seed: the initial states (the dynamics' own x0 distribution) are generated
        on the device by the counter-based generator (``hjb_sample_states``: environment i is sample ``first + i`` of the
        stream keyed by ``seed``), range by range ahead of each range's kernel; the per-environment results come back to
        pinned host memory.  Nothing crosses PCIe on the way in — what a many-GPU box needs, where eight 400 MB host
        arrays per step share one root complex."""
        t = self.torch
        pin = self._staging()
        chunks = max(1, min(int(chunks), self.N // 65536)) if self.xs is None else 1
        if chunks <= 1:
            self.dyn.sample_initial_states(self.N, seed, first, out=self._x0_dev)
            res = self.launch(self._x0_dev)
            if res.x_final is not None:
                pin["x_final"].copy_(res.x_final, non_blocking=True)
            if res.cost is not None:
                pin["cost"].copy_(res.cost, non_blocking=True)
            t.cuda.current_stream().synchronize()
            return pin["x_final"], pin["cost"]
        pipe = self._pipeline()
        cur = t.cuda.current_stream()
        start = t.cuda.Event()
        start.record(cur)
        for s_ in (pipe["d2h"], *pipe["run"]):
            s_.wait_event(start)
        step = -(-self.N // chunks)
        step = -(-step // 256) * 256
        for c, lo in enumerate(range(0, self.N, step)):
            hi = min(self.N, lo + step)
            sl = slice(lo, hi)
            run = pipe["run"][c & 1]
            with t.cuda.stream(run):
                self.dyn.sample_initial_states(hi - lo, seed, first + lo, out=self._x0_dev[sl])
                self._launch_range(self._x0_dev, lo, hi)
                done = t.cuda.Event()
                done.record(run)
            with t.cuda.stream(pipe["d2h"]):
                pipe["d2h"].wait_event(done)
                if self.x_final is not None:
                    pin["x_final"][sl].copy_(self.x_final[sl], non_blocking=True)
                if self.cost is not None:
                    pin["cost"][sl].copy_(self.cost[sl], non_blocking=True)
        pipe["d2h"].synchronize()
        return pin["x_final"], pin["cost"]

    def pinned_x0(self):
        return self._staging()["x0"]


def rollout(dynamics, controller, x0, steps: int, integrator: str = "euler", record_stride: int = 1,
            record_controls: bool = True, cost: Optional[RunningCost] = None, box: Optional[Box] = None,
            fast_trig: Optional[bool] = None) -> RolloutResult:
    """Closed-loop rollout of ``controller`` on ``dynamics`` from every row of ``x0`` ([N, n] or [n]).

    NumPy ``x0`` -> NumPy results (float32); CUDA tensor ``x0`` -> CUDA tensors (no host traffic)."""
    torch = L.require_cuda()
    on_device = isinstance(x0, torch.Tensor) and x0.is_cuda
    x0d = L.dev_f32(x0, (-1, dynamics.state_dim))
    plan = BatchedRollout(dynamics, controller, x0d.shape[0], steps, integrator, record_stride, record_controls,
                          cost, box, fast_trig)
    res = plan.launch(x0d)
    if on_device:
        return res
    to_np = lambda v: None if v is None else v.cpu().numpy()
    return RolloutResult(to_np(res.xs), to_np(res.us), to_np(res.x_final), to_np(res.cost), to_np(res.steps))

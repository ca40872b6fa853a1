"""Bang-bang controllers of the double integrator's minimum-time comparison, compiled into the rollout kernel
(reference: examples/double_integrator_optimal_time.ipynb cells 18-21 — the notebook compares the learned policy with
the saturated LQR, the analytic time-optimal law and a policy read from a level-set solver's value function, stepping
ten trajectories through Python functions; here every environment of a launch runs the law in the step loop).

``SwitchingCurveController``   cell 18 ``get_analytical_control``
``GridPolicyController``       cell 18 ``get_level_set_control`` (nearest node of dV/dvel on a regular grid)
``time_to_goal``               cell 20's bookkeeping: the first step after which the state is inside the goal ball

    dyn = LinearDynamics(...)                       # A = [[0, 1], [0, 0]], B = [[0], [1]], dt = 0.01, |u| <= 1
    res = dyn.rollout(SwitchingCurveController(dyn), x0, 500, integrator="discrete", record_stride=1)
    t = time_to_goal(res, dyn.dt, metric=1e-4)      # [N] seconds, 5.0 where the ball is never reached
"""
from __future__ import annotations

import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.controller.controller_basic import DeviceController


def _check_double_integrator_shape(dynamics):
    n, m = dynamics.get_dimension()
    if dynamics.KIND != L.SYS_LINEAR or (n, m) != (2, 1):
        raise ValueError("the minimum-time controllers act on a LinearDynamics with x = [pos, vel] and one input")


class SwitchingCurveController(DeviceController):
    """u = 0 inside x^T x <= metric; +amplitude when (vel < 0 and pos <= vel^2 / 2) or (vel >= 0 and pos < -vel^2 / 2);
    -amplitude otherwise (the notebook's ``get_analytical_control``, amplitude 1)."""

    def __init__(self, dynamics, metric: float = 1e-4, amplitude: float = 1.0) -> None:
        super().__init__()
        _check_double_integrator_shape(dynamics)
        self.dynamics, self.metric, self.amplitude = dynamics, float(metric), float(amplitude)

    def control_spec(self):
        c = L.HjbControl()
        c.kind, c.clip = L.CTL_SWITCH_CURVE, 0
        c.aux[0], c.aux[1] = self.metric, self.amplitude
        return c


class GridPolicyController(DeviceController):
    """u = -amplitude * sign(dV/dvel) read at the nearest node of a regular (vel, pos) grid, extrapolating by clamping —
    ``-np.sign(RegularGridInterpolator((vel, pos), dVdvel, method="nearest", bounds_error=False, fill_value=None)(flip(x)))``.

    ``dVdvel`` [nv, np] lives on the axes ``vel`` (nv nodes) and ``pos`` (np nodes), both equally spaced.
    ``from_value_function`` builds it the notebook's way: central differences of V over the velocity axis."""

    def __init__(self, dynamics, dVdvel, pos, vel, amplitude: float = 1.0) -> None:
        super().__init__()
        _check_double_integrator_shape(dynamics)
        self.dynamics, self.amplitude = dynamics, float(amplitude)
        self.table = np.ascontiguousarray(dVdvel, dtype=np.float32)
        pos, vel = np.asarray(pos, dtype=np.float64), np.asarray(vel, dtype=np.float64)
        if self.table.shape != (len(vel), len(pos)) or len(vel) < 2 or len(pos) < 2:
            raise ValueError("dVdvel must be [len(vel), len(pos)] with at least two nodes per axis")
        for ax in (pos, vel):
            d = np.diff(ax)
            if not (d > 0).all() or np.abs(d - d[0]).max() > 1e-9 * max(1.0, abs(d[0])):
                raise ValueError("the grid axes must be ascending and equally spaced")
        self.pos, self.vel = pos, vel
        self._dev = None

    @classmethod
    def from_value_function(cls, dynamics, V, pos, vel, amplitude: float = 1.0):
        """V [len(vel), len(pos)]: dV/dvel by central differences on vel[1:-1] (cell 18: ``diffVdiffvel_by_level_set``)."""
        V, vel = np.asarray(V, dtype=np.float64), np.asarray(vel, dtype=np.float64)
        dv = (vel[-1] - vel[0]) / (len(vel) - 1)
        return cls(dynamics, (V[2:, :] - V[:-2, :]) / (2 * dv), pos, vel[1:-1], amplitude)

    def control_spec(self):
        if self._dev is None:
            torch = L.require_cuda()
            self._dev = torch.as_tensor(self.table).cuda().contiguous()
        c = L.HjbControl()
        c.kind, c.clip = L.CTL_GRID_SIGN, 0
        c.aux[0] = self.pos[0]
        c.aux[1] = (len(self.pos) - 1) / (self.pos[-1] - self.pos[0])
        c.aux[2] = self.vel[0]
        c.aux[3] = (len(self.vel) - 1) / (self.vel[-1] - self.vel[0])
        c.aux[4] = self.amplitude
        c.ref, c.ref_steps, c.ref_offset = self._dev.data_ptr(), int(self.table.shape[0]), int(self.table.shape[1])
        return c


def time_to_goal(result, dt: float, metric: float = 1e-4, t_max=None):
    """Per-environment time to the goal ball of a recorded rollout (``record_stride=1``): ``k * dt`` for the first step k
    whose resulting state satisfies x^T x <= metric, ``t_max`` (default: steps * dt, the notebook's T) if there is none —
    ``optimal_t`` of the notebook's cell 20.  Runs on the device (``hjb_time_to_goal``); NumPy results in -> NumPy out."""
    torch = L.require_cuda()
    xs = result.xs
    if xs is None:
        raise ValueError("time_to_goal needs a recorded rollout (record_stride=1)")
    on_device = isinstance(xs, torch.Tensor) and xs.is_cuda
    xd = xs if on_device else torch.as_tensor(np.ascontiguousarray(xs, dtype=np.float32)).cuda()
    xd = xd.contiguous()
    rows, N, n = xd.shape
    t_max = (rows - 1) * float(dt) if t_max is None else float(t_max)
    out = torch.empty(N, device="cuda", dtype=torch.float32)
    L.check(L.lib().hjb_time_to_goal(L.ptr(xd), N, n, rows, float(metric), float(dt), t_max, L.ptr(out), L.stream_ptr()),
            "hjb_time_to_goal")
    return out if on_device else out.cpu().numpy().astype(np.float64)

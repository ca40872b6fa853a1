"""Hover LQR controllers for the planar quadrotor and the 10-D near-hover quadcopter:
u = clip(-K wrap(x - xf) + uf, umin, umax)  (reference: controller/quadrotors_model_based_controller.py:7-75).

``Quadrotors2DWaypointsPlanner`` (the reference's minimum-snap planner, :77-233) is a host-side linear solve, provided with
the reference's interface plus a batched ``plan(ts)``; ``Quadrotors2DTrackingController`` feeds that time-varying reference
to the rollout kernel: u_t = clip(u_ref(t) - K wrap(x - x_ref(t)))."""
import numpy as np

from q_learning_with_hjb_b200 import _lib as L
from q_learning_with_hjb_b200.controller.controller_basic import DeviceController, lqr_gain
from q_learning_with_hjb_b200.dynamics.quadrotors import NearHoverQuadcopter, Quadrotors2D


class _HoverLQR(DeviceController):
    def _finish(self, dynamics, xf, Q, R, n_pos):
        self.dynamics = dynamics
        self.xf, self.Q, self.R = np.asarray(xf), np.asarray(Q), np.asarray(R)
        self.umin, self.umax = dynamics.get_control_limit()
        if np.linalg.norm(self.xf[n_pos:]) > 0:
            raise ValueError("Final Velocity or Angle is not zero")

    def control_spec(self):
        c = L.HjbControl()
        c.kind, c.clip = L.CTL_FEEDBACK, 1
        L.fill(c.K, self.K)
        L.fill(c.xf, self.xf)
        L.fill(c.uf, self.uf)
        return c


class Quadrotors2DHoveringController(_HoverLQR):
    def __init__(self, dynamics: Quadrotors2D, xf, Q, R) -> None:
        super().__init__()
        self._finish(dynamics, xf, Q, R, n_pos=2)
        d = dynamics
        self.uf = d.m * d.g / 2 * np.ones(2)
        # hover linearisation (:25-31): d(ddx)/dtheta = -g, thrust sum drives ddy, thrust difference dtheta
        self.A = np.zeros((6, 6))
        self.A[:3, 3:] = np.eye(3)
        self.A[3, 2] = -d.g
        self.B = np.zeros((6, 2))
        self.B[4, :] = 1.0 / d.m
        self.B[5, :] = [d.r / d.I, -d.r / d.I]
        self.K, self.P = lqr_gain(self.A, self.B, self.Q, self.R)


class NearHoverQuadcopterHoveringController(_HoverLQR):
    def __init__(self, dynamics: NearHoverQuadcopter, xf, Q, R) -> None:
        super().__init__()
        self._finish(dynamics, xf, Q, R, n_pos=3)
        d = dynamics
        self.uf = np.array([d.g * d.m / d.kT, 0.0, 0.0])
        # hover linearisation (:58-68): tan(theta) ~ theta
        self.A = np.zeros((10, 10))
        self.A[:5, 5:] = np.eye(5)
        self.A[5, 3] = self.A[6, 4] = d.g
        self.B = np.zeros((10, 3))
        self.B[7, 0] = d.kT / d.m
        self.B[8, 1] = self.B[9, 2] = d.n0
        self.K, self.P = lqr_gain(self.A, self.B, self.Q, self.R)


class Quadrotors2DWaypointsPlanner:
    """Minimum-snap trajectory through way-points and its differential-flatness lift to states and inputs of the planar
    quadrotor (reference: controller/quadrotors_model_based_controller.py:77-233).

    Segment i is a 7th-order polynomial in the time since the segment began; segment times are distance / avg_speed.
    Constraints (8 per segment): end points fixed with zero velocity, acceleration and jerk; interior way-points fixed on
    both sides; derivatives 1..6 continuous across them.  ``coeff[d, i, k]`` multiplies t^k of coordinate d on segment i.
    """

    ORDER = 7

    def __init__(self, waypoints: np.ndarray, dynamics: Quadrotors2D, avg_speed=0.25, exact_theta_ddot: bool = False) -> None:
        """``exact_theta_ddot``: the reference's hand-derived second derivative of theta (:201-202) drops one factor of
        the third derivative of y in its last term (``2 x_ddot y_dddot / b^3`` where the derivative of atan2 gives
        ``2 x_ddot y_dddot^2 / b^3``), so its feed-forward torque is off by ~0.1-0.3 % and the planned (x, u) is not
        exactly a trajectory of the model.  False (default) reproduces the reference; True uses the consistent
        derivative."""
        self.exact_theta_ddot = bool(exact_theta_ddot)
        waypoints = np.asarray(waypoints, dtype=np.float64)
        if waypoints.ndim != 2 or waypoints.shape[1] not in (2, 3):
            raise ValueError("The waypoints dim should either be 2D or 3D")
        self.dim = waypoints.shape[1]
        self.points, self.dynamics, self.avg_speed = waypoints, dynamics, avg_speed
        self.points_num = waypoints.shape[0]
        self.displacements = np.diff(waypoints, axis=0)
        self.distants = np.linalg.norm(self.displacements, axis=1)
        self.interval_t = self.distants / self.avg_speed
        self.cumulated_t = np.concatenate([[0.0], np.cumsum(self.interval_t)])
        self.coeff = self.solve_minimum_snap_coefficient()

    def get_polynomial_term(self, t, n, order=7) -> np.ndarray:
        """z with d^n/dt^n (c . [1, t, ..., t^order]) = c . z: z_k = k (k-1) ... (k-n+1) t^(k-n) for k >= n, else 0."""
        k = np.arange(order + 1)
        falling = np.array([np.prod(np.arange(i, i - n, -1.0)) if i >= n else 0.0 for i in k])
        return falling * np.power(float(t), np.maximum(k - n, 0))

    def solve_minimum_snap_coefficient(self) -> np.ndarray:
        S, W = self.points_num - 1, self.ORDER + 1
        rows, rhs = [], []

        def row(seg, t, n, sign=1.0):
            r = np.zeros(S * W)
            r[seg * W:(seg + 1) * W] = sign * self.get_polynomial_term(t, n, self.ORDER)
            return r

        # the two ends: position, then rest (velocity, acceleration, jerk zero)
        rows += [row(0, 0.0, 0), row(S - 1, self.interval_t[-1], 0)]
        rhs += [self.points[0], self.points[-1]]
        for n in (1, 2, 3):
            rows += [row(0, 0.0, n), row(S - 1, self.interval_t[-1], n)]
            rhs += [np.zeros(self.dim)] * 2
        # interior way-points: reached by the segment that ends there and by the one that starts there ...
        for i in range(S - 1):
            rows += [row(i, self.interval_t[i], 0), row(i + 1, 0.0, 0)]
            rhs += [self.points[i + 1]] * 2
        # ... with derivatives 1..6 continuous
        for i in range(S - 1):
            for n in range(1, 7):
                rows.append(row(i, self.interval_t[i], n) + row(i + 1, 0.0, n, -1.0))
                rhs.append(np.zeros(self.dim))
        A, b = np.stack(rows), np.stack(rhs)                       # [8 S, 8 S], [8 S, dim]
        return np.linalg.solve(A, b).T.reshape(self.dim, S, W)

    def _flat_to_state_input(self, dyn, d):
        """d[k] = (x^(k), y^(k)) for k = 0..4 (arrays broadcast over time).  With a = x'', b = y'' + g the thrust direction
        gives theta = -atan2(a, b); total thrust m sqrt(a^2 + b^2); the torque follows from theta''."""
        (x, y), (xd, yd), (a, ydd), (ad, bd), (add, bdd) = d
        b = ydd + dyn.g
        rho = a * a + b * b
        cross, dot = ad * b - a * bd, a * ad + b * bd
        theta = -np.arctan2(a, b)
        theta_d = -cross / rho
        theta_dd = -((add * b - a * bdd) * rho - 2.0 * cross * dot) / (rho * rho)
        if not self.exact_theta_ddot:   # the reference's last term: 2 x_ddot y_dddot / b^3 instead of ... y_dddot^2 / b^3
            theta_dd = theta_dd + (b * b / rho) * (2.0 * a / (b * b * b)) * (bd * bd - bd)
        thrust = dyn.m * np.sqrt(rho)
        torque = dyn.I / dyn.r * theta_dd
        return np.stack([x, y, theta, xd, yd, theta_d], axis=-1), np.stack([(thrust + torque) / 2, (thrust - torque) / 2], axis=-1)

    def flat_output_to_full_states_and_inputs_2D(self, t, coeff):
        """([x, y, theta, dx, dy, dtheta], [u1, u2]) at time ``t`` since the start of the segment with coefficients
        ``coeff`` [2, order + 1]."""
        order = coeff.shape[1] - 1
        d = [tuple(float(np.dot(coeff[c], self.get_polynomial_term(t, n, order))) for c in range(2)) for n in range(5)]
        state, u = self._flat_to_state_input(self.dynamics, [tuple(np.asarray(v) for v in pair) for pair in d])
        return state, u

    def update(self, t):
        """Reference state and feed-forward input at time ``t``; past the last way-point: hover there."""
        if self.dim != 2:
            raise NotImplementedError
        seg = int(np.searchsorted(self.cumulated_t, t, side="right")) - 1
        if seg >= self.points_num - 1:
            coeff = np.zeros((2, 1))
            coeff[:, 0] = self.points[-1]
            return self.flat_output_to_full_states_and_inputs_2D(0.0, coeff)
        return self.flat_output_to_full_states_and_inputs_2D(t - self.cumulated_t[seg], self.coeff[:, seg, :])

    def plan(self, ts):
        """``update`` for a whole time grid at once: states [T, 6], inputs [T, 2] (the time-varying reference of a tracking
        rollout)."""
        if self.dim != 2:
            raise NotImplementedError
        ts = np.asarray(ts, dtype=np.float64)
        seg = np.clip(np.searchsorted(self.cumulated_t, ts, side="right") - 1, 0, self.points_num - 2)
        tau = np.where(ts >= self.cumulated_t[-1], self.interval_t[-1], ts - self.cumulated_t[seg])   # hold the end point
        k = np.arange(self.ORDER + 1)
        d = []
        for n in range(5):
            falling = np.array([np.prod(np.arange(i, i - n, -1.0)) if i >= n else 0.0 for i in k])
            basis = falling * np.power(tau[:, None], np.maximum(k - n, 0))                         # [T, 8]
            vals = np.einsum("dtk,tk->dt", self.coeff[:, seg, :], basis)
            if n > 0:
                vals = np.where(ts >= self.cumulated_t[-1], 0.0, vals)                               # at rest when hovering
            d.append((vals[0], vals[1]))
        return self._flat_to_state_input(self.dynamics, d)


class Quadrotors2DTrackingController(DeviceController):
    """Way-point tracking for the planar quadrotor: the hover LQR's feedback (reference :36-38) about the state and
    feed-forward input the minimum-snap planner returns for the current time (:77-233),

        x_ref, u_ref = planner.update(t);   u = clip(u_ref - K wrap(x - x_ref), umin, umax)

    (the reference ships the planner and the hover controller; this is the loop that joins them).  ``K`` is the hover gain
    for (Q, R).  The whole reference — ``planner.plan(i dt)``, i = 0 .. — is computed ONCE on the host in float64 and lives
    on the device as a table [steps][n + m]; the rollout kernel reads row t in step t (the same row for every
    environment), so N environments x T steps cost one planning pass, not N T planner calls.
    """

    def __init__(self, dynamics: Quadrotors2D, planner: "Quadrotors2DWaypointsPlanner", Q, R, settle_steps: int = 2) -> None:
        super().__init__()
        self.dynamics, self.planner = dynamics, planner
        hover = Quadrotors2DHoveringController(dynamics, np.zeros(6), Q, R)
        self.K, self.P, self.Q, self.R = hover.K, hover.P, hover.Q, hover.R
        self.umin, self.umax = dynamics.get_control_limit()
        dt = float(dynamics.dt)
        # rows 0 .. : t = i dt up to the end of the last segment, then `settle_steps` rows of hovering at the last
        # way-point (the kernel repeats the last row for later steps)
        self.steps = int(np.ceil(planner.cumulated_t[-1] / dt)) + 1 + int(settle_steps)
        self.ts = dt * np.arange(self.steps)
        self.x_ref, self.u_ref = planner.plan(self.ts)
        self.xf, self.uf = self.x_ref[-1], self.u_ref[-1]
        self._table = None

    def reference_table(self):
        """float32 CUDA tensor [steps, 8]: x_ref (6), u_ref (2) per step."""
        if self._table is None:
            torch = L.require_cuda()
            self._table = torch.as_tensor(np.concatenate([self.x_ref, self.u_ref], axis=1).astype(np.float32)).cuda().contiguous()
        return self._table

    def control_spec(self, offset: int = 0):
        c = L.HjbControl()
        c.kind, c.clip = L.CTL_TRACK, 1
        L.fill(c.K, self.K)
        L.fill(c.xf, self.xf)
        L.fill(c.uf, self.uf)
        tab = self.reference_table()
        c.ref, c.ref_steps, c.ref_offset = tab.data_ptr(), int(tab.shape[0]), int(offset)
        return c

    def get_control_efforts(self, x, t: float = 0.0):
        """u for state(s) ``x`` at time ``t`` (rounded to the step grid of the dynamics, like the per-step loop)."""
        step = int(round(float(t) / float(self.dynamics.dt)))
        return self._efforts(self.dynamics.system_spec(), self.control_spec(offset=max(step, 0)), x)

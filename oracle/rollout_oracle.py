"""TEST INFRASTRUCTURE — CPU oracle for the batched closed-loop rollout (hot path (a)).

This is a NumPy float64 restatement of the reference's rollout arithmetic, vectorised over a leading
environment axis.  It is the *checker* for the CUDA kernels in ``q_learning_with_hjb_b200/csrc`` and
the ``cpu_baseline`` leg of ``bench.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``
(cpu_baseline / ``--impl reference``) may import it; the product package never does.

Parity pin: ``tests/test_oracle_vs_reference.py`` runs this module against the UNMODIFIED reference
(loaded through ``oracle/ref_loader.py``) step by step, and ``tests/test_kat.py`` reproduces the
numbers printed in the reference's notebooks (K-CP, K-Q2, K-Q10, K-DI, K-ARE).

Every function cites the reference file:line it restates (paths under /root/reference).
The reference integrates with forward Euler only (dynamics/dynamics_basic.py:120); ``integrator="rk4"``
is the north-star extension: classical RK4 composed from the reference's ``dynamics_step`` with ``u``
held constant over the step and ``states_wrap`` applied once after the step.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

PI = np.pi
TWO_PI = 2.0 * np.pi

# Working precision.  float64 (the reference's NumPy branch) is what every parity check uses; bench.py's "best-effort CPU"
# leg switches to float32 — same arithmetic, half the memory traffic, NOT a parity oracle.
DT = np.float64


def set_dtype(dt):
    global DT
    DT = np.dtype(dt).type


def wrap_angle(a):
    """``np.remainder(a + pi, 2*pi) - pi`` — floor-mod into [-pi, pi)
    (dynamics/cartpole.py:60-64, dynamics/quadrotors.py:66-70, controller/acrobot_energy_shaping.py:6-7)."""
    return np.remainder(a + PI, TWO_PI) - PI


# ------------------------------------------------------------------------------------------------
# systems: x' = f(x) + g(x) u
# ------------------------------------------------------------------------------------------------

@dataclass
class OracleSystem:
    kind: str                      # linear | cartpole | acrobot | quad2d | quad10d
    n: int
    m: int
    dt: float
    umin: np.ndarray
    umax: np.ndarray
    par: dict = field(default_factory=dict)

    # angle components wrapped by states_wrap
    def wrap_index(self) -> Tuple[int, ...]:
        return {"linear": (), "cartpole": (1,), "acrobot": (0, 1), "quad2d": (2,), "quad10d": (3, 4)}[self.kind]

    def wrap(self, x: np.ndarray) -> np.ndarray:
        """states_wrap, returning a copy (linear.py:17-18, cartpole.py:52-64, acrobot.py:72-81,
        quadrotors.py:48-70,151-170)."""
        x = np.array(x, dtype=DT, copy=True)
        for i in self.wrap_index():
            x[..., i] = wrap_angle(x[..., i])
        return x

    def f_g(self, x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Batched ``get_control_affine_matrix``: x[B,n] -> f[B,n], g[B,n,m]."""
        x = np.asarray(x, dtype=DT)
        B = x.shape[0]
        p = self.par
        f = np.zeros((B, self.n))
        g = np.zeros((B, self.n, self.m))
        if self.kind == "linear":
            # dynamics/linear.py:20-22
            f = x @ np.asarray(p["A"], dtype=DT).T
            g[:] = np.asarray(p["B"], dtype=DT)[None]
        elif self.kind == "cartpole":
            # dynamics/cartpole.py:19-50 via the manipulator form dynamics_basic.py:64-94
            mc, mp, l, grav = p["mc"], p["mp"], p["l"], p["g"]
            th, dth = x[:, 1], x[:, 3]
            s, c = np.sin(th), np.cos(th)
            M11 = mc + mp
            M12 = mp * l * c
            M22 = mp * l ** 2
            det = M11 * M22 - M12 * M12
            # C dq + G
            h1 = -mp * l * dth * s * dth
            h2 = mp * grav * l * s
            # -M^-1 (C dq + G)
            f[:, 0] = x[:, 2]
            f[:, 1] = x[:, 3]
            f[:, 2] = -(M22 * h1 - M12 * h2) / det
            f[:, 3] = -(-M12 * h1 + M11 * h2) / det
            # M^-1 B, B = [1, 0]
            g[:, 2, 0] = M22 / det
            g[:, 3, 0] = -M12 / det
        elif self.kind == "acrobot":
            # dynamics/acrobot.py:39-58 via dynamics_basic.py:64-94
            M, Cdq, G = self.acrobot_terms(x)
            det = M[:, 0, 0] * M[:, 1, 1] - M[:, 0, 1] * M[:, 1, 0]
            h = Cdq + G
            f[:, 0] = x[:, 2]
            f[:, 1] = x[:, 3]
            f[:, 2] = -(M[:, 1, 1] * h[:, 0] - M[:, 0, 1] * h[:, 1]) / det
            f[:, 3] = -(-M[:, 1, 0] * h[:, 0] + M[:, 0, 0] * h[:, 1]) / det
            # B = [0, 1]
            g[:, 2, 0] = -M[:, 0, 1] / det
            g[:, 3, 0] = M[:, 0, 0] / det
        elif self.kind == "quad2d":
            # dynamics/quadrotors.py:17-46
            th = x[:, 2]
            f[:, 0:3] = x[:, 3:6]
            f[:, 4] = -p["g"]
            g[:, 3, :] = (-np.sin(th) / p["m"])[:, None]
            g[:, 4, :] = (np.cos(th) / p["m"])[:, None]
            g[:, 5, 0] = p["r"] / p["I"]
            g[:, 5, 1] = -p["r"] / p["I"]
        elif self.kind == "quad10d":
            # dynamics/quadrotors.py:118-149
            f[:, 0:5] = x[:, 5:10]
            f[:, 5] = p["g"] * np.tan(x[:, 3])
            f[:, 6] = p["g"] * np.tan(x[:, 4])
            f[:, 7] = -p["g"]
            g[:, 7, 0] = p["kT"] / p["m"]
            g[:, 8, 1] = p["n0"]
            g[:, 9, 2] = p["n0"]
        else:
            raise ValueError(self.kind)
        return f, g

    # -- acrobot helpers -----------------------------------------------------------------------
    def acrobot_terms(self, x):
        """M (B,2,2), C dq (B,2), G (B,2) of dynamics/acrobot.py:39-58."""
        p = self.par
        m1, m2, l1, l2, I1, I2, grav = p["m1"], p["m2"], p["l1"], p["l2"], p["I1"], p["I2"], p["g"]
        q1, q2, dq1, dq2 = x[:, 0], x[:, 1], x[:, 2], x[:, 3]
        a = m2 * l1 * l2 / 2
        c2, s2 = np.cos(q2), np.sin(q2)
        M = np.empty((x.shape[0], 2, 2))
        M[:, 0, 0] = I1 + I2 + m2 * l1 ** 2 + 2 * a * c2
        M[:, 0, 1] = I2 + a * c2
        M[:, 1, 0] = I2 + a * c2
        M[:, 1, 1] = I2
        Cdq = np.empty((x.shape[0], 2))
        Cdq[:, 0] = -2 * a * s2 * dq2 * dq1 - a * s2 * dq2 * dq2
        Cdq[:, 1] = a * s2 * dq1 * dq1
        G = np.empty((x.shape[0], 2))
        G[:, 0] = (m1 * l1 / 2 + m2 * l1) * grav * np.sin(q1) + m2 * grav * l2 / 2 * np.sin(q1 + q2)
        G[:, 1] = m2 * grav * l2 / 2 * np.sin(q1 + q2)
        return M, Cdq, G

    def acrobot_energy(self, x):
        """dynamics/acrobot.py:60-70."""
        p = self.par
        m1, m2, l1, l2, I1, I2, grav = p["m1"], p["m2"], p["l1"], p["l2"], p["I1"], p["I2"], p["g"]
        q1, q2, dq1, dq2 = x[:, 0], x[:, 1], x[:, 2], x[:, 3]
        c1, c2 = np.cos(q1), np.cos(q2)
        a = m2 * l1 * l2 / 2
        T1 = 0.5 * I1 * dq1 ** 2
        T2 = 0.5 * (m2 * l1 ** 2 + I2 + 2 * a * c2) * dq1 ** 2 + 0.5 * I2 * dq2 ** 2 + (I2 + a * c2) * dq1 * dq2
        U = -m1 * grav * l1 / 2 * c1 - m2 * grav * (l1 * c1 + l2 / 2 * np.cos(q1 + q2))
        return T1 + T2 + U

    # -- integrators ---------------------------------------------------------------------------
    def xdot(self, x, u):
        """dynamics_step (dynamics_basic.py:96-105): f + g @ u."""
        f, g = self.f_g(x)
        return f + np.einsum("bnm,bm->bn", g, u)

    def step(self, x, u, integrator="euler"):
        """``Dynamics.simulate`` (dynamics_basic.py:107-122): clip u, integrate one dt, wrap."""
        u = np.clip(np.asarray(u, dtype=DT), self.umin, self.umax)
        dt = self.dt
        if integrator == "euler":
            xn = x + self.xdot(x, u) * dt
        elif integrator == "rk4":
            k1 = self.xdot(x, u)
            k2 = self.xdot(x + 0.5 * dt * k1, u)
            k3 = self.xdot(x + 0.5 * dt * k2, u)
            k4 = self.xdot(x + dt * k3, u)
            xn = x + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        elif integrator == "discrete":
            # exact zero-order hold of a LINEAR system: examples/double_integrator_optimal_time.ipynb cell 4
            # (``scipy.signal.cont2discrete`` once, then ``dynamics_step(x, u) = Ad x + Bd u``)
            if self.kind != "linear":
                raise ValueError("integrator='discrete' needs a linear system")
            import scipy.signal
            A, B = np.asarray(self.par["A"], dtype=np.float64), np.asarray(self.par["B"], dtype=np.float64)
            Ad, Bd, *_ = scipy.signal.cont2discrete((A, B, np.eye(A.shape[0]), np.zeros((A.shape[0], B.shape[1]))), dt=dt)
            xn = x @ Ad.T.astype(DT) + u @ Bd.T.astype(DT)
        else:
            raise ValueError(integrator)
        return self.wrap(xn)


# ------------------------------------------------------------------------------------------------
# controllers
# ------------------------------------------------------------------------------------------------

@dataclass
class OracleController:
    """kind:
      feedback    u = -K wrap(x - xf) + uf, optionally clipped
                  (controller/lqr.py:29-30 [xf=0,uf=0,clip]; examples/cartpole_balancing.ipynb cell 4:24-25
                  [no clip]; controller/quadrotors_model_based_controller.py:36-38,73-75 [clip])
      cartpole_es controller/cartpole_energy_shaping.py:65-110
      acrobot_es  controller/acrobot_energy_shaping.py:74-121
      track       u_t = clip(u_ref(t) - K wrap(x - x_ref(t))): the feedback law of
                  controller/quadrotors_model_based_controller.py:36-38 about what Quadrotors2DWaypointsPlanner.update(t)
                  returns (:77-233); ``planner`` is an OraclePlanner, ``dt`` the step of the time grid
      switch_curve  examples/double_integrator_optimal_time.ipynb cell 18, ``get_analytical_control``: the double
                  integrator's time-optimal bang-bang law (x = [pos, vel]; 0 inside x^T x <= metric)
      grid_sign   same cell, ``get_level_set_control``: u = -sign(dV/dvel) read at the NEAREST node of a regular
                  (vel, pos) grid — scipy's RegularGridInterpolator(method="nearest", bounds_error=False,
                  fill_value=None), restated in closed form in ``nearest_node`` and pinned against scipy itself
                  (tests/test_kat.py); ``grid`` = the table [nv, np], ``grid_axes`` = (pos nodes, vel nodes)
    """
    kind: str
    K: Optional[np.ndarray] = None        # (m, n) LQR gain
    P: Optional[np.ndarray] = None        # (n, n) acrobot switch metric
    xf: Optional[np.ndarray] = None
    uf: Optional[np.ndarray] = None
    clip: bool = True
    Ke: Optional[np.ndarray] = None       # energy-shaping gains (3,)
    eps_energy: float = 1.0
    eps_state: float = 1.0
    eps: float = 1000.0
    planner: Optional["OraclePlanner"] = None
    metric: float = 1e-4                  # switch_curve: radius^2 of the goal ball
    amp: float = 1.0                      # switch_curve / grid_sign: |u|
    grid: Optional[np.ndarray] = None     # grid_sign: dV/dvel on the grid, [nv, np]
    grid_axes: Optional[Tuple[np.ndarray, np.ndarray]] = None

    def control(self, sys: OracleSystem, x: np.ndarray, t: float = 0.0) -> np.ndarray:
        x = np.asarray(x, dtype=DT)
        if self.kind == "switch_curve":
            p, v = x[:, 0], x[:, 1]
            plus = ((v < 0) & (p <= 0.5 * v ** 2)) | ((v >= 0) & (p < -0.5 * v ** 2))
            u = np.where(plus, self.amp, -self.amp)
            return np.where(p * p + v * v <= self.metric, 0.0, u)[:, None].astype(DT)
        if self.kind == "grid_sign":
            pos_axis, vel_axis = self.grid_axes
            g = np.asarray(self.grid)
            ip = nearest_node(x[:, 0], pos_axis)
            iv = nearest_node(x[:, 1], vel_axis)
            return (-self.amp * np.sign(g[iv, ip]))[:, None].astype(DT)
        if self.kind == "track":
            x_ref, u_ref = self.planner.update(t)
            dx = sys.wrap(x - x_ref)
            u = -dx @ np.asarray(self.K, dtype=DT).T + u_ref
            return np.clip(u, sys.umin, sys.umax)
        if self.kind == "feedback":
            dx = sys.wrap(x - self.xf)
            u = -dx @ np.asarray(self.K, dtype=DT).T + self.uf
            return np.clip(u, sys.umin, sys.umax) if self.clip else u
        if self.kind == "cartpole_es":
            p = sys.par
            xf = np.array([0.0, np.pi, 0.0, 0.0])
            dx = sys.wrap(x - xf)                                        # :75
            E = 0.5 * x[:, 3] ** 2 - np.cos(x[:, 1])                     # :90-95
            Ef = 0.5 * xf[3] ** 2 - np.cos(xf[1])
            de = E - Ef                                                   # :78
            near = (np.abs(de) < self.eps_energy) & (np.sqrt(dx[:, 1] ** 2 + dx[:, 3] ** 2) < self.eps_state)  # :79
            u_lqr = -dx @ np.asarray(self.K, dtype=DT).T          # :80
            u_bar = de * x[:, 3] * np.cos(x[:, 1])                        # :99
            Ke = np.asarray(self.Ke, dtype=DT)
            ddq1 = Ke[0] * (-x[:, 0]) + Ke[1] * (-x[:, 2]) + Ke[2] * u_bar  # :100
            ddq2 = -np.cos(x[:, 1]) / p["l"] * ddq1 - p["g"] * np.sin(x[:, 1]) / p["l"]  # :101
            u_es = (p["mc"] + p["mp"]) * ddq1 + p["mp"] * p["l"] * np.cos(x[:, 1]) * ddq2 \
                - p["mp"] * p["l"] * np.sin(x[:, 1]) * x[:, 3] ** 2       # :102-103
            u = np.where(near[:, None], u_lqr, u_es[:, None])
            return np.clip(u, sys.umin, sys.umax)                         # :86
        if self.kind == "acrobot_es":
            xf = np.array([np.pi, 0.0, 0.0, 0.0])
            d = x - xf
            dx = np.concatenate([wrap_angle(d[:, :2]), d[:, 2:]], axis=1)  # :109
            P = np.asarray(self.P, dtype=DT)
            quad = np.einsum("bi,ij,bj->b", dx, P, dx)                    # :114
            u_lqr = -dx @ np.asarray(self.K, dtype=DT).T          # :115
            M, Cdq, G = sys.acrobot_terms(x)                              # :83-86
            Ef = sys.acrobot_energy(xf[None])[0]
            ubar = (sys.acrobot_energy(x) - Ef) * x[:, 2]                 # :88
            Ks = np.asarray(self.Ke, dtype=DT)
            ddq2 = Ks[0] * (-wrap_angle(x[:, 1])) + Ks[1] * (-x[:, 3]) + Ks[2] * ubar   # :90
            h = G + Cdq
            u_sw = (M[:, 1, 1] - M[:, 0, 1] ** 2 / M[:, 0, 0]) * ddq2 + h[:, 1] - M[:, 1, 0] / M[:, 0, 0] * h[:, 0]  # :92
            u = np.where((quad < self.eps)[:, None], u_lqr, u_sw[:, None])
            return np.clip(u, sys.umin, sys.umax)                         # :119
        raise ValueError(self.kind)


def nearest_node(x, axis):
    """Index of the node of ``axis`` (ascending) that scipy's ``RegularGridInterpolator(method="nearest",
    bounds_error=False, fill_value=None)`` reads for coordinate ``x`` — its algorithm, restated: the cell is
    i = searchsorted(axis, x) - 1 clamped to [0, n - 2], the normalised distance y = (x - axis[i]) / (axis[i+1] - axis[i])
    (outside the axis y < 0 or y > 1: extrapolation), and the node is i when y <= 1/2, else i + 1.  On a regular axis this
    is ceil(t - 1/2) of the fractional index t, clamped — the form the CUDA controller evaluates."""
    axis = np.asarray(axis, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    i = np.clip(np.searchsorted(axis, x) - 1, 0, len(axis) - 2)
    y = (x - axis[i]) / (axis[i + 1] - axis[i])
    return np.where(y <= 0.5, i, i + 1).astype(np.int64)


class OraclePlanner:
    """Minimum-snap way-point planner and its differential-flatness lift for the planar quadrotor, restated from
    controller/quadrotors_model_based_controller.py:77-233 formula by formula (including the reference's hand-derived
    theta_ddot, :201-202, whose last term reads 2 x_ddot y_dddot / b^3).  ``update(t)`` -> (x_ref [6], u_ref [2])."""

    def __init__(self, waypoints, par, avg_speed=0.25):
        self.points = np.asarray(waypoints, dtype=np.float64)
        self.g, self.m, self.r, self.I = (float(par[k]) for k in ("g", "m", "r", "I"))
        n = self.points.shape[0]
        self.interval_t = np.sqrt(((self.points[1:] - self.points[:-1]) ** 2).sum(1)) / avg_speed      # :90-93
        self.cumulated_t = np.zeros(n)
        self.cumulated_t[1:] = np.cumsum(self.interval_t)                                                 # :94-95
        S = n - 1
        A = np.zeros((8 * S, 8 * S)); b = np.zeros((2, 8 * S))                                            # :107-109
        term = self.term
        A[0, :8] = term(0, 0); A[1, -8:] = term(self.interval_t[-1], 0)                                   # :113-117
        b[:, 0] = self.points[0]; b[:, 1] = self.points[-1]
        for k, nd in enumerate((1, 2, 3)):                                                                # :119-124
            A[2 + 2 * k, :8] = term(0, nd); A[3 + 2 * k, -8:] = term(self.interval_t[-1], nd)
        for i in range(S - 1):                                                                            # :127-132
            A[8 + 2 * i, 8 * i:8 * (i + 1)] = term(self.interval_t[i], 0)
            A[9 + 2 * i, 8 * (i + 1):8 * (i + 2)] = term(0, 0)
            b[:, 8 + 2 * i] = self.points[i + 1]; b[:, 9 + 2 * i] = self.points[i + 1]
        for i in range(S - 1):                                                                            # :135-153
            for k in range(6):
                row = 8 + (S - 1) * 2 + i * 6 + k
                A[row, 8 * i:8 * (i + 1)] = term(self.interval_t[i], k + 1)
                A[row, 8 * (i + 1):8 * (i + 2)] = -term(0, k + 1)
        self.coeff = np.stack([np.linalg.solve(A, b[d]).reshape(S, 8) for d in range(2)])                # :155-156

    @staticmethod
    def term(t, n, order=7):                                                                              # :160-176
        z = np.zeros(order + 1)
        for i in range(order + 1):
            if i - n >= 0:
                z[i] = t ** (i - n) * np.prod(np.arange(i, i - n, -1))
        return z

    def flat(self, t, coeff):                                                                             # :178-214
        o = coeff.shape[1] - 1
        d = [[float(np.dot(coeff[c], self.term(t, n, o))) for c in range(2)] for n in range(5)]
        (x, y), (xd, yd), (xdd, ydd), (x3, y3), (x4, y4) = d
        b = ydd + self.g
        theta = -np.arctan2(xdd, b)
        q = 1 + (xdd / b) ** 2
        w = x3 / b - xdd * y3 / b ** 2
        theta_d = -1 / q * w
        theta_dd = 2 * xdd / b / q ** 2 * w ** 2 - 1 / q * (x4 / b - x3 * y3 / b ** 2 - x3 * y3 / b ** 2 - xdd * y4 / b ** 2
                                                             + 2 * xdd * y3 / b ** 3)
        if not np.sin(theta) == 0:                                                                        # :207-212
            u1 = (self.I / self.r * theta_dd - self.m / np.sin(theta) * xdd) / 2
            u2 = (-self.I / self.r * theta_dd - self.m / np.sin(theta) * xdd) / 2
        else:
            u1 = (self.I / self.r * theta_dd + self.m / np.cos(theta) * b) / 2
            u2 = (-self.I / self.r * theta_dd + self.m / np.cos(theta) * b) / 2
        return np.array([x, y, theta, xd, yd, theta_d]), np.array([u1, u2])

    def update(self, t):                                                                                  # :216-231
        index = np.argwhere(t >= self.cumulated_t)[-1, 0]
        if index == self.cumulated_t.shape[0] - 1:
            coeff = np.zeros((2, 8)); coeff[:, 0] = self.points[-1]
            return self.flat(0.0, coeff)
        return self.flat(t - self.cumulated_t[index], self.coeff[:, index, :])


def control_scale(sys: OracleSystem, ctl: OracleController, x: np.ndarray) -> np.ndarray:
    """Per-sample magnitude of the largest intermediate term of the control law, [B] — the scale against which
    an fp32 evaluation of the law can be expected to be accurate (cancellation between large terms is a property
    of the law, not of the implementation).  Used only to normalise errors in the parity tests."""
    x = np.asarray(x, dtype=DT)
    if ctl.kind in ("switch_curve", "grid_sign"):
        return np.full(len(x), ctl.amp, dtype=DT)
    K = np.abs(np.asarray(ctl.K, dtype=DT))
    if ctl.kind == "feedback":
        dx = np.abs(sys.wrap(x - ctl.xf))
        return (dx @ K.T + np.abs(ctl.uf)).max(axis=1)
    if ctl.kind == "cartpole_es":
        p = sys.par
        xf = np.array([0.0, np.pi, 0.0, 0.0])
        lqr = (np.abs(sys.wrap(x - xf)) @ K.T)[:, 0]
        E = 0.5 * x[:, 3] ** 2 + 1 + 1
        Ke = np.abs(np.asarray(ctl.Ke, dtype=DT))
        a1 = Ke[0] * np.abs(x[:, 0]) + Ke[1] * np.abs(x[:, 2]) + Ke[2] * E * np.abs(x[:, 3])
        a2 = a1 / p["l"] + p["g"] / p["l"]
        es = (p["mc"] + p["mp"]) * a1 + p["mp"] * p["l"] * a2 + p["mp"] * p["l"] * x[:, 3] ** 2
        return np.maximum(lqr, es)
    if ctl.kind == "acrobot_es":
        xf = np.array([np.pi, 0.0, 0.0, 0.0])
        d = x - xf
        dx = np.abs(np.concatenate([wrap_angle(d[:, :2]), d[:, 2:]], axis=1))
        lqr = (dx @ K.T)[:, 0]
        M, Cdq, G = sys.acrobot_terms(x)
        pp = sys.par
        a = pp["m2"] * pp["l1"] * pp["l2"] / 2
        Emag = 0.5 * np.abs(M[:, 0, 0]) * x[:, 2] ** 2 + 0.5 * pp["I2"] * x[:, 3] ** 2 + np.abs(M[:, 0, 1] * x[:, 2] * x[:, 3]) \
            + 200.0 + 100.0
        Ks = np.abs(np.asarray(ctl.Ke, dtype=DT))
        a2 = Ks[0] * np.pi + Ks[1] * np.abs(x[:, 3]) + Ks[2] * Emag * np.abs(x[:, 2])
        hmag = np.abs(G) + np.abs(a * x[:, 3:4] * (2 * np.abs(x[:, 2:3]) + np.abs(x[:, 3:4]))) + np.abs(a * x[:, 2:3] ** 2)
        sw = np.abs(M[:, 1, 1]) * a2 + hmag[:, 1] + hmag[:, 0]
        return np.maximum(lqr, sw)
    raise ValueError(ctl.kind)


# ------------------------------------------------------------------------------------------------
# closed-loop rollout
# ------------------------------------------------------------------------------------------------

@dataclass
class OracleCost:
    """running cost l(x,u) = dx^T Q dx + (u-uf)^T R (u-uf), dx = wrap(x - xf)
    (controller/vhjb.py:162-165; same form in every notebook's ``running_cost``).  ``u`` is the
    controller's output as returned (before ``simulate`` clips it) — K-CP only reproduces that way."""
    Q: np.ndarray
    R: np.ndarray
    xf: np.ndarray
    uf: np.ndarray

    def running(self, sys: OracleSystem, x, u):
        dx = sys.wrap(x - self.xf)
        du = u - self.uf
        return np.einsum("bi,ij,bj->b", dx, self.Q, dx) + np.einsum("bi,ij,bj->b", du, self.R, du)


def rollout(sys: OracleSystem, ctl: OracleController, x0: np.ndarray, steps: int, integrator: str = "euler",
            record_stride: int = 1, cost: Optional[OracleCost] = None):
    """The reference's closed loop ``u = ctl(x); x = dyn.simulate(x, u)`` (scripts/test_vhjb_policy.py:146-151,
    controller/cartpole_energy_shaping.py:123-125, ...) over a batch of initial states.

    Returns ``xs [T_rec+1, N, n]`` (time-major; x0 first, then every ``record_stride``-th state),
    ``us [T_rec, N, m]`` (the controller output at the START of each recorded interval, i.e. at steps
    0, s, 2s, ...), ``x_final [N, n]`` and ``cost [N]`` (sum of l(x,u)*dt over all steps)."""
    x = np.array(x0, dtype=DT, copy=True)
    N = x.shape[0]
    xs, us = [x.copy()], []
    J = np.zeros(N, dtype=DT)
    for t in range(steps):
        u = ctl.control(sys, x, t * sys.dt) if ctl.kind == "track" else ctl.control(sys, x)
        if cost is not None:
            J += cost.running(sys, x, u) * sys.dt
        if record_stride and t % record_stride == 0 and t // record_stride < steps // record_stride:
            us.append(u.copy())
        x = sys.step(x, u, integrator)
        if record_stride and (t + 1) % record_stride == 0:
            xs.append(x.copy())
    xs = np.stack(xs) if record_stride else None
    us = np.stack(us) if (record_stride and us) else None
    return xs, us, x, J


# ------------------------------------------------------------------------------------------------
# host-side setup shared by tests (gains): SciPy ARE exactly as the reference calls it
# ------------------------------------------------------------------------------------------------

def lqr_gain(A, B, Q, R):
    """P = solve_continuous_are(A,B,Q,R); K = R^-1 B^T P (controller/lqr.py:25-26 and every model-based ctor)."""
    import scipy.linalg

    P = scipy.linalg.solve_continuous_are(A, B, Q, R)
    K = np.dot(scipy.linalg.inv(R), np.dot(B.T, P))
    return K, P


def cartpole_linearisation(par):
    """controller/cartpole_energy_shaping.py:21-45 (= examples/cartpole_balancing.ipynb cell 4:11-22 with
    xf[1] = 3.1415926 there)."""
    mc, mp, l, g = par["mc"], par["mp"], par["l"], par["g"]
    th = par.get("theta_f", np.pi)
    M = np.array([[mc + mp, mp * l * np.cos(th)], [mp * l * np.cos(th), mp * l ** 2]])
    pGpq = np.array([[0, 0], [0, -mp * g * l]])
    Alin = np.vstack([np.array([[0, 0, 1, 0], [0, 0, 0, 1]]),
                      np.hstack([-np.linalg.inv(M) @ pGpq, np.zeros((2, 2))])])
    Blin = np.hstack([np.zeros(2), np.linalg.inv(M) @ np.array([1, 0])]).reshape(4, 1)
    return Alin, Blin


def acrobot_linearisation(par):
    """controller/acrobot_energy_shaping.py:23-46."""
    m1, m2, l1, l2, I1, I2, g = par["m1"], par["m2"], par["l1"], par["l2"], par["I1"], par["I2"], par["g"]
    a = m2 * l1 * l2 / 2
    c2 = 1.0  # q2 = 0 at xf
    M = np.array([[I1 + I2 + m2 * l1 ** 2 + 2 * a * c2, I2 + a * c2], [I2 + a * c2, I2]])
    Minv = np.linalg.inv(M)
    pGpq1 = np.array([-m1 * g * l1 / 2 - m2 * g * l1 - m2 * g * l2 / 2, -m2 * g * l2 / 2])
    pGpq2 = np.array([-m2 * g * l2 / 2, -m2 * g * l2 / 2])
    Alin = np.vstack([np.array([0, 0, 1, 0]), np.array([0, 0, 0, 1]),
                      np.hstack([-Minv @ pGpq1.reshape(2, 1), -Minv @ pGpq2.reshape(2, 1), np.zeros((2, 2))])])
    Blin = np.hstack([np.zeros(2), Minv @ np.array([0, 1])]).reshape(4, 1)
    return Alin, Blin


def quad2d_hover_AB(par):
    """controller/quadrotors_model_based_controller.py:25-31."""
    A = np.vstack([np.hstack([np.zeros((3, 3)), np.eye(3)]), np.array([0, 0, -par["g"], 0, 0, 0]), np.zeros((2, 6))])
    B = np.vstack([np.zeros((4, 2)), np.ones((1, 2)) / par["m"], np.array([par["r"] / par["I"], -par["r"] / par["I"]])])
    return A, B


def quad10d_hover_AB(par):
    """controller/quadrotors_model_based_controller.py:58-68."""
    g = par["g"]
    A = np.vstack([np.hstack([np.zeros((5, 5)), np.eye(5)]),
                   np.array([0, 0, 0, g, 0, 0, 0, 0, 0, 0]),
                   np.array([0, 0, 0, 0, g, 0, 0, 0, 0, 0]),
                   np.zeros((3, 10))])
    B = np.vstack([np.zeros((7, 3)), np.array([par["kT"] / par["m"], 0, 0]),
                   np.array([0, par["n0"], 0]), np.array([0, 0, par["n0"]])])
    return A, B


# canonical parameter sets = the reference's gin files / module constants, float32-rounded where the
# reference's config dataclass casts to float32 (configs/dynamics/dynamics_config.py:15-21)
def _f32(v):
    return np.asarray(v, dtype=np.float32).astype(np.float64)


def std_system(kind: str) -> OracleSystem:
    if kind == "linear":      # configs/dynamics/linear.gin
        return OracleSystem("linear", 2, 1, 0.02, _f32([-5]), _f32([5]),
                            {"A": _f32([[0, 1], [0, 0]]), "B": _f32([[0], [1]])})
    if kind == "cartpole":    # configs/dynamics/cartpole.gin
        return OracleSystem("cartpole", 4, 1, 0.02, _f32([-10]), _f32([10]),
                            {"mc": 1, "mp": 0.1, "l": 1, "g": 9.81})
    if kind == "acrobot":     # dynamics/acrobot.py:7-16
        return OracleSystem("acrobot", 4, 1, 0.05, np.array([-25.0]), np.array([25.0]),
                            {"l1": 0.5, "l2": 1, "m1": 8, "m2": 8, "I1": 2, "I2": 8, "g": 10})
    if kind == "quad2d":      # configs/dynamics/quadrotors2D.gin
        return OracleSystem("quad2d", 6, 2, 0.05, _f32([-20, -20]), _f32([20, 20]),
                            {"m": 1, "r": 0.25, "g": 9.81, "I": 0.0625})
    if kind == "quad10d":     # configs/dynamics/near_hover_quadcopter.gin
        return OracleSystem("quad10d", 10, 3, 0.05, _f32([0, -10, -10]), _f32([14.715, 10, 10]),
                            {"g": 9.81, "m": 1, "kT": 0.91, "n0": 10})
    raise ValueError(kind)


# the way-points of the tracking tests / bench extras (a slalom; ~7.6 s at 0.5 m/s: 152 steps of dt = 0.05)
TRACK_WAYPOINTS = np.array([[0.0, 0.0], [1.0, 0.5], [2.0, -0.3], [2.5, 1.0]])
TRACK_SPEED = 0.5


def std_controller(kind: str, sys: OracleSystem) -> OracleController:
    """The model-based controller each BASELINE config pairs with ``sys``."""
    if kind == "lqr":             # A6 on linear.gin, Q=I,R=I
        K, P = lqr_gain(sys.par["A"], sys.par["B"], np.eye(2), np.eye(1))
        return OracleController("feedback", K=K, P=P, xf=np.zeros(2), uf=np.zeros(1), clip=True)
    if kind == "cartpole_lqr":    # A7 (notebook): xf uses 3.1415926, unclipped
        A, B = cartpole_linearisation({**sys.par, "theta_f": 3.1415926})
        K, P = lqr_gain(A, B, np.eye(4), np.eye(1))
        return OracleController("feedback", K=K, P=P, xf=np.array([0, 3.1415926, 0, 0]), uf=np.zeros(1), clip=False)
    if kind == "cartpole_es":     # A8
        A, B = cartpole_linearisation(sys.par)
        K, P = lqr_gain(A, B, np.eye(4), np.eye(1))
        return OracleController("cartpole_es", K=K, P=P, Ke=np.array([4.0, 4.0, 10.0]), eps_energy=1, eps_state=1)
    if kind == "acrobot_es":      # A9
        A, B = acrobot_linearisation(sys.par)
        K, P = lqr_gain(A, B, np.eye(4), np.eye(1))
        return OracleController("acrobot_es", K=K, P=P, Ke=np.array([1.0, 2.0, 1.0]), eps=1000)
    if kind == "quad2d_hover":    # A10
        A, B = quad2d_hover_AB(sys.par)
        K, P = lqr_gain(A, B, np.eye(6), np.eye(2))
        uf = sys.par["m"] * sys.par["g"] / 2 * np.ones(2)
        return OracleController("feedback", K=K, P=P, xf=np.zeros(6), uf=uf, clip=True)
    if kind == "quad2d_track":    # hover gain about the planner's time-varying reference
        A, B = quad2d_hover_AB(sys.par)
        K, P = lqr_gain(A, B, np.eye(6), np.eye(2))
        return OracleController("track", K=K, P=P, planner=OraclePlanner(TRACK_WAYPOINTS, sys.par, avg_speed=TRACK_SPEED))
    if kind == "quad10d_hover":   # A10
        A, B = quad10d_hover_AB(sys.par)
        K, P = lqr_gain(A, B, np.eye(10), np.eye(3))
        uf = np.array([sys.par["g"] * sys.par["m"] / sys.par["kT"], 0, 0])
        return OracleController("feedback", K=K, P=P, xf=np.zeros(10), uf=uf, clip=True)
    raise ValueError(kind)

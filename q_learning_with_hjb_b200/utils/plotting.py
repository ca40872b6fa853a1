"""Trajectory animation for the ``plot_trajectory`` methods of the dynamics classes (reference:
dynamics/{cartpole,acrobot,quadrotors}.py ``plot_trajectory``).  Off the hot path and matplotlib-optional: the geometry
(``*_frame`` functions: state -> polylines) is plain NumPy and testable without matplotlib; ``animate`` imports
matplotlib only when called."""
from __future__ import annotations

import numpy as np


def cartpole_frame(x, l, cart_width=0.4, cart_height=0.2):
    """Cart outline and pole of x = [p, theta, ...]; theta = 0 hangs down, theta = pi is upright."""
    p, th = float(x[0]), float(x[1])
    w, h = cart_width / 2, cart_height / 2
    cart = np.array([[p - w, -h], [p + w, -h], [p + w, h], [p - w, h], [p - w, -h]])
    pole = np.array([[p, 0.0], [p + l * np.sin(th), -l * np.cos(th)]])
    return [cart, pole]


def acrobot_frame(x, l1, l2):
    """Two links of x = [q1, q2, ...]; q1 = 0 hangs down, q2 is relative to link 1."""
    q1, q2 = float(x[0]), float(x[1])
    knee = l1 * np.array([np.cos(q1 - np.pi / 2), np.sin(q1 - np.pi / 2)])
    toe = knee + l2 * np.array([np.cos(q1 + q2 - np.pi / 2), np.sin(q1 + q2 - np.pi / 2)])
    return [np.array([[0.0, 0.0], knee, toe])]


def quad2d_frame(x, r):
    """Body of the planar quadrotor x = [x, y, theta, ...]: a bar of half-length r tilted by theta, with two rotor stubs."""
    c = np.array([float(x[0]), float(x[1])])
    th = float(x[2])
    along, up = np.array([np.cos(th), np.sin(th)]), np.array([-np.sin(th), np.cos(th)])
    a, b = c - r * along, c + r * along
    return [np.array([a, b]), np.array([a, a + 0.3 * r * up]), np.array([b, b + 0.3 * r * up])]


def quad10d_frame(x, arm=0.25):
    """Cross of the near-hover quadcopter x = [px, py, pz, theta_x, theta_y, ...] in 3-D (small-angle body axes)."""
    p = np.asarray(x[:3], dtype=float)
    tx, ty = float(x[3]), float(x[4])
    ex = np.array([np.cos(ty), 0.0, -np.sin(ty)])
    ey = np.array([0.0, np.cos(tx), np.sin(tx)])
    return [np.array([p - arm * ex, p + arm * ex]), np.array([p - arm * ey, p + arm * ey])]


def animate(ts, xs, frame, xlim=None, ylim=None, zlim=None, three_d=False, trail=None):
    """matplotlib FuncAnimation over the rows of ``xs`` (one frame per row, titled with ``ts``) -> (anim, fig)."""
    try:
        import matplotlib.pyplot as plt
        from matplotlib.animation import FuncAnimation
    except ImportError as e:   # pragma: no cover
        raise ImportError("plot_trajectory needs matplotlib (optional dependency)") from e
    ts, xs = np.asarray(ts), np.asarray(xs)
    fig = plt.figure()
    ax = fig.add_subplot(projection="3d") if three_d else plt.axes()

    def draw(i):
        ax.clear()
        for poly in frame(xs[i]):
            ax.plot(*poly.T, "k-", linewidth=2)
        if trail is not None:
            ax.plot(*np.asarray([trail(x) for x in xs[: i + 1]]).T, "b:", linewidth=1)
        if xlim is not None:
            ax.set_xlim(*xlim)
        if ylim is not None:
            ax.set_ylim(*ylim)
        if three_d and zlim is not None:
            ax.set_zlim(*zlim)
        if not three_d:
            ax.set_aspect("equal")
        ax.set_title(f"t = {float(ts[i]):.2f} s")

    interval = 1000.0 * float(ts[1] - ts[0]) if len(ts) > 1 else 50.0
    anim = FuncAnimation(fig, draw, frames=len(ts), interval=interval, repeat=False)
    return anim, fig          # the reference's plot_trajectory returns this pair


def span(values, margin):
    lo, hi = float(np.min(values)), float(np.max(values))
    return lo - margin, hi + margin

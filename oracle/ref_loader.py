"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (``/root/reference``) through import stubs.

Only usable in the build container (``/root/reference`` does not exist on the GPU box). Used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by the
``not gpu`` tests that pin the restated oracle (``oracle/rollout_oracle.py``) to the reference itself.

Nothing in the product package imports this module.

How it works: the reference needs ``jax``, ``gin`` and ``matplotlib`` at import time;
``oracle/_stubs`` provides inert stand-ins so that every ``isinstance(x, jnp.ndarray)`` is False and
the reference runs its NumPy (float64) branch.  ``Acrobot.__init__`` is broken at HEAD
(``dynamics/acrobot.py:22`` calls ``super().__init__()`` without the required config), so the object is
built with ``__new__`` and the attributes its methods read are injected.
"""
from __future__ import annotations

import importlib
import os
import sys

def _default_root() -> str:
    """/root/reference in the build container; on the GPU box the travelling copy under baseline/_ref (made by
    tools/install_reference_baseline.py — used ONLY by bench.py's reference-literal CPU baseline)."""
    if os.path.isdir("/root/reference/dynamics"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


REFERENCE_ROOT = os.environ.get("HJB_REFERENCE_ROOT") or _default_root()
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")

_REF_TOPLEVEL = ("dynamics", "controller", "configs", "utils")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dynamics"))


def _ensure_path():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if _STUBS not in sys.path:
        sys.path.insert(0, _STUBS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)


def ref_import(module: str):
    """Import ``module`` (e.g. ``"dynamics.cartpole"``) from the reference tree."""
    _ensure_path()
    mod = importlib.import_module(module)
    origin = getattr(mod, "__file__", "") or ""
    if not origin.startswith(REFERENCE_ROOT):
        raise RuntimeError(f"{module} resolved to {origin}, not the reference tree")
    return mod


def gin_parse(relpath: str):
    """Parse a reference .gin file (path relative to the reference root) with the stub gin."""
    _ensure_path()
    import gin  # the stub

    gin.parse_config_file(os.path.join(REFERENCE_ROOT, relpath))


# --------------------------------------------------------------------------------------------
# constructors for the reference objects used by the rollout path
# --------------------------------------------------------------------------------------------

def make_cartpole():
    gin_parse("configs/dynamics/cartpole.gin")
    cfg = ref_import("configs.dynamics.dynamics_config").CartpoleDynamicsConfig()
    return ref_import("dynamics.cartpole").Cartpole(cfg)


def make_linear():
    gin_parse("configs/dynamics/linear.gin")
    cfg = ref_import("configs.dynamics.dynamics_config").LinearDynamicsConfig()
    return ref_import("dynamics.linear").LinearDynamics(cfg)


def make_quad2d():
    gin_parse("configs/dynamics/quadrotors2D.gin")
    cfg = ref_import("configs.dynamics.dynamics_config").Quadrotors2DConfig()
    return ref_import("dynamics.quadrotors").Quadrotors2D(cfg)


def make_quad10d():
    gin_parse("configs/dynamics/near_hover_quadcopter.gin")
    cfg = ref_import("configs.dynamics.dynamics_config").NearHoverQuadcopterConfig()
    return ref_import("dynamics.quadrotors").NearHoverQuadcopter(cfg)


def make_acrobot():
    """``Acrobot()`` cannot be constructed at HEAD (dynamics/acrobot.py:22); replay the rest of its ctor."""
    import numpy as np

    mod = ref_import("dynamics.acrobot")
    obj = mod.Acrobot.__new__(mod.Acrobot)
    p = mod.p
    obj.dim = 2
    obj.control_dim = 1
    obj.state_dim = 4
    obj.p = p
    obj.m1, obj.m2, obj.l1, obj.l2, obj.I1, obj.I2, obj.umax = (
        p["m1"], p["m2"], p["l1"], p["l2"], p["I1"], p["I2"], p["umax"])
    obj.g, obj.dt = p["g"], p["dt"]
    # Dynamics.simulate clips with self.umin/self.umax (dynamics_basic.py:118); umin is never set by
    # the reference ctor -- the controller clips to +-umax (acrobot_energy_shaping.py:119).
    obj.umin = -p["umax"]
    np.random.seed(0)
    return obj


class CachedLqrTerm:
    """Wrap an energy-shaping controller so ``get_lqr_term`` (constant) is solved once, not per step
    (cartpole_energy_shaping.py:77, acrobot_energy_shaping.py:112)."""

    def __init__(self, ctl):
        self.ctl = ctl
        K, P = ctl.get_lqr_term()
        ctl.get_lqr_term = lambda: (K, P)
        self.K, self.P = K, P

    def __getattr__(self, name):
        return getattr(self.ctl, name)


def reference_rollout(dyn, ctl_fn, x0, steps):
    """The reference's own closed loop (scripts/test_vhjb_policy.py:146-151): one env, forward Euler."""
    import numpy as np

    xs = [np.array(x0, dtype=np.float64)]
    us = []
    for _ in range(steps):
        u = np.atleast_1d(np.asarray(ctl_fn(xs[-1]), dtype=np.float64))
        us.append(u)
        xs.append(np.array(dyn.simulate(xs[-1].copy(), u), dtype=np.float64))
    return np.stack(xs), np.stack(us)

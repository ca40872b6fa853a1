"""Pins the restated rollout oracle to the UNMODIFIED reference (loaded through import stubs).

Runs only where /root/reference exists (the build container); the committed fixtures under
tests/golden/ carry the same evidence to the GPU box."""
import numpy as np
import pytest

from oracle import ref_loader as R
from oracle import rollout_oracle as O

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not R.available(), reason="reference tree not present")]


def _rand_states(sys, B, seed, scale=None):
    rng = np.random.default_rng(seed)
    scale = np.ones(sys.n) * 2.0 if scale is None else np.asarray(scale)
    return rng.uniform(-1, 1, size=(B, sys.n)) * scale


CASES = {
    "linear": (R.make_linear, None),
    "cartpole": (R.make_cartpole, [3, 4, 3, 5]),
    "acrobot": (R.make_acrobot, [4, 4, 6, 6]),
    "quad2d": (R.make_quad2d, [2, 2, 4, 3, 3, 3]),
    "quad10d": (R.make_quad10d, [2, 2, 2, 1.2, 1.2, 2, 2, 2, 2, 2]),
}


@pytest.mark.parametrize("kind", list(CASES))
def test_f_g_and_step_match_reference(kind):
    make, scale = CASES[kind]
    dyn = make()
    sys = O.std_system(kind)
    xs = _rand_states(sys, 64, 1, scale)
    rng = np.random.default_rng(2)
    us = rng.uniform(-1.5, 1.5, size=(64, sys.m)) * np.maximum(np.abs(sys.umin), np.abs(sys.umax))
    f, g = sys.f_g(xs)
    xn = sys.step(xs, us, "euler")
    for i in range(xs.shape[0]):
        fr, gr = dyn.get_control_affine_matrix(xs[i].copy())
        np.testing.assert_allclose(f[i], fr, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(g[i], np.asarray(gr).reshape(sys.n, sys.m), rtol=1e-12, atol=1e-12)
        xr = dyn.simulate(xs[i].copy(), us[i].copy())
        np.testing.assert_allclose(xn[i], xr, rtol=1e-12, atol=1e-12)
        if kind != "acrobot":  # Acrobot.states_wrap only accepts (4,) (acrobot.py:73)
            np.testing.assert_allclose(sys.wrap(xs[i:i + 1] * 3)[0], dyn.states_wrap(xs[i].copy() * 3), atol=1e-14)
        else:
            np.testing.assert_allclose(sys.wrap(xs[i:i + 1] * 3)[0], dyn.states_wrap(xs[i] * 3), atol=1e-14)


def _ref_controller(kind, dyn):
    if kind == "lqr":
        return R.ref_import("controller.lqr").LQR(dyn, np.eye(2), np.eye(1))
    if kind == "cartpole_es":
        return R.CachedLqrTerm(R.ref_import("controller.cartpole_energy_shaping").CartpoleEnergyShapingController(dyn))
    if kind == "acrobot_es":
        return R.CachedLqrTerm(R.ref_import("controller.acrobot_energy_shaping").AcrobotEnergyShapingController(dyn))
    mod = R.ref_import("controller.quadrotors_model_based_controller")
    if kind == "quad2d_hover":
        return mod.Quadrotors2DHoveringController(dyn, np.zeros(6), np.eye(6), np.eye(2))
    if kind == "quad10d_hover":
        return mod.NearHoverQuadcopterHoveringController(dyn, np.zeros(10), np.eye(10), np.eye(3))
    raise ValueError(kind)


CTL_CASES = [("linear", "lqr"), ("cartpole", "cartpole_es"), ("acrobot", "acrobot_es"),
             ("quad2d", "quad2d_hover"), ("quad10d", "quad10d_hover")]


@pytest.mark.parametrize("skind,ckind", CTL_CASES)
def test_controller_matches_reference(skind, ckind):
    make, scale = CASES[skind]
    dyn = make()
    sys = O.std_system(skind)
    ctl = O.std_controller(ckind, sys)
    rc = _ref_controller(ckind, dyn)
    xs = _rand_states(sys, 256, 3, scale)
    if ckind in ("cartpole_es", "acrobot_es"):  # make sure both branches are exercised
        xf = np.array([0, np.pi, 0, 0]) if ckind == "cartpole_es" else np.array([np.pi, 0, 0, 0])
        xs[:128] = xf + np.random.default_rng(4).uniform(-0.3, 0.3, size=(128, 4))
    u = ctl.control(sys, xs)
    for i in range(xs.shape[0]):
        ur = np.atleast_1d(rc.get_control_efforts(xs[i].copy()))
        np.testing.assert_allclose(u[i], ur, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(ctl.K, np.asarray(rc.K if ckind not in ("cartpole_es", "acrobot_es") else rc.K), rtol=1e-12)


@pytest.mark.parametrize("skind,ckind,steps", [("linear", "lqr", 300), ("cartpole", "cartpole_es", 400),
                                               ("acrobot", "acrobot_es", 120), ("quad2d", "quad2d_hover", 300),
                                               ("quad10d", "quad10d_hover", 300)])
def test_closed_loop_trajectory_matches_reference(skind, ckind, steps):
    make, _ = CASES[skind]
    dyn = make()
    sys = O.std_system(skind)
    ctl = O.std_controller(ckind, sys)
    rc = _ref_controller(ckind, dyn)
    if skind == "acrobot":
        x0 = np.array([[0.001, 0, 0, 0], [0.05, -0.02, 0.1, 0.0]])  # acrobot_energy_shaping.py:131
    else:
        x0 = np.stack([dyn.get_initial_state() for _ in range(3)])
    xs, us, xf, _ = O.rollout(sys, ctl, x0, steps, "euler", record_stride=1)
    for e in range(x0.shape[0]):
        xr, ur = R.reference_rollout(dyn, rc.get_control_efforts, x0[e], steps)
        # chaotic systems amplify 1e-16 rounding differences (closed-form M^-1 vs np.linalg.inv)
        tol = 1e-6 if skind == "acrobot" else 1e-9
        np.testing.assert_allclose(xs[:, e], xr, rtol=tol, atol=tol)
        np.testing.assert_allclose(us[:, e], ur, rtol=tol * 100, atol=tol * 100)


def test_config_classes_and_gin_files_bind_like_the_reference():
    """Drop-in boundary: every config class has the reference's fields in the reference's order, and every
    .gin file of this repo binds exactly the values the reference's file of the same name binds."""
    import ast
    import dataclasses
    import os
    import re
    from q_learning_with_hjb_b200.configs.controller import vhjb_controller_config as CC
    from q_learning_with_hjb_b200.configs.dynamics import dynamics_config as DC
    ref = "/root/reference/configs"
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "q_learning_with_hjb_b200", "configs")

    def ref_fields(path):            # {class: [field, ...]} from the reference's dataclass source (annotated assignments)
        out, tree = {}, ast.parse(open(path).read())
        for node in tree.body:
            if isinstance(node, ast.ClassDef):
                out[node.name] = ([b.id for b in node.bases if isinstance(b, ast.Name)],
                                  [s.target.id for s in node.body if isinstance(s, ast.AnnAssign)])
        return out

    for mod, path in ((DC, f"{ref}/dynamics/dynamics_config.py"), (CC, f"{ref}/controller/vhjb_controller_config.py")):
        classes = ref_fields(path)
        for name, (bases, own) in classes.items():
            inherited = [f for b in bases for f in classes.get(b, ([], []))[1]]
            ours = [f.name for f in dataclasses.fields(getattr(mod, name))]
            assert ours == inherited + own, name

    def bindings(path):
        out = {}
        for raw in open(path):
            line = raw.split("#", 1)[0].strip()
            if line:
                k, v = line.split("=", 1)
                out[k.strip()] = ast.literal_eval(v.strip())
        return out

    checked = 0
    for sub in ("dynamics", "controller"):
        for f in sorted(os.listdir(f"{ref}/{sub}")):
            if f.endswith(".gin"):
                assert bindings(os.path.join(pkg, sub, f)) == bindings(f"{ref}/{sub}/{f}"), f
                checked += 1
    assert checked == 7


def test_waypoints_planner_matches_the_reference():
    """Quadrotors2DWaypointsPlanner (host-side minimum-snap planner + differential flatness) against
    the reference class, run unmodified through the import stubs: coefficients, and reference state / feed-forward input
    along and past the trajectory; the batched plan(ts) against update(t)."""
    import importlib
    R.make_quad2d()                                              # sets up the stubs / sys.path for the reference tree
    ref_mod = importlib.import_module("controller.quadrotors_model_based_controller")
    from q_learning_with_hjb_b200.controller.quadrotors_model_based_controller import Quadrotors2DWaypointsPlanner
    dyn_ref = R.make_quad2d()
    for pts in (np.array([[0.0, 0.0], [1.0, 0.5], [2.0, -0.3], [2.5, 1.0]]), np.array([[0.0, 0.0], [0.7, 1.1]])):
        ref = ref_mod.Quadrotors2DWaypointsPlanner(pts, dyn_ref, avg_speed=0.4)
        ours = Quadrotors2DWaypointsPlanner(pts, dyn_ref, avg_speed=0.4)
        np.testing.assert_allclose(ours.cumulated_t, ref.cumulated_t, rtol=1e-14)
        np.testing.assert_allclose(ours.coeff, ref.coeff, rtol=1e-6, atol=1e-8 * np.abs(ref.coeff).max())
        ts = np.linspace(0.0, ref.cumulated_t[-1] * 1.2, 41)
        xs_plan, us_plan = ours.plan(ts)
        for i, t in enumerate(ts):
            xr, ur = ref.update(t)
            xo, uo = ours.update(t)
            np.testing.assert_allclose(xo, xr, rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(uo, ur, rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(xs_plan[i], xr, rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(us_plan[i], ur, rtol=1e-6, atol=1e-7)
        # with the consistent second derivative of theta (the reference's drops a factor in one term) the plan is a
        # trajectory of the model: x' = f(x) + g(x) u along it (central differences of the states); the reference's is not
        exact = Quadrotors2DWaypointsPlanner(pts, dyn_ref, avg_speed=0.4, exact_theta_ddot=True)
        h = 1e-5
        t0 = 0.37 * ref.cumulated_t[-1]
        xm, _ = exact.update(t0 - h); xp, _ = exact.update(t0 + h); x0, u0 = exact.update(t0)
        f, g = O.std_system("quad2d").f_g(x0[None])
        np.testing.assert_allclose((xp - xm) / (2 * h), f[0] + g[0] @ u0, rtol=1e-5, atol=1e-7)
        xr, ur = ref.update(t0)
        assert abs(((xp - xm) / (2 * h))[5] - (f[0] + g[0] @ ur)[5]) > 1e-5


def test_host_side_class_methods_match_the_reference_classes():
    """The drop-in classes against the reference's, method by method, for everything that stays on the host: manipulator
    matrices, energy, control limits, dimensions, the initial-state stream of NumPy's global RNG, NumPy ``states_wrap``
    (in place, returns its argument) and the linearisations behind the model-based gains."""
    from tests.helpers import make_dynamics
    rng = np.random.default_rng(7)
    makers = {"cartpole": R.make_cartpole, "acrobot": R.make_acrobot, "quad2d": R.make_quad2d, "quad10d": R.make_quad10d,
              "linear": R.make_linear}
    for kind, make in makers.items():
        ref = make()
        ours = make_dynamics(kind)
        n, m = ours.get_dimension()
        assert tuple(ref.get_dimension()) == (n, m)
        ulo, uhi = ours.get_control_limit()
        if kind != "acrobot":                             # (the reference's Acrobot never sets umin: acrobot.py:29-30)
            rlo, rhi = ref.get_control_limit()
            np.testing.assert_allclose(ulo, rlo); np.testing.assert_allclose(uhi, rhi)
        for _ in range(5):
            x = rng.uniform(-2, 2, size=n)
            if kind in ("cartpole", "acrobot"):
                np.testing.assert_allclose(ours.get_M(x), ref.get_M(x), rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(ours.get_C(x), ref.get_C(x), rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(ours.get_G(x), ref.get_G(x), rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(np.asarray(ours.get_B()).ravel(), np.asarray(ref.get_B()).ravel())
            if kind == "acrobot":
                assert abs(ours.energy(x) - ref.energy(x)) < 1e-10
            if kind in ("cartpole", "quad2d", "quad10d"):
                y = x * 3.0
                a, b = y.copy(), y.copy()
                ra, rb = ours.states_wrap(a), ref.states_wrap(b)
                np.testing.assert_allclose(ra, rb, rtol=1e-12, atol=1e-12)
                assert ra is a                                # NumPy input: wrapped in place and returned
        if kind != "acrobot":                             # both constructors seed NumPy's global RNG with the config's seed
            ref2, ours2 = make(), None
            seq_ref = [ref2.get_initial_state() for _ in range(4)]
            ours2 = make_dynamics(kind)
            seq_ours = [ours2.get_initial_state() for _ in range(4)]
            np.testing.assert_allclose(np.stack(seq_ours), np.stack(seq_ref), rtol=1e-12, atol=1e-12)
    # linearisations about the goal (used once, on the host, for every model-based gain) against finite differences of
    # the REFERENCE's dynamics_step
    for kind, xf, uf in (("cartpole", np.array([0, np.pi, 0, 0.0]), np.zeros(1)), ("quad2d", np.zeros(6), np.array([4.905, 4.905])),
                         ("quad10d", np.zeros(10), np.array([9.81 / 0.91, 0, 0]))):
        ref, ours = makers[kind](), make_dynamics(kind)
        A, B = ours.linearize(xf, uf)
        h = 1e-6
        n, m = ours.get_dimension()
        Afd = np.stack([(ref.dynamics_step(xf + h * e, uf) - ref.dynamics_step(xf - h * e, uf)) / (2 * h) for e in np.eye(n)], axis=1)
        Bfd = np.stack([(ref.dynamics_step(xf, uf + h * e) - ref.dynamics_step(xf, uf - h * e)) / (2 * h) for e in np.eye(m)], axis=1)
        np.testing.assert_allclose(A, Afd, atol=1e-6)
        np.testing.assert_allclose(B, Bfd, atol=1e-6)


def test_controller_gains_match_the_reference_classes():
    """Constructor-time quantities of the model-based controllers (solved once on the host) against the reference's own
    classes: LQR K / P, the hover controllers' K and trim input, the energy-shaping controllers' catch gains and energy."""
    import importlib
    from tests.helpers import make_controller, make_dynamics
    # LQR on the double integrator (controller/lqr.py:25-26)
    rdyn, odyn = R.make_linear(), make_dynamics("linear")
    rlqr = importlib.import_module("controller.lqr").LQR(rdyn, np.eye(2), np.eye(1))
    from q_learning_with_hjb_b200.controller.lqr import LQR
    olqr = LQR(odyn, np.eye(2), np.eye(1))
    np.testing.assert_allclose(olqr.K, rlqr.K, rtol=1e-10)
    np.testing.assert_allclose(olqr.P, rlqr.P, rtol=1e-10)
    # hover controllers (controller/quadrotors_model_based_controller.py:11-34, :40-71)
    rmod = importlib.import_module("controller.quadrotors_model_based_controller")
    from q_learning_with_hjb_b200.controller import quadrotors_model_based_controller as omod
    for kind, rmake, cls, n, m in (("quad2d", R.make_quad2d, "Quadrotors2DHoveringController", 6, 2),
                                   ("quad10d", R.make_quad10d, "NearHoverQuadcopterHoveringController", 10, 3)):
        rc = getattr(rmod, cls)(rmake(), np.zeros(n), np.eye(n), np.eye(m))
        oc = getattr(omod, cls)(make_dynamics(kind), np.zeros(n), np.eye(n), np.eye(m))
        np.testing.assert_allclose(oc.K, rc.K, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(oc.uf, rc.uf, rtol=1e-12)
        with pytest.raises(ValueError):
            getattr(omod, cls)(make_dynamics(kind), np.ones(n), np.eye(n), np.eye(m))     # moving goal: rejected like :20-21
    # energy-shaping controllers: the LQR catch (get_lqr_term) and the pole energy
    rcp = importlib.import_module("controller.cartpole_energy_shaping").CartpoleEnergyShapingController(R.make_cartpole())
    ocp = make_controller("cartpole_es", make_dynamics("cartpole"))
    for a, b in zip(ocp.get_lqr_term(), rcp.get_lqr_term()):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
    x = np.array([0.3, 2.5, -0.4, 1.2])
    assert abs(ocp.energy(x) - rcp.energy(x)) < 1e-14
    rac = importlib.import_module("controller.acrobot_energy_shaping").AcrobotEnergyShapingController(R.make_acrobot())
    oac = make_controller("acrobot_es", make_dynamics("acrobot"))
    for a, b in zip(oac.get_lqr_term(), rac.get_lqr_term()):
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10)


def test_riccati_for_the_terminal_cost_matches_the_reference_solver_in_float32():
    """controller/vhjb.py:156-160: P comes from the reference's OWN ordered-Schur solver (utils/utils.py:67-80) fed with
    float32 Alin, Blin (jax.jacobian of the float32 dynamics), Q, R (float32 config arrays).  The product's
    utils.solve_continuous_are computes in the inputs' type like the reference's: same P on the same float32 inputs."""
    from q_learning_with_hjb_b200.utils import utils as U
    rutils = R.ref_import("utils.utils")
    rng = np.random.default_rng(0)
    cases = [(np.array([[0, 1], [0, 0]]), np.array([[0], [1]]))]
    for n, m in ((4, 1), (6, 2), (10, 3)):
        cases.append((rng.normal(size=(n, n)), rng.normal(size=(n, m))))
    for A, B in cases:
        n, m = B.shape
        A32, B32, Q32, R32 = (np.asarray(v, dtype=np.float32) for v in (A, B, np.eye(n), np.eye(m)))
        ours = U.solve_continuous_are(A32, B32, Q32, R32)
        ref = rutils.solve_continuous_are(A32, B32, Q32, R32)
        assert ours.dtype == ref.dtype == np.float32
        np.testing.assert_allclose(ours, ref, rtol=2e-5, atol=2e-5)     # two single-precision Schur forms
        ours64 = U.solve_continuous_are(A, B, np.eye(n), np.eye(m))
        np.testing.assert_allclose(ours64, rutils.solve_continuous_are(np.asarray(A, float), np.asarray(B, float),
                                                                       np.eye(n), np.eye(m)), rtol=1e-9, atol=1e-9)


def test_tracking_loop_matches_the_reference_classes():
    """the reference ships the minimum-snap planner (quadrotors_model_based_controller.py:77-233) and the
    hover LQR (:7-38); the tracking loop joins them — x_ref, u_ref = planner.update(t); u = clip(u_ref - K wrap(x - x_ref));
    x = dynamics.simulate(x, u).  Oracle planner vs the reference's planner (bit for bit: same formulas), product planner
    vs both (1e-9), and the oracle's tracking rollout vs the loop run with the reference's own classes."""
    from q_learning_with_hjb_b200.controller import quadrotors_model_based_controller as QC
    from tests.helpers import make_dynamics
    rdyn = R.make_quad2d()
    rmod = R.ref_import("controller.quadrotors_model_based_controller")
    rplan = rmod.Quadrotors2DWaypointsPlanner(O.TRACK_WAYPOINTS, rdyn, avg_speed=O.TRACK_SPEED)
    rhover = rmod.Quadrotors2DHoveringController(rdyn, np.zeros(6), np.eye(6), np.eye(2))
    osys = O.std_system("quad2d")
    octl = O.std_controller("quad2d_track", osys)
    pplan = QC.Quadrotors2DWaypointsPlanner(O.TRACK_WAYPOINTS, make_dynamics("quad2d"), avg_speed=O.TRACK_SPEED)
    np.testing.assert_allclose(octl.K, rhover.K, rtol=1e-12)
    ts = np.arange(0, rplan.cumulated_t[-1] + 0.3, 0.05)
    px, pu = pplan.plan(ts)
    for i, t in enumerate(ts):
        xr, ur = rplan.update(t)
        xo, uo = octl.planner.update(t)
        np.testing.assert_array_equal(xo, xr)
        np.testing.assert_array_equal(uo, ur)
        np.testing.assert_allclose(px[i], xr, rtol=0, atol=1e-9)
        np.testing.assert_allclose(pu[i], ur, rtol=0, atol=1e-8)
    # the loop with the reference's own classes, three starts, against the oracle's batched rollout
    rng = np.random.default_rng(0)
    x0 = rng.uniform(-0.3, 0.3, size=(3, 6))
    T = len(ts) - 1
    xs, us, xf, _ = O.rollout(osys, octl, x0, T, "euler", record_stride=1)
    umin, umax = rdyn.get_control_limit()
    for e in range(3):
        x = x0[e].copy()
        for i in range(T):
            xr, ur = rplan.update(ts[i])
            u = np.clip(-rhover.K @ rdyn.states_wrap(x - xr) + ur, umin, umax)
            np.testing.assert_allclose(us[i, e], u, rtol=0, atol=1e-9)
            x = rdyn.simulate(x, u)
            np.testing.assert_allclose(xs[i + 1, e], x, rtol=0, atol=1e-9)
    # ... and it tracks: the end of the slalom is reached
    assert np.abs(xf[:, :2] - O.TRACK_WAYPOINTS[-1]).max() < 0.05

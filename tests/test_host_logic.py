"""CPU-side checks of the product package: config loading, gains, C-ABI struct packing, library exports.
No compute call is made (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.helpers import PAIRS, make_controller, make_dynamics, oracle_pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from q_learning_with_hjb_b200 import _lib as L
    header = open(os.path.join(ROOT, "include", "hjb_b200.h")).read()
    declared = set(re.findall(r"\b(hjb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in include/hjb_b200.h"
    assert declared == set(L.SYMBOLS), f"_lib.SYMBOLS out of sync with the header: {declared ^ set(L.SYMBOLS)}"
    handle = ctypes.CDLL(L.lib_path())
    for name in declared:
        assert hasattr(handle, name), f"{name} not exported by libhjb_b200.so"
    assert L.lib().hjb_abi_version() == 2
    assert L.lib().hjb_status_string(-2).startswith(b"hjb: unsupported")


def test_struct_sizes_match_header_layout():
    from q_learning_with_hjb_b200 import _lib as L
    assert ctypes.sizeof(L.HjbSystem) == 4 * (4 + 3 + 3 + 8 + 16 + 8)
    # (69 four-byte fields, padding to the pointer's alignment, the reference-table pointer, two int32)
    assert ctypes.sizeof(L.HjbControl) == 280 + 8 + 2 * 4 and L.HjbControl.ref.offset == 280
    assert ctypes.sizeof(L.HjbCost) == 4 * (100 + 9 + 10 + 3)
    assert ctypes.sizeof(L.HjbRolloutOpts) == 4 * (4 + 30)


@pytest.mark.parametrize("skind,ckind", PAIRS)
def test_gains_and_specs_match_oracle(skind, ckind):
    dyn = make_dynamics(skind)
    ctl = make_controller(ckind, dyn)
    osys, octl = oracle_pair(skind, ckind)
    spec = ctl.control_spec()
    n, m = dyn.get_dimension()
    np.testing.assert_allclose(np.array(spec.K[:n * m]).reshape(m, n), octl.K, rtol=2e-7)
    s = dyn.system_spec()
    assert (s.n, s.m) == (osys.n, osys.m) and abs(s.dt - osys.dt) < 1e-9
    np.testing.assert_allclose(np.array(s.umin[:m]), osys.umin, rtol=1e-7)
    np.testing.assert_allclose(np.array(s.umax[:m]), osys.umax, rtol=1e-7)


def test_initial_state_stream_matches_reference_semantics():
    # same RNG stream as Dynamics.get_initial_state (dynamics_basic.py:28-29): seed 0, U(-std, std) + mean, wrapped
    dyn = make_dynamics("cartpole")
    x = dyn.get_initial_state()
    np.testing.assert_allclose(x, [0.23430483, -3.12166627, 0.20552675, 0.00448832], atol=1e-8)
    assert x.dtype == np.float64


def test_states_wrap_numpy_is_in_place():
    dyn = make_dynamics("quad10d")
    x = np.arange(10, dtype=np.float64)
    y = dyn.states_wrap(x)
    assert y is x
    np.testing.assert_allclose(x[3:5], [3, 4 - 2 * np.pi])
    xb = np.tile(np.arange(10, dtype=np.float64), (3, 1))
    dyn.states_wrap(xb)
    np.testing.assert_allclose(xb[:, 4], 4 - 2 * np.pi)


def test_compute_calls_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    dyn = make_dynamics("cartpole")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dyn.simulate(np.zeros(4), np.zeros(1))
    ctl = make_controller("cartpole_es", dyn)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ctl.get_control_efforts(np.zeros(4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dyn.rollout(ctl, np.zeros((2, 4)), 10)


def test_hover_controller_rejects_moving_goal():
    from q_learning_with_hjb_b200.controller.quadrotors_model_based_controller import Quadrotors2DHoveringController
    dyn = make_dynamics("quad2d")
    with pytest.raises(ValueError):
        Quadrotors2DHoveringController(dyn, np.array([0, 0, 0.1, 0, 0, 0]), np.eye(6), np.eye(2))


def test_gin_files_bind_like_the_reference():
    from q_learning_with_hjb_b200.configs import gin_compat as gin
    from q_learning_with_hjb_b200.configs.controller.vhjb_controller_config import VHJBControllerConfig
    cfgdir = os.path.join(ROOT, "q_learning_with_hjb_b200", "configs", "controller")
    gin.parse_config_file(os.path.join(cfgdir, "quadrotors2DHovering_vhjb_controller.gin"))
    cfg = VHJBControllerConfig()
    assert list(cfg.features) == [128, 128, 64] and cfg.batch_size == 256 and cfg.epsilon == 1e-10
    assert cfg.Q.shape == (6, 6) and cfg.Q.dtype == np.float32
    np.testing.assert_allclose(cfg.uf, [4.905, 4.905], rtol=1e-6)


def test_plot_trajectory_geometry_and_optional_matplotlib():
    """plot_trajectory is kept for drop-in compatibility (off the hot path): the frame geometry is plain NumPy, and the
    methods exist on every system the reference animates; without matplotlib they raise ImportError, not AttributeError."""
    from q_learning_with_hjb_b200.utils import plotting as P
    cart, pole = P.cartpole_frame([0.5, np.pi, 0, 0], l=1.0)
    np.testing.assert_allclose(pole[1], [0.5, 1.0], atol=1e-12)            # theta = pi: upright
    (links,) = P.acrobot_frame([np.pi, 0.0, 0, 0], 0.5, 1.0)
    np.testing.assert_allclose(links[-1], [0.0, 1.5], atol=1e-12)          # both links up
    (links,) = P.acrobot_frame([0.0, 0.0, 0, 0], 0.5, 1.0)
    np.testing.assert_allclose(links[-1], [0.0, -1.5], atol=1e-12)         # hanging
    body = P.quad2d_frame([1.0, 2.0, 0.0, 0, 0, 0], 0.25)[0]
    np.testing.assert_allclose(body, [[0.75, 2.0], [1.25, 2.0]])
    assert len(P.quad10d_frame(np.zeros(10))) == 2
    for kind in ("cartpole", "acrobot", "quad2d", "quad10d"):
        dyn = make_dynamics(kind)
        assert callable(dyn.plot_trajectory)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            make_dynamics("cartpole").plot_trajectory(np.arange(3) * 0.1, np.zeros((3, 4)))


def test_sgdr_schedule_of_the_product_matches_the_oracle():
    """The termination-loss weight schedule of the product (controller/vhjb.py::sgdr_schedule, a host scalar handed to the
    kernels) against the oracle's restatement of optax.sgdr_schedule with the gin files' values (vhjb.py:123-126)."""
    from oracle import vhjb_oracle as V
    from q_learning_with_hjb_b200.controller.vhjb import sgdr_schedule
    for step in list(range(0, 4200, 37)) + [999, 1000, 1001, 1999, 2000, 2001, 19999, 20000, 20001, 10 ** 6]:
        ours = sgdr_schedule(step, init=0.0, peak=1e-5, end=0.0, cycles=10, warmup=1000, total=2000)
        assert abs(ours - V.sgdr_schedule(step)) < 1e-18, step


def test_header_is_plain_c_and_a_c_caller_links():
    """The boundary is a C ABI: include/hjb_b200.h compiles as C99 (no C++ or torch types in the signatures) and a C
    translation unit that takes the address of every declared entry point links against libhjb_b200.so (no call is made:
    there is no GPU here)."""
    import shutil
    import subprocess
    import tempfile
    from q_learning_with_hjb_b200 import _lib as L
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    header = os.path.join(ROOT, "include", "hjb_b200.h")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", header], check=True)
    names = sorted(L.SYMBOLS)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "caller.c")
        with open(src, "w") as fh:
            fh.write('#include "hjb_b200.h"\n#include <stdio.h>\nint main(void) {\n  const void* table[] = {\n')
            fh.write("".join(f"    (const void*)&{n},\n" for n in names))
            fh.write('  };\n  printf("%d\\n", (int)(sizeof(table) / sizeof(table[0])));\n  return hjb_abi_version() == 2 ? 0 : 1;\n}\n')
        exe = os.path.join(tmp, "caller")
        libdir = os.path.dirname(L.lib_path())
        subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", libdir, "-lhjb_b200",
                        f"-Wl,-rpath,{libdir}"], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip()
        assert int(out) == len(names)


def test_vectorised_initial_states_are_the_reference_stream():
    """Dynamics.get_initial_states(count) is ONE NumPy draw; it must produce the numbers of `count` successive
    get_initial_state() calls (dynamics_basic.py:28-29) — the notebook known-answer tests depend on the draw offsets."""
    from tests.helpers import make_dynamics
    for kind in ("cartpole", "quad2d", "quad10d", "linear"):
        dyn = make_dynamics(kind)
        np.random.seed(3)
        loop = np.stack([dyn.get_initial_state() for _ in range(257)])
        np.random.seed(3)
        batch = dyn.get_initial_states(257)
        np.testing.assert_array_equal(loop, batch)
        assert np.random.uniform() == np.random.RandomState(3).uniform(size=257 * dyn.state_dim + 1)[-1]   # same stream position
    assert make_dynamics("quad2d").get_initial_states(0).shape == (0, 6)


def test_counter_based_state_stream_twin():
    """oracle/x0_stream.py (the NumPy twin of hjb_sample_states): Philox4x32-10 against the published Random123 known
    answers; sample i does not depend on how the batch is split; distribution = wrap(U(-std, std) + mean)."""
    from oracle import x0_stream as X
    kat = [((0, 0), (0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff, 0xffffffff), (0xffffffff,) * 4, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for key, ctr, want in kat:
        got = X.philox4x32_10(key[0], key[1], *[np.array([c]) for c in ctr])
        assert tuple(int(v[0]) for v in got) == want
    mean, std = np.float32([0, 3.14, 0, 0]), np.float32([2.4, 0.05, 1, 0.05])
    whole = X.sample_states("cartpole", mean, std, 1234, 0, 5000)
    parts = np.concatenate([X.sample_states("cartpole", mean, std, 1234, lo, 1250) for lo in range(0, 5000, 1250)])
    np.testing.assert_array_equal(whole, parts)
    assert whole.dtype == np.float32 and not np.array_equal(whole, X.sample_states("cartpole", mean, std, 1235, 0, 5000))
    assert (np.abs(whole[:, 0]) <= 2.4).all() and (whole[:, 1] >= -np.pi).all() and (whole[:, 1] < np.pi).all()
    raw = np.where(whole[:, 1] < 0, whole[:, 1] + 2 * np.pi, whole[:, 1])          # un-wrap: 3.14 +- 0.05
    assert abs(raw.mean() - 3.14) < 2e-3 and raw.min() >= 3.09 - 1e-6 and raw.max() <= 3.19 + 1e-6
    assert abs(whole[:, 2].std() - 1 / np.sqrt(3)) < 0.02


def test_kernel_source_hash_is_independent_of_the_checkout_directory(tmp_path):
    """profiles/traffic.json stamps each ncu capture with the hash of the kernel sources; bench.py quotes the measured DRAM
    traffic only when that hash equals the hash of the sources it runs on — on the GPU box the repository lives under another
    path, so the hash must cover file names and contents only."""
    import shutil
    from q_learning_with_hjb_b200 import build
    a, b = tmp_path / "a", tmp_path / "some" / "where" / "else"
    a.mkdir()
    b.mkdir(parents=True)
    names = build.KERNEL_FAMILIES["rollout"]
    for d in (a, b):
        for f in names:
            shutil.copy(os.path.join(build.CSRC, f), d / f)
    assert build._digest([str(a / f) for f in names]) == build._digest([str(b / f) for f in names])
    assert build._digest([str(a / f) for f in names])[:16] == build.source_hash("rollout")


def test_bench_workloads_are_the_problems_the_tests_check():
    """bench.py builds what it times from q_learning_with_hjb_b200/workloads.py (nothing of oracle/ or tests/ on the measured
    arm); the parity tests build the same problems from the oracle's descriptions (tests/helpers_vhjb.py).  Same goal, same
    forms, the same random weights and the same synthetic batch, bit for bit."""
    from oracle import vhjb_oracle as V
    from q_learning_with_hjb_b200 import workloads as WL
    from tests import helpers_vhjb as H
    forms = {"clip": "clipped", "bang": "bangbang"}
    for name in ("linear", "cartpole", "quad2d", "quad10d", "di_mintime"):
        w, p = WL.vhjb_workload(name), H.problem(name)
        assert np.array_equal(w.xf, p.xf) and np.allclose(w.uf, p.uf, rtol=1e-15, atol=0) and w.act == p.act
        assert w.control_form == forms[p.control_form] and w.residual_form == p.residual_form
        assert w.eps == p.eps and w.eps_s == p.eps_s
        assert np.array_equal(p.Q, np.eye(p.sys.n)) and np.array_equal(p.R, np.eye(p.sys.m))
        for a, b in zip(WL.sample_vhjb_batch(name, 777, seed=5), H.sample_batch(name, 777, seed=5)):
            assert a.dtype == np.float32 and np.array_equal(a, b)
        for a, b in zip(WL.init_weights(p.sys.n, seed=3), V.init_weights(p.sys.n, seed=3)):
            assert a.dtype == np.float32 and np.array_equal(a, b.astype(np.float32))     # (the oracle keeps float64)
    assert np.array_equal(WL.flat_params(WL.init_weights(2)), H.flat_params(V.init_weights(2)))


def test_min_time_controllers_host_side():
    """controller/min_time.py without a device: argument checks, the notebook's central-difference table
    (double_integrator_optimal_time.ipynb cell 18) and the switching-curve controller's parameter block."""
    from q_learning_with_hjb_b200 import _lib as L
    from q_learning_with_hjb_b200.controller.min_time import GridPolicyController, SwitchingCurveController
    from tests.helpers import make_dynamics
    dyn = make_dynamics("linear")
    c = SwitchingCurveController(dyn, metric=1e-4, amplitude=1.0).control_spec()
    assert c.kind == L.CTL_SWITCH_CURVE and c.aux[0] == np.float32(1e-4) and c.aux[1] == 1.0
    with pytest.raises(ValueError):
        SwitchingCurveController(make_dynamics("cartpole"))
    pos, vel = np.linspace(-1, 1, 11), np.linspace(-2, 2, 21)
    V = np.add.outer(vel ** 2, pos)                                            # dV/dvel = 2 vel
    g = GridPolicyController.from_value_function(dyn, V, pos, vel)
    assert g.table.shape == (19, 11) and np.allclose(g.table, 2 * vel[1:-1, None], atol=1e-6)
    assert np.allclose(g.vel, vel[1:-1]) and np.allclose(g.pos, pos)
    with pytest.raises(ValueError):
        GridPolicyController(dyn, np.zeros((19, 11)), pos, vel)                 # axes / table mismatch
    with pytest.raises(ValueError):
        GridPolicyController(dyn, np.zeros((3, 3)), [0.0, 1.0, 3.0], [0.0, 1.0, 2.0])   # not equally spaced


def test_enum_values_of_the_header_match_the_ctypes_constants():
    """include/hjb_b200.h is the contract: the controller / system kinds the Python side passes must be the header's."""
    from q_learning_with_hjb_b200 import _lib as L
    header = open(os.path.join(ROOT, "include", "hjb_b200.h")).read()
    enums = {name: int(val) for name, val in re.findall(r"\b(HJB_(?:CTL|SYS|INT)_[A-Z0-9_]+)\s*=\s*(\d+)", header)}
    want = {"HJB_CTL_FEEDBACK": L.CTL_FEEDBACK, "HJB_CTL_CARTPOLE_ES": L.CTL_CARTPOLE_ES, "HJB_CTL_ACROBOT_ES": L.CTL_ACROBOT_ES,
            "HJB_CTL_TRACK": L.CTL_TRACK, "HJB_CTL_SWITCH_CURVE": L.CTL_SWITCH_CURVE, "HJB_CTL_GRID_SIGN": L.CTL_GRID_SIGN,
            "HJB_SYS_LINEAR": L.SYS_LINEAR}
    for name, value in want.items():
        assert enums.get(name) == value, (name, enums.get(name), value)
    for name, code in L.INTEGRATORS.items():
        assert enums["HJB_INT_" + name.upper()] == code, name

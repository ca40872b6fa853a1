"""The example scripts import without a GPU and their host-side pieces behave (the training itself is covered on the GPU by
tests/test_training_quality_gpu.py)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_examples_and_scripts_import():
    for name in ("onpolicy_hjb", "double_integrator_min_time", "cartpole_balancing", "quadcopter_10d", "drone_hovering"):
        mod = importlib.import_module(name)
        assert mod.__doc__ and "reference" in mod.__doc__
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    script = importlib.import_module("test_vhjb_policy")
    assert set(script.ENVS) == {"lqr", "cartpole", "quadrotors2DHovering"}      # the reference script's --env_name choices
    for dyn_gin, cfg_cls, ctl_gin in script.ENVS.values():
        assert os.path.exists(os.path.join(script.CONFIGS, "dynamics", dyn_gin))
        assert os.path.exists(os.path.join(script.CONFIGS, "controller", ctl_gin))
        assert hasattr(script.DC, cfg_cls)


def test_double_integrator_host_policies():
    """The comparison policies of the double-integrator example: the analytic switching curve reaches the origin from every
    start in the unit box, faster than the saturated LQR, in about the notebook's time (cell 21: 1.57 +- 0.54 s)."""
    D = importlib.import_module("double_integrator_min_time")
    x0 = np.random.default_rng(1).uniform(-1, 1, size=(200, 2))
    t_opt = D.time_to_origin(D.analytic_control, x0)
    t_lqr = D.time_to_origin(D.lqr_control(), x0)
    assert (t_opt < 5.0).all() and 1.3 < t_opt.mean() < 2.1
    assert t_lqr.mean() > 2.0 * t_opt.mean()
    # minimum time of the double integrator from rest at distance d: 2 sqrt(d); the sampled bang-bang law (dt = 0.01, target
    # radius 0.01) chatters along the switching curve and needs a little longer, never less
    t = D.time_to_origin(D.analytic_control, np.array([[0.64, 0.0], [-0.25, 0.0]]))
    ideal = np.array([1.6, 1.0])
    assert (t > ideal - 0.05).all() and (t < ideal + 0.4).all(), t

"""SURVEY.md 8d, CPU baseline (1), "reference-literal": the UNMODIFIED reference's own per-environment loop
    u = controller.get_control_efforts(x); x = dynamics.simulate(x, u)
loaded from /root/reference through the import stubs (oracle/ref_loader.py), one process per core.  Runs only where the
reference tree exists (the build container; it does not travel to the GPU box — bench.py's reference arm there is the NumPy
port).  Prints env-steps/s per core and for all cores."""
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np


def work(args):
    kind, envs, T = args
    from oracle import ref_loader as R
    import importlib
    if kind == "quad2d":
        dyn = R.make_quad2d()
        mod = importlib.import_module("controller.quadrotors_model_based_controller")
        ctl = mod.Quadrotors2DHoveringController(dyn, np.zeros(6), np.eye(6), np.eye(2))
    elif kind == "cartpole":
        dyn = R.make_cartpole()
        mod = importlib.import_module("controller.cartpole_energy_shaping")
        es = mod.CartpoleEnergyShapingController(dyn)
        K, _ = es.get_lqr_term()
        xf = np.array([0, np.pi, 0, 0])

        class Lqr:                                   # the notebook's LQR about xf (cartpole_balancing.ipynb cell 4)
            def get_control_efforts(self, x):
                return -K @ dyn.states_wrap(x - xf)
        ctl = Lqr()
    else:
        raise ValueError(kind)
    t0 = time.perf_counter()
    for _ in range(envs):
        x = dyn.get_initial_state()
        for _ in range(T):
            x = dyn.simulate(x, ctl.get_control_efforts(x))
    return envs * T / (time.perf_counter() - t0)


if __name__ == "__main__":
    from oracle import ref_loader as R
    if not R.available():
        raise SystemExit("reference tree not present")
    cores = os.cpu_count() or 1
    for kind, envs, T in (("quad2d", 16, 1000), ("cartpole", 16, 500)):
        with mp.get_context("spawn").Pool(cores) as pool:
            rates = pool.map(work, [(kind, envs, T)] * cores)
        print(f"{kind}: reference-literal loop, {cores} processes x {envs} envs x {T} steps: "
              f"{np.mean(rates):.3e} env-steps/s per core, {np.sum(rates):.3e} on {cores} cores")

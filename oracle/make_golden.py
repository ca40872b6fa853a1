"""TEST INFRASTRUCTURE — generates tests/golden/rollout_reference.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Every array below is produced by the reference's own classes (NumPy float64 branch, loaded through
oracle/ref_loader.py): per-state f, g, simulate(), get_control_efforts(), and closed-loop trajectories.
The fixtures travel to the GPU box, where /root/reference does not exist.
"""
import os

import numpy as np

from oracle import ref_loader as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "rollout_reference.npz")

SCALES = {
    "linear": [2, 2],
    "cartpole": [3, 4, 3, 5],
    "acrobot": [4, 4, 6, 6],
    "quad2d": [2, 2, 4, 3, 3, 3],
    "quad10d": [2, 2, 2, 1.2, 1.2, 2, 2, 2, 2, 2],
}
MAKERS = {"linear": R.make_linear, "cartpole": R.make_cartpole, "acrobot": R.make_acrobot,
          "quad2d": R.make_quad2d, "quad10d": R.make_quad10d}


def ref_controller(kind, dyn):
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


PAIRS = [("linear", "lqr", 300), ("cartpole", "cartpole_es", 500), ("acrobot", "acrobot_es", 100),
         ("quad2d", "quad2d_hover", 300), ("quad10d", "quad10d_hover", 300)]


def main():
    out = {}
    B = 96
    for skind, ckind, steps in PAIRS:
        dyn = MAKERS[skind]()
        n = len(SCALES[skind])
        m = 1 if skind in ("linear", "cartpole", "acrobot") else (2 if skind == "quad2d" else 3)
        rng = np.random.default_rng({"linear": 11, "cartpole": 12, "acrobot": 13, "quad2d": 14, "quad10d": 15}[skind])
        xs = rng.uniform(-1, 1, size=(B, n)) * np.asarray(SCALES[skind])
        if ckind == "cartpole_es":
            xs[:B // 2] = np.array([0, np.pi, 0, 0]) + rng.uniform(-0.4, 0.4, size=(B // 2, 4))
        if ckind == "acrobot_es":
            xs[:B // 2] = np.array([np.pi, 0, 0, 0]) + rng.uniform(-0.3, 0.3, size=(B // 2, 4))
        umax = np.broadcast_to(np.abs(np.asarray(dyn.umax, dtype=np.float64)), (m,))
        us = rng.uniform(-1.5, 1.5, size=(B, m)) * umax
        ctl = ref_controller(ckind, dyn)
        f = np.zeros((B, n)); g = np.zeros((B, n, m)); xn = np.zeros((B, n)); uc = np.zeros((B, m))
        for i in range(B):
            fi, gi = dyn.get_control_affine_matrix(xs[i].copy())
            f[i], g[i] = fi, np.asarray(gi).reshape(n, m)
            xn[i] = dyn.simulate(xs[i].copy(), us[i].copy())
            uc[i] = np.atleast_1d(ctl.get_control_efforts(xs[i].copy()))
        out[f"{skind}/x"], out[f"{skind}/u"] = xs, us
        out[f"{skind}/f"], out[f"{skind}/g"], out[f"{skind}/x_next"] = f, g, xn
        out[f"{skind}/{ckind}/u_ctl"] = uc
        out[f"{skind}/{ckind}/K"] = np.asarray(ctl.K, dtype=np.float64)
        # closed loop, 4 environments
        if skind == "acrobot":
            x0 = np.array([[0.001, 0, 0, 0], [0.05, -0.02, 0.1, 0.0], [-0.08, 0.03, 0.0, -0.05], [0.02, 0.02, 0.02, 0.02]])
        else:
            x0 = np.stack([dyn.get_initial_state() for _ in range(4)])
        trajs, ctrls = [], []
        for e in range(4):
            xr, ur = R.reference_rollout(dyn, ctl.get_control_efforts, x0[e], steps)
            trajs.append(xr); ctrls.append(ur)
        out[f"{skind}/{ckind}/traj_x"] = np.stack(trajs, axis=1)    # [T+1, 4, n] time-major
        out[f"{skind}/{ckind}/traj_u"] = np.stack(ctrls, axis=1)    # [T, 4, m]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE — generates tests/golden/tracking_reference.npz from the UNMODIFIED reference (build container only:
python -m oracle.make_golden_tracking).  The reference ships the minimum-snap planner and the hover LQR
(controller/quadrotors_model_based_controller.py:7-38, :77-233); the arrays are produced by ITS classes:
planner.update(t) on the step grid, and the closed loop  u = clip(u_ref - K wrap(x - x_ref)); x = dynamics.simulate(x, u)."""
import os

import numpy as np

from oracle import ref_loader as R
from oracle import rollout_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tracking_reference.npz")


def main():
    dyn = R.make_quad2d()
    mod = R.ref_import("controller.quadrotors_model_based_controller")
    plan = mod.Quadrotors2DWaypointsPlanner(O.TRACK_WAYPOINTS, dyn, avg_speed=O.TRACK_SPEED)
    hover = mod.Quadrotors2DHoveringController(dyn, np.zeros(6), np.eye(6), np.eye(2))
    T = int(np.ceil(plan.cumulated_t[-1] / dyn.dt)) + 20
    ts = dyn.dt * np.arange(T + 1)
    ref = [plan.update(t) for t in ts]
    umin, umax = dyn.get_control_limit()
    rng = np.random.default_rng(7)
    x0 = rng.uniform(-0.3, 0.3, size=(4, 6))
    x0[0] = 0.0
    trajs, ctrls = [], []
    for e in range(4):
        x, xs, us = x0[e].copy(), [x0[e].copy()], []
        for i in range(T):
            xr, ur = ref[i]
            u = np.clip(-hover.K @ dyn.states_wrap(x - xr) + ur, umin, umax)
            x = dyn.simulate(x, u)
            us.append(u); xs.append(np.array(x))
        trajs.append(np.stack(xs)); ctrls.append(np.stack(us))
    np.savez_compressed(OUT, waypoints=O.TRACK_WAYPOINTS, avg_speed=O.TRACK_SPEED, ts=ts, K=hover.K,
                        x_ref=np.stack([r[0] for r in ref]), u_ref=np.stack([r[1] for r in ref]),
                        traj_x=np.stack(trajs, axis=1), traj_u=np.stack(ctrls, axis=1))
    print("wrote", OUT, "T =", T)


if __name__ == "__main__":
    main()

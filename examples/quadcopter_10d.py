"""10-D near-hover quadcopter value learning, end to end on the GPU — the "ours" run of the reference's
examples/10D_quadcopte.ipynb (cells 3-10, 13-14): relu value net (the tcgen05 kernels' home configuration, BASELINE C5),
clipped control around the hover thrust, normalised HJB residual, 200 epochs of 20 on-policy trajectories (200 steps,
stopped when the drone is far away) + one shuffled pass of minibatches of 256.

The notebook prints (cell 10) a loss falling from 0.84 (epoch 10) through 0.18 (50), 0.055 (100), 0.032 (150) to 0.019
(190) while the policy only starts to keep the drone inside the observation box after ~150 epochs (collected trajectory
length 2 -> 15 -> 147 states), and (cell 14) "lqr cost 9.085334056081662" for the evaluation state drawn next.

    python examples/quadcopter_10d.py [--epochs 200]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import onpolicy_hjb as H  # noqa: E402


def make_problem():
    from q_learning_with_hjb_b200.configs import gin_compat as gin
    from q_learning_with_hjb_b200.configs.dynamics.dynamics_config import NearHoverQuadcopterConfig
    from q_learning_with_hjb_b200.dynamics.quadrotors import NearHoverQuadcopter
    import q_learning_with_hjb_b200 as pkg
    gin.parse_config_file(os.path.join(os.path.dirname(pkg.__file__), "configs", "dynamics", "near_hover_quadcopter.gin"))
    dyn = NearHoverQuadcopter(NearHoverQuadcopterConfig())
    uf = np.array([dyn.g * dyn.m / dyn.kT, 0, 0])
    p = H.Problem(dyn, np.zeros(10), uf, np.array([-2, -2, -2, -0.5, -0.5, -4, -4, -4, -2, -2.0]),
                  np.array([2, 2, 2, 0.5, 0.5, 4, 4, 4, 2, 2.0]), act="relu",
                  far_away=np.array([10, 10, 10, 4, 4, 20, 20, 20, 20, 20.0]))
    return p, p.kernels()


def evaluate(p, k, params, T=20.0):
    """Cells 13-14: closed-loop cost over 20 s from the next initial state, learned policy and (clipped) hover LQR."""
    from q_learning_with_hjb_b200.controller.controller_basic import lqr_gain
    A, B = p.dyn.linearize(p.xf, p.uf)
    K, _ = lqr_gain(A, B, np.eye(10), np.eye(3))
    x0 = p.dyn.get_initial_state()[None]
    steps = int(round(T / p.dyn.dt))
    return (H.closed_loop_cost(p, H.Policy(k, params), x0, steps)[0], H.closed_loop_cost(p, H.lqr_policy(p, K, clip=True), x0, steps)[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=200)
    args = ap.parse_args()
    p, k = make_problem()
    t0 = time.time()
    params, history = H.train(p, k, args.epochs)
    print(f"trained {args.epochs} epochs in {time.time() - t0:.1f} s")
    learned, lqr = evaluate(p, k, params)
    print("learned cost", learned)
    print("lqr cost", lqr)


if __name__ == "__main__":
    main()

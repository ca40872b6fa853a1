"""Cart-pole balancing by HJB value learning, end to end on the GPU — the "ours" experiment of the reference's
examples/cartpole_balancing.ipynb (cells 3-10, 15-16): a tanh value net V(x) = |MLP(z)|^2 + 1e-3 |z|^2, z = wrap(x - xf),
trained on the normalised HJB residual |dV/dx . (f + g u) / (l(x, u) + 1e-10) + 1| with u = clip(-R^-1 g^T dV/dx / 2).
Every epoch 20 closed-loop trajectories of 200 steps under the CURRENT policy are added to the data set (states inside the
observation box), then one shuffled pass of minibatches of 256 (Adam 1e-3).  All 20 trajectories advance together: one
value-net launch and one dynamics launch per step.

The notebook prints (cell 10) loss 0.177, 0.064, 0.040, 0.032, 0.024, 0.021, 0.017, 0.015, 0.013, 0.012 at epochs 10..100,
cumulated rollout costs between 4.9 and 9.4, trajectories of full length, and (cell 16) a mean closed-loop cost of the
learned policy of 9.1409 on ten initial states — the LQR's is 9.1410.

    python examples/cartpole_balancing.py [--epochs 100]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import onpolicy_hjb as H  # noqa: E402  (examples/ is on sys.path when run as a script or from the tests)

XF = np.array([0, 3.1415926, 0, 0])


def make_problem():
    from q_learning_with_hjb_b200.configs import gin_compat as gin
    from q_learning_with_hjb_b200.configs.dynamics.dynamics_config import CartpoleDynamicsConfig
    from q_learning_with_hjb_b200.dynamics.cartpole import Cartpole
    import q_learning_with_hjb_b200 as pkg
    gin.parse_config_file(os.path.join(os.path.dirname(pkg.__file__), "configs", "dynamics", "cartpole.gin"))
    dyn = Cartpole(CartpoleDynamicsConfig())          # seeds NumPy's global RNG with 0 like the reference
    p = H.Problem(dyn, XF, np.zeros(1), np.array([-4.8, -0.418, -1000, -1000]), np.array([4.8, 0.418, 1000, 1000]), act="tanh")
    return p, p.kernels()


def train(p, k, epochs=100, log=print):
    return H.train(p, k, epochs, log=log)


def make_soft_pd(p, seed=1):
    """The soft-PD baseline of cell 11: unconstrained tanh net with biases and a Dense(1) head; warm-up target z^T P z with
    the P of cell 4 (upright linearisation + Riccati)."""
    from q_learning_with_hjb_b200.controller.cartpole_energy_shaping import CartpoleEnergyShapingController
    from q_learning_with_hjb_b200.controller.soft_pd import SoftPDController
    K, P = CartpoleEnergyShapingController(p.dyn).get_lqr_term()
    return SoftPDController(p.dyn, XF, np.zeros(1), np.eye(4), np.eye(1), activation="tanh", normalized_residual=False,
                            K=K, P=P, seed=seed)


def train_soft_pd(p, ctl, epochs=100, warmup_epochs=20, log=print):
    """Cell 11 (``warmup_epochs=20``) / cell 12 (``warmup_epochs=0``).  The notebook prints, with the warm-up: cumulated cost
    82.0 at warm-up epoch 10, 5.66 at 20, then 5.3-7.4 with full-length trajectories; without it the policy never balances
    (costs 95-224, ~43 collected states per trajectory).  Like every on-policy run here it depends on the initialisation: seeds
    1 and 2 follow the notebook's curve (loss 0.40, 0.29, 0.19, 0.15, 0.13, 0.11, 0.096 at epochs 40..100 with seed 1, every
    trajectory of full length), seeds 0 and 3 lose some trajectories after epoch 40 — the unconstrained net has no
    positive-definiteness guarantee, which is the notebook's point."""
    return H.train_soft_pd(p, ctl, epochs, warmup_epochs=warmup_epochs, warmup_form="value_match", regularization=1.0, seed=1,
                           log=log)


def evaluate(p, k, params, notebook_draws=4000, soft=None):
    """Cell 16: closed-loop cost over 10 s from ten initial states, learned policy and LQR.  The notebook trains two more
    nets (2 x 100 epochs x 20 initial states) between cell 10 and cell 16; the same number of draws is skipped here so
    that the ten states are THE ten of the notebook (its LQR line reads 9.140986134043468)."""
    from q_learning_with_hjb_b200.controller.cartpole_energy_shaping import CartpoleEnergyShapingController
    dyn = p.dyn
    for _ in range(notebook_draws):
        dyn.get_initial_state()
    x0 = np.stack([dyn.get_initial_state() for _ in range(10)])
    steps = int(round(10 / dyn.dt))
    K, _ = CartpoleEnergyShapingController(dyn).get_lqr_term()     # the upright linearisation + Riccati of cell 4
    out = (H.closed_loop_cost(p, H.Policy(k, params), x0, steps), H.closed_loop_cost(p, H.lqr_policy(p, K), x0, steps))
    if soft is not None:           # cell 16 also prints "mean soft pd: 9.15915881211543"
        out += (H.closed_loop_cost(p, H.soft_pd_policy(soft), x0, steps),)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--soft-pd", action="store_true", help="also train the notebook's soft-PD baseline (cell 11)")
    args = ap.parse_args()
    p, k = make_problem()
    p.dyn.get_initial_state()                            # the notebook draws one initial state before training
    t0 = time.time()
    params, history = train(p, k, args.epochs)
    print(f"trained {args.epochs} epochs in {time.time() - t0:.1f} s")
    draws, soft = 4000, None
    if args.soft_pd:
        t0 = time.time()
        soft = make_soft_pd(p)
        train_soft_pd(p, soft, args.epochs)
        print(f"soft-PD: trained {args.epochs} epochs in {time.time() - t0:.1f} s")
        draws -= 20 * args.epochs
    costs = evaluate(p, k, params, notebook_draws=draws, soft=soft)
    print("mean pd: ", costs[0].mean())
    if soft is not None:
        print("mean soft pd: ", costs[2].mean())
    print("mean lqr: ", costs[1].mean())


if __name__ == "__main__":
    main()

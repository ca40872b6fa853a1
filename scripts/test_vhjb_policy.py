"""Train a vhjb controller from the gin configs, then roll the learned policy out next to the model-based one — the
reference's scripts/test_vhjb_policy.py (same flags: --env_name lqr | cartpole | quadrotors2DHovering, --dynamics_config,
--vhjb_controller_config), on the CUDA library: training is VHJBController.train() (fused updates, device replay buffer,
all of an epoch's trajectories rolled out together), the comparison loop calls the same per-state interface as the
reference's (get_control_efforts / simulate / running_cost).  Plots need matplotlib and are skipped without it
(or with --no-plot); the reference's closing pdb session is not reproduced.

    python scripts/test_vhjb_policy.py --env_name cartpole --no-plot
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from q_learning_with_hjb_b200.configs import gin_compat as gin  # noqa: E402
from q_learning_with_hjb_b200.configs.controller.vhjb_controller_config import VHJBControllerConfig  # noqa: E402
from q_learning_with_hjb_b200.configs.dynamics import dynamics_config as DC  # noqa: E402

CONFIGS = os.path.join(ROOT, "q_learning_with_hjb_b200", "configs")
ENVS = {   # env_name -> (dynamics gin, dynamics config class, controller gin)
    "lqr": ("linear.gin", "LinearDynamicsConfig", "linear_vhjb_controller.gin"),
    "cartpole": ("cartpole.gin", "CartpoleDynamicsConfig", "cartpole_vhjb_controller.gin"),
    "quadrotors2DHovering": ("quadrotors2D.gin", "Quadrotors2DConfig", "quadrotors2DHovering_vhjb_controller.gin"),
}


def load_systems(env_name, dynamics_config=None, vhjb_controller_config=None):
    """(dynamics, nn_policy, model_based_policy) as the reference's load_* functions build them (:20-130)."""
    from q_learning_with_hjb_b200.controller.vhjb import VHJBController
    dyn_gin, cfg_cls, ctl_gin = ENVS[env_name]
    gin.parse_config_file(dynamics_config or os.path.join(CONFIGS, "dynamics", dyn_gin))
    dcfg = getattr(DC, cfg_cls)()
    gin.parse_config_file(vhjb_controller_config or os.path.join(CONFIGS, "controller", ctl_gin))
    ccfg = VHJBControllerConfig()
    if env_name == "lqr":
        from q_learning_with_hjb_b200.controller.lqr import LQR
        from q_learning_with_hjb_b200.dynamics.linear import LinearDynamics
        dynamics = LinearDynamics(dcfg)
        model_based = LQR(dynamics, np.asarray(ccfg.Q), np.asarray(ccfg.R))         # same Q, R as the learned controller
    elif env_name == "cartpole":
        from q_learning_with_hjb_b200.controller.cartpole_energy_shaping import CartpoleEnergyShapingController
        from q_learning_with_hjb_b200.dynamics.cartpole import Cartpole
        dynamics = Cartpole(dcfg)
        model_based = CartpoleEnergyShapingController(dynamics, np.asarray(ccfg.Q), np.asarray(ccfg.R))
    else:
        from q_learning_with_hjb_b200.controller.quadrotors_model_based_controller import Quadrotors2DHoveringController
        from q_learning_with_hjb_b200.dynamics.quadrotors import Quadrotors2D
        dynamics = Quadrotors2D(dcfg)
        model_based = Quadrotors2DHoveringController(dynamics, np.asarray(ccfg.xf), np.asarray(ccfg.Q), np.asarray(ccfg.R))
    return dynamics, VHJBController(dynamics, ccfg), model_based


def test_policy(nn_policy, dynamics, model_based_controller, T=5, plot=True):
    """Side-by-side closed loops from one initial state (:132-225); returns (t, xs_learned, xs_model_based, cost_learned,
    cost_model_based)."""
    nn_policy.train_mode = False
    t_span = np.arange(0, T, dynamics.dt)
    n, m = dynamics.get_dimension()
    xs_mb, xs_nn = np.zeros((len(t_span), n)), np.zeros((len(t_span), n))
    us_mb, us_nn = np.zeros((len(t_span) - 1, m)), np.zeros((len(t_span) - 1, m))
    cost_mb, cost_nn = np.zeros(len(t_span) - 1), np.zeros(len(t_span) - 1)
    xs_mb[0] = xs_nn[0] = dynamics.get_initial_state()
    for i in range(1, len(t_span)):
        us_nn[i - 1] = nn_policy.get_control_efforts(xs_nn[i - 1])
        us_mb[i - 1] = model_based_controller.get_control_efforts(xs_mb[i - 1])
        xs_nn[i] = dynamics.simulate(xs_nn[i - 1], us_nn[i - 1])
        xs_mb[i] = dynamics.simulate(xs_mb[i - 1], us_mb[i - 1])
        cost_mb[i - 1] = nn_policy.running_cost(xs_mb[i - 1], us_mb[i - 1]) * dynamics.dt
        cost_nn[i - 1] = nn_policy.running_cost(xs_nn[i - 1], us_nn[i - 1]) * dynamics.dt
    if plot and hasattr(dynamics, "plot_trajectory"):
        try:
            dynamics.plot_trajectory(t_span, xs_nn)
        except ImportError:
            pass
    return t_span, xs_nn, xs_mb, cost_nn, cost_mb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env_name", default="lqr", choices=sorted(ENVS), help="Environment name")
    ap.add_argument("--dynamics_config", help="The path to the dynamics config")
    ap.add_argument("--vhjb_controller_config", help="The path to the config of vhjb controller")
    ap.add_argument("--no-plot", action="store_true")
    ap.add_argument("--epochs", type=int, default=0, help="override the config's number of epochs")
    args = ap.parse_args()
    dynamics, nn_policy, model_based = load_systems(args.env_name, args.dynamics_config, args.vhjb_controller_config)
    if args.epochs:
        nn_policy.epochs = args.epochs
    t0 = time.time()
    cost_mean, cost_std, lengths, total_loss, hjb_loss, term_loss = nn_policy.train()
    print(f"trained {nn_policy.epochs} epochs ({nn_policy.update_counter} updates) in {time.time() - t0:.1f} s; "
          f"final losses: total {total_loss[-1]:.5f}, hjb {hjb_loss[-1]:.5f}, termination {term_loss[-1]:.5f}")
    t, xs_nn, xs_mb, cost_nn, cost_mb = test_policy(nn_policy, dynamics, model_based, plot=not args.no_plot)
    print(f"closed-loop cost over {t[-1] + dynamics.dt:.1f} s from one initial state: learned {cost_nn.sum():.4f}, "
          f"model-based {cost_mb.sum():.4f}")
    if not args.no_plot:
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            return
        mean, std = np.array(cost_mean), np.array(cost_std)
        plt.figure(); plt.plot(mean, color="blue", label="average trajectory cost")
        plt.fill_between(range(len(mean)), mean + std, mean - std, color="lightblue", label="1-std")
        plt.xlabel("epochs"); plt.ylabel("average trajectory cost"); plt.legend(); plt.title("trajectory cost vs epoch")
        plt.figure()
        for series, label in ((total_loss, "total loss"), (hjb_loss, "hjb loss"), (term_loss, "termination loss")):
            plt.plot(series, label=label)
        plt.xlabel("epoch"); plt.ylabel("loss"); plt.legend(); plt.title("loss vs epoch")
        plt.show()


if __name__ == "__main__":
    main()

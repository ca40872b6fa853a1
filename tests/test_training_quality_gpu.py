"""End-to-end training on the GPU against numbers the REAL reference (JAX) printed.

examples/double_integrator_optimal_time.ipynb, cell 11 output, holds the loss the reference's own implementation reached
every 10 epochs when it trained its sin value net on the minimum-time HJB residual (65,536 states, batches of 256, Adam
1e-3, 100 epochs), and cell 21 the time the learned bang-bang policy needs to reach the origin.  JAX's PRNG (initial
weights) and torch's shuffle order cannot be reproduced here, so the comparison is statistical — but it is the only
place where numbers produced by the reference's JAX code pin the vhjb path (forward, input gradient, bang-bang control,
residual, parameter gradient, Adam) as a whole."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# cell 11 output of the notebook ("epoch:10, loss:0.07184618711471558, time to origin:4.253..."), epochs 10..100
NOTEBOOK_LOSS = [0.07184618711471558, 0.05127125605940819, 0.050909679383039474, 0.044496916234493256, 0.035330913960933685,
                 0.04974454268813133, 0.02857941947877407, 0.025896722450852394, 0.025598838925361633, 0.026374680921435356]
NOTEBOOK_LEARNED_TIME = 2.6149999999999998      # cell 21: "mean pd"
NOTEBOOK_LQR_TIME = 4.104000000000001           # cell 21: "mean lqr"
NOTEBOOK_ANALYTIC_TIME = 1.572                  # cell 21: "mean analytical set"


def test_double_integrator_training_reproduces_the_notebook():
    import torch
    assert torch.cuda.is_available()
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import double_integrator_min_time as D
    dyn, k = D.make_problem()
    params, losses = D.train(k, epochs=100, batch=256, log=None)
    ours = np.array(losses[9::10])
    ref = np.array(NOTEBOOK_LOSS)
    # the loss falls like the notebook's: every 10-epoch reading within a factor 1.6 (its own curve is that noisy: epoch 60
    # reads 0.050 between 0.035 and 0.029), the geometric mean of the ratio within 20 %
    ratio = ours / ref
    assert (ratio > 1 / 1.6).all() and (ratio < 1.6).all(), ratio
    assert abs(np.exp(np.log(ratio).mean()) - 1.0) < 0.2, ratio
    assert ours[-1] < 0.45 * ours[0]
    # the learned policy: clearly better than the saturated LQR, within reach of the analytic optimum, like the notebook's
    x0 = np.random.default_rng(1).uniform(-1, 1, size=(200, 2))
    t_learned = D.time_to_origin(D.learned_control(k, params), x0)
    t_lqr = D.time_to_origin(D.lqr_control(), x0)
    t_opt = D.time_to_origin(D.analytic_control, x0)
    assert t_opt.mean() < t_learned.mean() < 0.75 * t_lqr.mean()
    assert abs(t_learned.mean() / t_opt.mean() - NOTEBOOK_LEARNED_TIME / NOTEBOOK_ANALYTIC_TIME) < 0.45
    assert (t_learned < 15.0).mean() > 0.97            # (almost) every trajectory reaches the origin

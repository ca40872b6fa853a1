"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding by state and the two collectives of the vhjb
train step.  The per-shard arithmetic is done by the oracle here (no GPU in this container); what is under test is
q_learning_with_hjb_b200.parallel and the normalise-after-all-reduce algebra the CUDA path relies on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from q_learning_with_hjb_b200 import parallel


def test_shard_bounds_partition():
    for total in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 2)


def test_single_process_is_not_distributed():
    """No process group: VhjbKernels.train_step takes its one-call path (hjb_vhjb_train_step) and the collectives are no-ops."""
    assert parallel.dist_info() == (0, 1) and not parallel.is_distributed()
    t = torch.tensor([3.0, 5.0], dtype=torch.float64)
    assert parallel.global_counts(t.clone(), 0.5).tolist() == [3.5, 5.5]
    assert parallel.sum_across_ranks(t.clone()).tolist() == [3.0, 5.0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import vhjb_oracle as V
        from tests.helpers_vhjb import problem, sample_batch
        assert parallel.dist_info() == (rank, world) and parallel.is_distributed()   # -> the multi-GPU branch of train_step
        p = problem("quad2d")
        W = V.init_weights(p.sys.n, seed=3)
        xs, dones, costs = sample_batch("quad2d", 1000, seed=4)       # the GLOBAL batch, identical on every rank
        lo, hi = parallel.shard_bounds(len(xs), rank, world)
        d = torch.tensor(dones[lo:hi], dtype=torch.float64)
        # collective 1: done-counts -> global normalisers
        norm = parallel.global_counts(torch.stack([(1 - d).sum(), d.sum()]), p.eps)
        # this rank's shard, normalised by the GLOBAL counts (what hjb_vhjb_loss_grad does with `norm`)
        orc = V.VhjbOracle(p, W)
        q = orc.pieces(xs[lo:hi], create_graph=True)
        cst = torch.tensor(costs[lo:hi], dtype=torch.float64)
        hjb_sum = (q["r"].abs() * (1 - d)).sum()
        term_sum = ((q["V"] / (cst + p.eps) - 1).abs() * d).sum()
        reg = 0.3
        local = hjb_sum / norm[0] + reg * term_sum / norm[1]
        grads = torch.autograd.grad(local, orc.W)
        flat = torch.cat([g.reshape(-1) for g in grads] + [hjb_sum.detach().reshape(1), term_sum.detach().reshape(1)])
        # collective 2: gradient + loss sums
        parallel.sum_across_ranks(flat)
        if rank == 0:
            np.save(out, np.concatenate([flat.numpy(), norm.numpy()]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_equals_full_batch(tmp_path):
    from oracle import vhjb_oracle as V
    from tests.helpers_vhjb import problem, sample_batch
    out = str(tmp_path / "rank0.npy")
    tmp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    p = problem("quad2d")
    orc = V.VhjbOracle(p, V.init_weights(p.sys.n, seed=3))
    xs, dones, costs = sample_batch("quad2d", 1000, seed=4)
    total, hjb, term, grads, _ = orc.loss_and_grad(xs, dones, costs, 0.3)
    ref = np.concatenate([g.reshape(-1) for g in grads])
    P = ref.size
    np.testing.assert_allclose(got[:P], ref, rtol=1e-9, atol=1e-12)
    norm = got[P + 2:]
    d = dones.astype(np.float64)
    np.testing.assert_allclose(norm, [(1 - d).sum() + p.eps, d.sum() + p.eps], rtol=1e-12)
    assert abs(got[P] / norm[0] - hjb) < 1e-10 * hjb and abs(got[P + 1] / norm[1] - term) < 1e-10 * term

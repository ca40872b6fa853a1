"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The hot path shards by initial state / sampled state; the ONLY data-path collectives are the two in
``VhjbKernels.train_step``: an all-reduce of the two done-counts (so that every rank normalises by the GLOBAL batch,
controller/vhjb.py:241, 253) and an all-reduce of the flat value-net gradient with the two loss sums appended
(~101 KB).  Rollouts need no collective at all.
"""
from __future__ import annotations

from typing import Tuple


def dist_info() -> Tuple[int, int]:
    """(rank, world_size) of the default process group; (0, 1) when torch.distributed is not initialised."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def is_distributed() -> bool:
    """True when a default process group with more than one rank is initialised."""
    return dist_info()[1] > 1


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``total`` items owned by ``rank`` — sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_counts(local_counts, eps: float, group=None):
    """local [sum(1-done), sum(done)] (no eps) -> global normalisers [sum(1-done) + eps, sum(done) + eps], in place."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(local_counts, group=group)
    local_counts += eps
    return local_counts


def sum_across_ranks(flat, group=None):
    """All-reduce(sum) of a flat buffer (gradient + loss sums), in place; a no-op on a single process."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(flat, group=group)
    return flat

"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The hot path shards by initial state / sampled state; the ONLY data-path collectives are the two in
``VhjbKernels.train_step``: the two done-counts (so that every rank normalises by the GLOBAL batch,
controller/vhjb.py:241, 253) and the flat value-net gradient with the two loss sums appended (~101 KB).  Rollouts need no
collective at all.  On NVLink-connected GPUs the exchange is done by the library's own kernel over peer memory
(``PeerBuffers`` + ``hjb_vhjb_train_step_peer``: reduce, exchange and Adam in one launch); NCCL all-reduces are the fallback
(``HJB_VHJB_EXCHANGE=nccl``, or when symmetric memory cannot be set up).
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


class PeerBuffers:
    """Exchange buffers and flags of ``hjb_vhjb_train_step_peer`` in torch symmetric memory (peer-mapped over NVLink):
    every rank can store into every rank's buffer.  ``bufs`` / ``flags`` are device arrays [world] of the peers' pointers."""

    def __init__(self, n: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from q_learning_with_hjb_b200 import _lib as L
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nf = int(L.lib().hjb_vhjb_peer_exchange_floats(n, self.world))
        ng = int(L.lib().hjb_vhjb_peer_exchange_flags(n, self.world))
        self.mem = symm.empty(nf + ng, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        self.mem.zero_()
        self.handle = symm.rendezvous(self.mem, group if group is not None else dist.group.WORLD)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.bufs = torch.tensor(ptrs, dtype=torch.int64, device="cuda")
        self.flags = torch.tensor([p + 4 * nf for p in ptrs], dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        dist.barrier(group)          # every rank's flags are zero before anyone's first step


def peer_buffers(n: int, group=None):
    """``PeerBuffers`` for a value net of state dimension n, or None (NCCL fallback: HJB_VHJB_EXCHANGE=nccl, a non-NCCL
    process group as in the CPU tests, or symmetric memory unavailable)."""
    import os
    import torch.distributed as dist
    if os.environ.get("HJB_VHJB_EXCHANGE", "") == "nccl" or not is_distributed():
        return None
    import torch
    if dist.get_backend(group) != "nccl":
        return None
    px = None
    try:
        px = PeerBuffers(n, group)
    except Exception as exc:   # noqa: BLE001 — any set-up failure means: use NCCL
        import warnings
        warnings.warn(f"peer-memory exchange unavailable ({exc!r}); falling back to NCCL all-reduces")
    # every rank must take the same path: peer exchange only if it was set up everywhere
    ok = torch.tensor([1.0 if px is not None else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    return px if float(ok.item()) > 0 else None

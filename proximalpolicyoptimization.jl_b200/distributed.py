"""Minibatch-sharded data parallelism (no reference counterpart).

One process per GPU.  The buffer is sharded by contiguous transition ranges cut at EPISODE
boundaries (so the returns scan needs no communication); every rank draws its own local
permutation (seed + rank) and contributes B/G rows to each global minibatch; gradients are
summed with one NCCL all-reduce per minibatch inside ``ppo_step_epoch``.  torch.distributed is
used only to distribute the NCCL unique id (and works with the gloo backend on CPU for tests of
the host-side logic).
"""
from __future__ import annotations

import numpy as np


def shard_bounds_at_episode_ends(terminal, nranks):
    """Cut [0, N) into ``nranks`` contiguous ranges of near-equal length whose right ends fall on
    episode ends (terminal[i] == True is the LAST transition of an episode, reference
    src/collect_rollouts.jl:34).  Returns a list of (start, stop)."""
    t = np.asarray(terminal).astype(bool)
    n = t.size
    ends = np.flatnonzero(t) + 1          # candidate cut positions (exclusive stops)
    if ends.size == 0 or ends[-1] != n:
        ends = np.append(ends, n)
    bounds, start = [], 0
    for r in range(nranks):
        if r == nranks - 1:
            stop = n
        else:
            target = (n * (r + 1)) // nranks
            j = np.searchsorted(ends, target, side="left")
            cands = [ends[k] for k in (j - 1, j) if 0 <= k < ends.size and ends[k] > start]
            stop = min(cands, key=lambda e: abs(int(e) - target)) if cands else start
            stop = int(min(stop, n))
        bounds.append((int(start), int(stop)))
        start = stop
    return bounds


def equalize_counts(bounds):
    """Rows every rank uses so that all ranks run the same number of minibatches: min shard length."""
    return min(b - a for a, b in bounds)


def local_seed(seed, rank):
    return (int(seed) * 1000003 + int(rank)) & (2 ** 64 - 1)


def enable_p2p_gradients(policy, backend_group=None):
    """Exchange the minibatch gradient over NVLink peer memory, fused into the Adam kernel (csrc/dp_p2p.cu), instead of an
    NCCL all-reduce per minibatch: every rank exports a CUDA-IPC handle of its exchange buffer, the handles are
    all-gathered over torch.distributed, and every rank maps its peers' buffers.  Call after ``init_comm``, on every
    rank, before the first update.  Returns False (and leaves the NCCL path in place) in a single-rank job."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return False
    handles = [None] * world
    dist.all_gather_object(handles, policy.p2p_export(), group=backend_group)
    policy.p2p_connect(world, rank, b"".join(handles))
    return True


def init_comm(ctx, backend_group=None):
    """Create the library's NCCL communicator across the ranks of torch.distributed: rank 0 draws
    the unique id, it is broadcast with the process group already initialised by the launcher."""
    import torch
    import torch.distributed as dist
    from .context import unique_id
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    payload = [unique_id() if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0, group=backend_group)
    ctx.comm_init(world, rank, payload[0])

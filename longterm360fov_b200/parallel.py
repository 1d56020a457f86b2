"""Data-parallel plumbing: samples (viewer x window) are independent, so
training shards the batch contiguously by rank and sum-allreduces ONE flat fp32 gradient
bucket per step; inference shards by batch with no collective.  torch.distributed is the
transport (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous [lo, hi) slice of n samples owned by `rank` (remainder to the low ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(arrays, rank, world):
    lo, hi = shard_bounds(len(arrays[0]), rank, world)
    return [a[lo:hi] for a in arrays]


def allreduce_gradients(flat_grads, group=None):
    """Sum the flat gradient bucket over the group; returns the scale (1/world) the optimiser
    step must apply so the update equals the gradient of the GLOBAL-batch mean loss (Keras
    losses are means over the batch; equal shard sizes assumed)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def gather_predictions(local, group=None):
    """Inference: gather per-rank prediction arrays on every rank (reporting only)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    out = [None] * world
    dist.all_gather_object(out, local, group=group)
    return np.concatenate(out, axis=0)

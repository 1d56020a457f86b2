"""Data-parallel plumbing: samples (viewer x window) are independent, so
training shards the batch contiguously by rank and sum-allreduces ONE flat fp32 gradient
bucket per step; inference shards by batch with no collective.

Transports
  * ``FovComm``   - the C ABI's own NCCL communicator (fov_dp_init / fov_dp_allreduce /
                    fov_dp_broadcast / fov_dp_destroy in include/fov360.h): NCCL over NVLink 5 /
                    NVSwitch, asynchronous on the compute stream.  The product path on GPUs.
  * ``TorchComm`` - a torch.distributed process group (gloo in the CPU tests, where no CUDA exists).

Weighting: Keras losses are means over the batch, so the gradient of the GLOBAL-batch mean is
sum_r (n_r / n) g_r with g_r the gradient of rank r's local mean.  Each rank therefore back-propagates
its loss with seed n_r (gradients come out multiplied by the local sample count), writes n_r into the
reserved last element of the bucket, the bucket is summed, and the optimiser kernel divides by the
summed count it finds there (a device scalar: no host synchronisation).  Unequal shards - the remainder
of ``shard_bounds`` or a kept last partial batch - are therefore exact, not approximately averaged.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous [lo, hi) slice of n samples owned by `rank` (remainder to the low ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(arrays, rank, world):
    lo, hi = shard_bounds(len(arrays[0]), rank, world)
    return [a[lo:hi] for a in arrays]


class TorchComm:
    """torch.distributed transport (gloo on CPU, or torch's own NCCL group)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)

    def allreduce_sum(self, flat):
        import torch.distributed as dist
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)

    def broadcast(self, flat, root=0):
        import torch.distributed as dist
        if self.world > 1:
            dist.broadcast(flat, src=root, group=self.group)

    def destroy(self):
        pass


class FovComm:
    """The C ABI's NCCL communicator.  ``bootstrap`` ships rank 0's unique id to every rank: by default a
    torch.distributed object broadcast on the default group (any backend), or pass ``exchange(bytes_or_None) ->
    bytes`` to use your own transport (MPI, a file, a TCP store ...)."""

    def __init__(self, rank, world, exchange=None):
        import torch
        from . import _lib
        lib = _lib.load()
        nbytes = lib.fov_dp_unique_id_bytes()
        buf = (C.c_char * nbytes)()
        if rank == 0:
            _lib.check(lib.fov_dp_get_unique_id(C.cast(buf, C.c_void_p)), "fov_dp_get_unique_id")
        payload = bytes(buf) if rank == 0 else None
        if exchange is None:
            import torch.distributed as dist
            box = [payload]
            dist.broadcast_object_list(box, src=0)
            payload = box[0]
        else:
            payload = exchange(payload)
        buf = (C.c_char * nbytes).from_buffer_copy(payload)
        _lib.check(lib.fov_dp_init(C.cast(buf, C.c_void_p), int(rank), int(world)), "fov_dp_init")
        self._lib, self._check = lib, _lib.check
        self.world, self.rank = int(world), int(rank)
        self._torch = torch

    def _stream(self):
        return self._torch.cuda.current_stream().cuda_stream

    def allreduce_sum(self, flat):
        self._check(self._lib.fov_dp_allreduce(flat.data_ptr(), flat.numel(), self._stream()), "fov_dp_allreduce")

    def broadcast(self, flat, root=0):
        self._check(self._lib.fov_dp_broadcast(flat.data_ptr(), flat.numel(), int(root), self._stream()),
                    "fov_dp_broadcast")

    def destroy(self):
        self._check(self._lib.fov_dp_destroy(), "fov_dp_destroy")


def allreduce_gradients(flat_grads, comm, local_count):
    """Count-weighted gradient allreduce.  ``flat_grads``: the flat bucket whose LAST element is the reserved
    count slot and whose gradients are already multiplied by the rank's sample count (Model._backward scales the
    bucket after BPTT).  After the call the bucket holds the sums and ``flat_grads[-1]`` the global
    sample count the optimiser divides by (ops.adam_step(..., grad_div=flat_grads[-1:]))."""
    flat_grads[-1] = float(local_count)
    comm.allreduce_sum(flat_grads)
    return flat_grads[-1:]


def gather_predictions(local, group=None):
    """Inference: gather per-rank prediction arrays on every rank (reporting only)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    out = [None] * world
    dist.all_gather_object(out, local, group=group)
    return np.concatenate(out, axis=0)

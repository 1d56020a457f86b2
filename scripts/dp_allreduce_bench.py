#!/usr/bin/env python
"""Gradient-bucket allreduce alone: the C ABI's communicator (fov_dp_allreduce) vs torch.distributed's NCCL group.
torchrun --nproc-per-node N scripts/dp_allreduce_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from longterm360fov_b200 import parallel

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl")
fc = parallel.FovComm(rank, world)
tc = parallel.TorchComm() if hasattr(parallel, "TorchComm") else None
for n in (921858 + 64, 15182814):
    buf = torch.ones(n, device="cuda")
    for name, fn in (("fov_dp_allreduce", lambda: fc.allreduce_sum(buf)), ("torch.distributed", lambda: dist.all_reduce(buf))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print("%-18s %9d floats, %d ranks: %.1f us per allreduce" % (name, n, world, e0.elapsed_time(e1) / 50 * 1e3))
        buf.fill_(1.0)
dist.destroy_process_group()

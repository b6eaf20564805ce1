#!/usr/bin/env python
"""tools/gather_probe.py -- development aid: rank-0 ingress of the BEV gather (NCCL send/recv over
NVLink) for one cfg-2 output batch per rank, under whatever NCCL_* environment the caller sets.

    torchrun --nproc-per-node N tools/gather_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bev_b200 import sharding  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world, rank = dist.get_world_size(), dist.get_rank()
out = torch.empty((256, 1024, 1024, 3), dtype=torch.uint8, device=dev)
sharding.gather_to_rank0(out[:2])
for chunks in (1, 4, 8):
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g = sharding.gather_to_rank0(out, chunks=chunks)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
        del g
    if rank == 0:
        nbytes = (world - 1) * out.numel()
        print("N=%d chunks=%d: %.2f ms  %.0f GB/s into rank 0  [%s]" % (
            world, chunks, best, nbytes / best / 1e6,
            " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("NCCL_"))), flush=True)
dist.destroy_process_group()

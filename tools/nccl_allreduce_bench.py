#!/usr/bin/env python
"""All-reduce / reduce-scatter of the projector gradient message (54 512 062 fp32 = 218 MB, and its bf16 form) over NCCL
on the GPUs of one box — to pick NCCL settings for dist.allreduce_gradients.  Launch with torchrun; NCCL_* environment
variables are read when the process group is created, so every setting is its own launch."""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 54512062
N8 = (N + 8 * 128 - 1) // (8 * 128) * (8 * 128)
out = []
for name, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
    x = torch.randn(N8, device=dev, dtype=torch.float32).to(dtype)
    shard = torch.empty(N8 // world, device=dev, dtype=dtype)
    for op in ("all_reduce", "reduce_scatter", "rs+ag"):
        def run():
            if op == "all_reduce":
                dist.all_reduce(x)
            elif op == "reduce_scatter":
                dist.reduce_scatter_tensor(shard, x)
            else:
                dist.reduce_scatter_tensor(shard, x)
                dist.all_gather_into_tensor(x, shard)
        for _ in range(5):
            run()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        nbytes = N8 * x.element_size()
        factor = 2 * (world - 1) / world if op != "reduce_scatter" else (world - 1) / world
        out.append("%s %s %.3f ms busbw %.0f GB/s" % (name, op, ms, nbytes * factor / ms / 1e6))
if rank == 0:
    print(os.environ.get("TAG", ""), "|", " | ".join(out), flush=True)
dist.destroy_process_group()

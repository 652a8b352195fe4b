"""All-reduce probe (torchrun): time a sum all-reduce of the gradient volume of BASELINE config #3
(708 MB FP32) as one call and as the 5-tensor split the backward uses; prints algbw / busbw."""
import os
import sys
import time

import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
world = dist.get_world_size()
P = 3_000_000
sizes = {"flat": [P * 59], "split": [P * 3, P * 3, P * 4, P, P * 48]}
for name, szs in sizes.items():
    ts = [torch.ones(n, device="cuda") for n in szs]
    for _ in range(3):
        for t in ts:
            dist.all_reduce(t)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        for t in ts:
            dist.all_reduce(t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = sum(szs) * 4
    if dist.get_rank() == 0:
        alg = nbytes / ms / 1e6
        print(f"{os.environ.get('NCCL_ALGO', 'default'):8s} {name:6s} {nbytes / 1e6:.0f} MB  {ms:.3f} ms  algbw {alg:.0f} GB/s  "
              f"busbw {alg * 2 * (world - 1) / world:.0f} GB/s", flush=True)
dist.destroy_process_group()

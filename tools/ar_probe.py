"""All-reduce probe (torchrun): the gradient volume of BASELINE config #3 (708 MB FP32) summed over the ranks with
  nccl        torch.distributed.all_reduce (NCCL's own choice of algorithm; NCCL_ALGO in the environment overrides)
  multimem    torch.ops.symm_mem.multimem_all_reduce_ on a symmetric-memory buffer (NVSwitch in-switch reduction:
              multimem.ld_reduce of a 1/N slice + multimem.st broadcast)
  two_shot    torch.ops.symm_mem.two_shot_all_reduce_ (peer loads over NVLink, no multicast)
Prints ms, algbw and busbw per variant; a variant that is not available on the box prints why."""
import os
import sys

import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
world, rank = dist.get_world_size(), dist.get_rank()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000 * 59


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms):
    if rank == 0:
        alg = N * 4 / ms / 1e6
        print(f"N={world} {name:10s} {N * 4 / 1e6:.0f} MB  {ms:.3f} ms  algbw {alg:.0f} GB/s  busbw {alg * 2 * (world - 1) / world:.0f} GB/s",
              flush=True)


t = torch.ones(N, device="cuda")
report("nccl" + ("/" + os.environ["NCCL_ALGO"] if "NCCL_ALGO" in os.environ else ""), timed(lambda: dist.all_reduce(t)))
try:
    import torch.distributed._symmetric_memory as symm_mem
    group = dist.group.WORLD
    buf = symm_mem.empty(N, dtype=torch.float32, device=torch.device(f"cuda:{local}"))
    hdl = symm_mem.rendezvous(buf, group.group_name)
    buf.fill_(1.0)
    if rank == 0:
        print(f"symm_mem: multicast_ptr = {hdl.multicast_ptr:#x}, world {hdl.world_size}", flush=True)
    for name, op in (("multimem", "multimem_all_reduce_"), ("two_shot", "two_shot_all_reduce_")):
        try:
            f = getattr(torch.ops.symm_mem, op)
            buf.fill_(1.0)
            f(buf, "sum", group.group_name)
            torch.cuda.synchronize()
            ok = bool((buf[:1000] == world).all()) and bool((buf[-1000:] == world).all())
            ms = timed(lambda: f(buf, "sum", group.group_name))
            report(name + ("" if ok else "(WRONG)"), ms)
        except Exception as e:  # noqa
            if rank == 0:
                print(f"{name}: unavailable: {repr(e)[:300]}", flush=True)
except Exception as e:  # noqa
    if rank == 0:
        print(f"symmetric memory unavailable: {repr(e)[:300]}", flush=True)
dist.barrier()
dist.destroy_process_group()

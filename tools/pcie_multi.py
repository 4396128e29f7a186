"""Raw pinned host<->device copy rates with every rank copying at the same time (is the e2e leg at N > 1 limited by the
host path of the box?).  torchrun --nproc-per-node N tools/pcie_multi.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 28
h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
d_a, d_b = torch.empty(n, dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


both(); torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(4):
    both()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 4
gbs = n * 4 / dt / 1e9
t = torch.tensor([gbs], device=dev)
if world > 1:
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    if rank == 0:
        v = [float(x) for x in out]
        print(f"{world} ranks copying both ways at once: per-rank GB/s each way {['%.1f' % x for x in v]}, total each way {sum(v):.1f}")
    dist.destroy_process_group()
else:
    print(f"1 rank: {gbs:.1f} GB/s each way")

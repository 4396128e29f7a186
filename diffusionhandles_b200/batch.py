"""Batched edit -> warp pipeline: the public call for sweeps of independent (image x transform) edits.

``EditWarpPipeline.run_host`` takes HOST (pinned) buffers - depths, masks, rigid transforms and one
activation stack per edit - and returns the warped stacks and the correspondence counts in HOST buffers.
Edits are processed in chunks on a small ring of CUDA streams so that the host->device copies, the kernels
(K1 -> K2 -> masks -> correspondences -> dense maps -> K3) and the device->host copies of neighbouring
chunks overlap.  Edits are independent, so a multi-GPU sweep is plain sharding (``shard_edits``) with no
collective on the data path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _native as N
from . import warp
from .engine import EditEngine


def shard_edits(n_edits: int, rank: int, world_size: int) -> List[int]:
    """Static round-robin partition of independent edits: edit e -> rank e mod world_size."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, n_edits, world_size))


def bind_host_to_gpu_numa_node(device) -> Optional[dict]:
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off, so that pinned staging buffers
    allocated afterwards are node-local (first touch) and H2D/D2H DMA does not cross the socket interconnect.
    Call before allocating pinned memory.  Returns {'numa_node', 'cpus'} or None when the topology is not visible
    (containers without sysfs NUMA information); never raises."""
    import os
    try:
        props = torch.cuda.get_device_properties(device)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception:
        return None


class EditWarpPipeline:
    def __init__(self, device: torch.device, S: int, level_shapes: Sequence[Tuple[int, int]], chunk: int = 16,
                 n_streams: int = 3, full_winner_map: bool = True):
        """level_shapes: [(channels, side), ...] of the activation stack, e.g. SD2-depth config 2:
        [(320,64),(640,32),(1280,16),(1280,8)]."""
        self.device = torch.device(device)
        self.S, self.chunk, self.level_shapes = S, chunk, list(level_shapes)
        self.full_winner_map = full_winner_map
        self.slots = []
        f32 = torch.float32
        for _ in range(n_streams):
            slot = dict(
                stream=torch.cuda.Stream(device=self.device),
                engine=EditEngine(self.device, chunk, S, S),
                depth=torch.empty((chunk, S, S), dtype=f32, device=self.device),
                bg=torch.empty((chunk, S, S), dtype=f32, device=self.device),
                mask=torch.empty((chunk, S, S), dtype=f32, device=self.device),
                levels=[torch.empty((chunk, c, s, s), dtype=f32, device=self.device) for c, s in self.level_shapes],
                outs=[torch.empty((chunk, c, s, s), dtype=f32, device=self.device) for c, s in self.level_shapes],
                done=torch.cuda.Event(),
            )
            self.slots.append(slot)
        self.launches_per_chunk = 4 + 2 + 2 + 4 + 2 + 1 + 2 * len(self.level_shapes) + 1

    def h2d_bytes_per_edit(self) -> int:
        return 3 * self.S * self.S * 4 + sum(c * s * s * 4 for c, s in self.level_shapes)

    def d2h_bytes_per_edit(self) -> int:
        return sum(c * s * s * 4 for c, s in self.level_shapes) + 4

    def run_host(self, depth_h: torch.Tensor, bg_h: torch.Tensor, mask_h: torch.Tensor, intrinsics: torch.Tensor,
                 rigids: Sequence[N.dh_rigid], levels_h: Sequence[torch.Tensor], outs_h: Sequence[torch.Tensor],
                 n_corr_h: torch.Tensor) -> None:
        """All *_h tensors are pinned host tensors with a leading edit dimension E (E % chunk == 0).
        On return (after a final synchronisation) outs_h / n_corr_h hold the results."""
        E = depth_h.shape[0]
        if E % self.chunk:
            raise ValueError(f"number of edits {E} must be a multiple of the chunk size {self.chunk}")
        sides = [s for _, s in self.level_shapes]
        for ci, e0 in enumerate(range(0, E, self.chunk)):
            slot = self.slots[ci % len(self.slots)]
            e1 = e0 + self.chunk
            with torch.cuda.stream(slot["stream"]):
                slot["depth"].copy_(depth_h[e0:e1], non_blocking=True)
                slot["bg"].copy_(bg_h[e0:e1], non_blocking=True)
                slot["mask"].copy_(mask_h[e0:e1], non_blocking=True)
                for d, h in zip(slot["levels"], levels_h):
                    d.copy_(h[e0:e1], non_blocking=True)
                eng: EditEngine = slot["engine"]
                res = eng.run(slot["depth"], slot["bg"], slot["mask"], intrinsics, list(rigids[e0:e1]), poisson=False,
                              sync_counts=False)
                maps = warp.dense_source_maps(res.corr, res.n_corr, self.S, sides,
                                              res.winner_src if self.full_winner_map else None)
                warp.warp_stacks(slot["levels"], maps, slot["outs"])
                for d, h in zip(slot["outs"], outs_h):
                    h[e0:e1].copy_(d, non_blocking=True)
                n_corr_h[e0:e1].copy_(res.n_corr, non_blocking=True)
        for slot in self.slots:
            slot["stream"].synchronize()


def gather_records(records: torch.Tensor, dst: int = 0):
    """One collective after a sharded sweep: gathers a fixed-size per-edit record tensor (n_local, k) from every
    rank to ``dst`` and re-interleaves the rows into global edit order (edit e lives on rank e mod world).
    Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = [torch.empty_like(records) for _ in range(world)] if rank == dst else None
    dist.gather(records, buf, dst=dst)
    if rank != dst:
        return None
    n_local = records.shape[0]
    out = torch.empty((n_local * world,) + tuple(records.shape[1:]), dtype=records.dtype, device=records.device)
    for r in range(world):
        out[r::world] = buf[r]
    return out

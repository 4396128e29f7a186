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
        self._table = None
        self._idx_pinned = None

    def h2d_bytes_per_edit(self, edits_per_scene: float = 1.0) -> float:
        """depth + bg + mask (shared by the `edits_per_scene` edits of a scene when a scene table is used) + the stack."""
        return 3 * self.S * self.S * 4 / edits_per_scene + sum(c * s * s * 4 for c, s in self.level_shapes)

    def d2h_bytes_per_edit(self) -> int:
        return sum(c * s * s * 4 for c, s in self.level_shapes) + 4

    def run_host(self, depth_h: torch.Tensor, bg_h: torch.Tensor, mask_h: torch.Tensor, intrinsics: torch.Tensor,
                 rigids: Sequence[N.dh_rigid], levels_h: Sequence[torch.Tensor], outs_h: Sequence[torch.Tensor],
                 n_corr_h: torch.Tensor, scene_index: Optional[Sequence[int]] = None) -> None:
        """All *_h tensors are pinned host tensors.  levels_h / outs_h / n_corr_h have a leading edit dimension E
        (E % chunk == 0).  Without ``scene_index`` depth_h / bg_h / mask_h are per EDIT (E,S,S).  With ``scene_index``
        (E ints) they are a SCENE TABLE (n_scenes,S,S): an edit sweep is typically a few scenes x many transforms, so every
        depth / background / mask crosses PCIe once per scene instead of once per edit and the edits pick their scene on the
        device.  On return (after a final synchronisation) outs_h / n_corr_h hold the results."""
        E = levels_h[0].shape[0]
        if E % self.chunk:
            raise ValueError(f"number of edits {E} must be a multiple of the chunk size {self.chunk}")
        sides = [s for _, s in self.level_shapes]
        table = None
        if scene_index is not None:
            if len(scene_index) != E:
                raise ValueError("scene_index needs one entry per edit")
            n_scenes = depth_h.shape[0]
            if self._table is None or self._table[0].shape[0] < n_scenes:
                self._table = tuple(torch.empty((n_scenes, self.S, self.S), dtype=torch.float32, device=self.device) for _ in range(3))
                self._table_ready = torch.cuda.Event()
            up = self.slots[0]["stream"]
            with torch.cuda.stream(up):
                for d, h in zip(self._table, (depth_h, bg_h, mask_h)):
                    d[:n_scenes].copy_(h, non_blocking=True)
                self._table_ready.record(up)
            table = self._table
            if self._idx_pinned is None or self._idx_pinned.numel() < E:       # (pinned allocations are slow: keep the buffer)
                self._idx_pinned = torch.empty(E, dtype=torch.int64).pin_memory()
            idx_all = self._idx_pinned[:E]
            idx_all.copy_(torch.as_tensor(list(scene_index), dtype=torch.int64))
            if int(idx_all.min()) < 0 or int(idx_all.max()) >= n_scenes:
                raise IndexError(f"scene_index outside the scene table of {n_scenes} scenes")
        for ci, e0 in enumerate(range(0, E, self.chunk)):
            slot = self.slots[ci % len(self.slots)]
            e1 = e0 + self.chunk
            with torch.cuda.stream(slot["stream"]):
                if table is None:
                    slot["depth"].copy_(depth_h[e0:e1], non_blocking=True)
                    slot["bg"].copy_(bg_h[e0:e1], non_blocking=True)
                    slot["mask"].copy_(mask_h[e0:e1], non_blocking=True)
                else:
                    slot["stream"].wait_event(self._table_ready)
                    idx = idx_all[e0:e1].to(self.device, non_blocking=True)
                    torch.index_select(table[0], 0, idx, out=slot["depth"])
                    torch.index_select(table[1], 0, idx, out=slot["bg"])
                    torch.index_select(table[2], 0, idx, out=slot["mask"])
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


class DeviceSweep:
    """A sweep of independent edits whose inputs (depths, masks, transforms, activation stacks) are RESIDENT on one GPU:
    K1 -> K2 -> masks -> correspondences -> dense maps -> K3 per chunk, all on the current stream.  The launch chain of a
    chunk (~25 launches) is captured once in a CUDA graph and replayed, so that a small shard (a strong-scaling run leaves 32
    edits per GPU at 8 GPUs) is not bound by host launch overhead.  This is the per-rank body of BASELINE config 4."""

    def __init__(self, device: torch.device, S: int, level_shapes: Sequence[Tuple[int, int]], depth: torch.Tensor, bg: torch.Tensor,
                 mask: torch.Tensor, intrinsics: torch.Tensor, rigids: Sequence[N.dh_rigid], levels: Sequence[torch.Tensor],
                 chunk: Optional[int] = None, use_graph: bool = True, full_winner_map: bool = True, branches: int = 1):
        """``branches`` > 1 (graphs only): a chunk is cut into that many sub-chunks whose launch chains are captured on forked
        streams, i.e. as parallel branches of the graph.  At small chunks every kernel of the chain is latency bound (the exact
        sequential centroid alone is ~0.17 ms whatever the batch), so independent sub-chains overlap almost perfectly."""
        self.device = torch.device(device)
        E = depth.shape[0]
        self.E, self.S = E, S
        self.chunk = chunk or E
        if E % self.chunk:
            raise ValueError(f"number of edits {E} must be a multiple of the chunk size {self.chunk}")
        self.branches = branches if use_graph else 1
        if self.chunk % self.branches:
            raise ValueError(f"chunk size {self.chunk} must be a multiple of the number of branches {self.branches}")
        self.sub = self.chunk // self.branches
        self.sides = [s for _, s in level_shapes]
        self.depth, self.bg, self.mask, self.K, self.rigids, self.levels = depth, bg, mask, intrinsics, list(rigids), list(levels)
        self.outs = [torch.empty_like(l) for l in levels]
        self.n_corr = torch.zeros(E, dtype=torch.int32, device=self.device)
        self.rec = torch.zeros((E, 2), dtype=torch.int64, device=self.device)     # per-edit result records (filled inside the chain)
        self.full_winner_map = full_winner_map
        # one engine (scratch) per sub-chunk position when graphs are used: a captured chain is bound to its buffers
        n_chunks = E // self.chunk
        self.engines = [EditEngine(self.device, self.sub, S, S) for _ in range(n_chunks * self.branches if use_graph else 1)]
        self.graphs: List[Optional[torch.cuda.CUDAGraph]] = [None] * n_chunks
        self.use_graph = use_graph
        self.launches_per_chunk = 25 * self.branches

    def _part(self, ci: int, bi: int):
        e0 = ci * self.chunk + bi * self.sub
        e1 = e0 + self.sub
        eng = self.engines[ci * self.branches + bi if self.use_graph else 0]
        res = eng.run(self.depth[e0:e1], self.bg[e0:e1], self.mask[e0:e1], self.K, self.rigids[e0:e1], poisson=False,
                      sync_counts=False)
        maps = warp.dense_source_maps(res.corr, res.n_corr, self.S, self.sides, res.winner_src if self.full_winner_map else None)
        warp.warp_stacks([l[e0:e1] for l in self.levels], maps, [o[e0:e1] for o in self.outs])
        self.n_corr[e0:e1].copy_(res.n_corr)
        # fixed-size result record of every edit: (n_corr, 64-bit checksum of the smallest warped level) - part of the captured chain
        self.rec[e0:e1, 0].copy_(res.n_corr)
        torch.sum(self.outs[-1][e0:e1].flatten(1).view(torch.int32), 1, dtype=torch.int64, out=self.rec[e0:e1, 1])

    def _chunk(self, ci: int, forks=None):
        if forks is None or self.branches == 1:
            for bi in range(self.branches):
                self._part(ci, bi)
            return
        main = torch.cuda.current_stream(self.device)
        for bi, st in enumerate(forks):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                self._part(ci, bi)
        for st in forks:
            main.wait_stream(st)

    def capture(self):
        """Warm up every chunk once, then capture it (idempotent)."""
        if not self.use_graph or all(g is not None for g in self.graphs):
            return
        side = torch.cuda.Stream(device=self.device)
        forks = [torch.cuda.Stream(device=self.device) for _ in range(self.branches)] if self.branches > 1 else None
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for ci in range(len(self.graphs)):
                self._chunk(ci)
            side.synchronize()
            for ci in range(len(self.graphs)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    self._chunk(ci, forks)
                self.graphs[ci] = g
        torch.cuda.current_stream(self.device).wait_stream(side)

    def run(self):
        """Enqueue the whole sweep on the current stream (no synchronisation)."""
        if self.use_graph:
            self.capture()
            for g in self.graphs:
                g.replay()
        else:
            for ci in range(len(self.graphs)):
                self._chunk(ci)

    def records(self) -> torch.Tensor:
        """Fixed-size per-edit result records (n_corr, 64-bit checksum of the first warped level) for the result gather."""
        return self.rec


def gather_records(records: torch.Tensor, dst: int = 0):
    """One collective after a sharded sweep: gathers a fixed-size per-edit record tensor (n_local, k) from every
    rank to ``dst`` and re-interleaves the rows into global edit order (edit e lives on rank e mod world).
    Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = torch.empty((world,) + tuple(records.shape), dtype=records.dtype, device=records.device) if rank == dst else None
    dist.gather(records, list(buf.unbind(0)) if rank == dst else None, dst=dst)
    if rank != dst:
        return None
    # (world, n_local, k) -> (n_local, world, k) -> (n_local * world, k): row e = edit e, one copy
    return buf.transpose(0, 1).reshape((records.shape[0] * world,) + tuple(records.shape[1:]))

"""The activation warp: K3 behind a small Python API.

* ``gather_list(A, y, x)``       W[c,n] = A[c, y[n], x[n]]  - the gather the reference's losses perform
  implicitly (losses.py:46-47, :80);
* ``dense_source_maps(...)``     per-level int32 source maps derived from an edit's splat result;
* ``warp_stacks(levels, maps)``  D[e,c,q] = A[e,c,map[e,q]] (0 where map < 0) for every level of a batch
  of activation stacks - one persistent TMA-staged kernel launch for the whole batch.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import ctypes as C

import torch

from . import _native as N


def gather_list(A: torch.Tensor, y, x) -> torch.Tensor:
    """A (C,h,w) fp32 on CUDA, y/x integer index arrays (tensor / ndarray / list) -> (C,N) fp32."""
    if A.dim() != 3:
        raise ValueError("expected a (C,h,w) activation tensor")
    lib = N.load()
    A = A.contiguous()
    C_, h, w = A.shape
    dev = A.device
    yy = torch.as_tensor(y, device=dev).to(torch.int64)
    xx = torch.as_tensor(x, device=dev).to(torch.int64)
    idx = (yy * w + xx).to(torch.int32).contiguous()
    n = idx.numel()
    out = torch.empty((C_, n), dtype=torch.float32, device=dev)
    if n == 0:                                   # an empty index list gathers nothing (A[:, y, x] -> (C, 0))
        if not A.is_cuda:
            raise N.NativeLibraryError("gather_list runs on CUDA only; there is no CPU fallback")
        return out
    N.check(lib.dh_warp_gather_list(N.ptr(A, torch.float32, "A"), C_, h * w, N.ptr(idx), n, N.ptr(out),
                                    N.stream_handle(dev)), "dh_warp_gather_list")
    return out


def dense_source_maps(corr: torch.Tensor, n_corr: torch.Tensor, img_res: int, sides: Sequence[int],
                      winner_src: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """corr (B,cap,4) int64 device, n_corr (B,) int32 device -> one (B, side*side) int32 map per level.

    ``src(q)`` = source cell of the lowest-n correspondence whose destination cell is q; with ``winner_src``
    (B,img_res*img_res int32: source pixel of each target pixel's splat winner) cells without a correspondence
    fall back to the first target pixel of the cell that has a winner (full winner-index map)."""
    lib = N.load()
    dev = corr.device
    B, cap = corr.shape[0], corr.shape[1]
    sides = [int(v) for v in sides]
    if not 1 <= len(sides) <= 8:
        raise ValueError("between 1 and 8 levels are supported")
    # one buffer, level-major; one scatter + one finalize launch for the whole stack
    buf = torch.empty(B * sum(v * v for v in sides), dtype=torch.int32, device=dev)
    arr = (C.c_int * len(sides))(*sides)
    N.check(lib.dh_dense_source_maps(N.ptr(corr, torch.int64, "corr"), N.ptr(n_corr, torch.int32, "n_corr"), cap,
                                     N.ptr(winner_src, torch.int32, "winner_src") if winner_src is not None else None,
                                     B, img_res, arr, len(sides), N.ptr(buf), N.stream_handle(dev)), "dh_dense_source_maps")
    maps, o = [], 0
    for v in sides:
        maps.append(buf[o:o + B * v * v].view(B, v * v))
        o += B * v * v
    return maps


def warp_stacks(levels: Sequence[torch.Tensor], maps: Sequence[torch.Tensor],
                out: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
    """levels[l]: (B,C_l,h_l,w_l) fp32; maps[l]: (B,h_l*w_l) int32.  Returns the warped levels (same shapes)."""
    if len(levels) != len(maps):
        raise ValueError("one source map per level is required")
    lib = N.load()
    dev = levels[0].device
    B = levels[0].shape[0]
    if out is None:
        out = [torch.empty_like(l) for l in levels]
    arr = (N.dh_warp_level * len(levels))()
    for i, (a, m, o) in enumerate(zip(levels, maps, out)):
        if a.dim() != 4 or a.shape[0] != B:
            raise ValueError("levels must be (B,C,h,w) with a common batch size")
        hw = a.shape[2] * a.shape[3]
        if tuple(m.shape) != (B, hw):
            raise ValueError(f"map {i} must have shape {(B, hw)}, got {tuple(m.shape)}")
        if tuple(o.shape) != tuple(a.shape):
            raise ValueError("output shape mismatch")
        arr[i].inp = N.ptr(a, torch.float32, f"levels[{i}]")
        arr[i].out = N.ptr(o, torch.float32, f"out[{i}]")
        arr[i].src_map = N.ptr(m, torch.int32, f"maps[{i}]")
        arr[i].channels = a.shape[1]
        arr[i].hw = hw
    N.check(lib.dh_warp_gather_dense(arr, len(levels), B, N.stream_handle(dev)), "dh_warp_gather_dense")
    return list(out)

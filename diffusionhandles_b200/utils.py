"""Mirror of the reference's ``diffhandles/utils.py`` for the pieces on the hot path.

``pack_correspondences`` / ``unpack_correspondences`` (utils.py:111-117) define the wire format of the
path: an (N,4) int64 tensor with columns [x_src, y_src, x_dst, y_dst] in full-resolution pixels.
"""
from __future__ import annotations

import torch


def pack_correspondences(original_x, original_y, transformed_x, transformed_y):
    return torch.stack((original_x, original_y, transformed_x, transformed_y), dim=-1)


def unpack_correspondences(correspondences):
    original_x, original_y, transformed_x, transformed_y = torch.split(correspondences, 1, dim=-1)
    return original_x, original_y, transformed_x, transformed_y


def solve_laplacian_depth(fg_depth, bg_depth, mask):
    """utils.py:49-102 - replace the masked region of ``fg_depth`` by the solution of the Poisson problem whose
    Laplacian is that of ``bg_depth`` and whose boundary values are the surrounding ``fg_depth``.
    NumPy in / NumPy out like the reference; solved on the current CUDA device (fp64 CG, tolerance parity)."""
    import numpy as np
    from .depth_transform import _pack_mask, _poisson_device
    dev = torch.device("cuda", torch.cuda.current_device())
    fg = np.asarray(fg_depth)
    img = torch.as_tensor(fg, dtype=torch.float32, device=dev)[None].contiguous()
    src = torch.as_tensor(np.asarray(bg_depth), dtype=torch.float32, device=dev)[None].contiguous()
    m = torch.as_tensor(np.asarray(mask) != 0, device=dev).to(torch.float32)[None].contiguous()
    out = _poisson_device(img, _pack_mask(m), lap_source=src, check=True, what="solve_laplacian_depth")
    return out[0].cpu().numpy().astype(fg.dtype)

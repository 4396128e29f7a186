"""Synthetic inputs of the benchmark / smoke workloads (SURVEY.md 8(d) config-1 recipe, scaled with S):
bg_depth = 4 + row/S + 0.05*rand, foreground = 2 + 0.3*rand inside a disc.  Seeded torch CPU generator, so the
same arrays are produced on every box.  (Input generation only - no part of the hot path.)"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch


def synthetic_scene(S: int = 512, seed: int = 0, cx: Optional[float] = None, cy: Optional[float] = None,
                    radius: Optional[float] = None, quantize: Optional[float] = None):
    g = torch.Generator().manual_seed(seed)
    rows = torch.arange(S, dtype=torch.float32)[:, None]
    bg = 4 + rows / S + 0.05 * torch.rand(S, S, generator=g)
    fgv = 2 + 0.3 * torch.rand(S, S, generator=g)
    k = S / 512.0
    cx = 256 * k if cx is None else cx
    cy = 280 * k if cy is None else cy
    radius = 120 * k if radius is None else radius
    col = torch.arange(S, dtype=torch.float32)[None, :]
    mask = ((col - cx) ** 2 + (rows - cy) ** 2) < radius ** 2
    depth = torch.where(mask, fgv, bg)
    if quantize:
        depth = torch.round(depth / quantize) * quantize
        bg = torch.round(bg / quantize) * quantize
    return depth.numpy().astype(np.float32), bg.numpy().astype(np.float32), mask.numpy().astype(np.float32)

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

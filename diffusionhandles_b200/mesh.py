"""Minimal mesh container: the argument type of ``Renderer.update_scene`` (reference ``diffhandles/mesh.py``;
only ``verts``, ``faces`` and named per-vertex attributes are used on the hot path, depth_transform.py:63-69)."""
from __future__ import annotations

from typing import Dict

import torch


class Mesh:
    def __init__(self, verts: torch.Tensor, faces: torch.Tensor, device=None):
        self.device = verts.device if device is None else torch.device(device)
        self.verts = verts.to(self.device)
        self.faces = faces.to(self.device) if faces is not None else None
        self.vert_attributes: Dict[str, torch.Tensor] = {}

    def add_vert_attribute(self, name: str, values: torch.Tensor, faces: torch.Tensor = None):
        if values.shape[0] != self.verts.shape[0]:
            raise RuntimeError("Number of attribute values does not match the number of vertices.")
        self.vert_attributes[name] = values.to(self.device)

    def has_vert_attribute(self, name: str) -> bool:
        return name in self.vert_attributes

    def to(self, device):
        m = Mesh(self.verts.to(device), None if self.faces is None else self.faces.to(device))
        m.vert_attributes = {k: v.to(device) for k, v in self.vert_attributes.items()}
        return m

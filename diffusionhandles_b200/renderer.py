"""Renderer interface of the reference (``diffhandles/renderer.py:9-60``) and the splat renderer behind it.

``Camera`` / ``RendererArgs`` / ``Renderer`` keep the reference's names and methods.  ``SplatRenderer`` renders a
scene of meshes by z-buffer splatting their VERTICES with K2 (bit-exact (z, index) tie-break, meshes earlier in
the list win ties - like ``join_meshes_as_scene([bg, fg])``) and exposes the two layers the path uses
(depth_transform.py:152): ``world_position`` and ``flat_vertex_color`` as (B,H,W,4) tensors, alpha in channel 3.
``MeshRenderer`` (= ``PyTorch3DRenderer``) is the hard z-buffer TRIANGLE rasteriser of mesh mode (``dh_raster.cu``), written to
pytorch3d's published ``rasterize_meshes`` semantics; pytorch3d itself is not available here, so that parity is unpinned
(SURVEY.md 8(c), 8(f) rank 2).
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from . import _native as N
from .mesh import Mesh


class Camera:
    def __init__(self, intrinsics, extrinsics_R=None, extrinsics_t=None):
        self.intrinsics = intrinsics
        self.extrinsics_R = extrinsics_R if extrinsics_R is not None else torch.eye(3, device=intrinsics.device, dtype=intrinsics.dtype)
        self.extrinsics_t = extrinsics_t if extrinsics_t is not None else torch.zeros(3, device=intrinsics.device, dtype=intrinsics.dtype)


@dataclass
class RendererArgs:
    device: Union[str, torch.device] = "cpu"


class Renderer(torch.nn.Module, ABC):
    def __init__(self, args: RendererArgs = RendererArgs()):
        super().__init__()
        self.device = torch.device(args.device)
        self.args = args

    @abstractmethod
    def update_scene(self, scene_elements: Dict[str, Any], ignore_unsupported_elements: bool = False):
        """Update the renderer's scene representation with new/updated elements."""

    @abstractmethod
    def set_output_layers(self, output_names: List[str]):
        """Define which layers to output."""

    @abstractmethod
    def render(self) -> Dict[str, torch.Tensor]:
        """Render the scene from all cameras; returns a dictionary of named (B,H,W,C) layers."""


@dataclass
class SplatRendererArgs(RendererArgs):
    """Field names follow ``PyTorch3DRendererArgs`` (pytorch3d_renderer.py:31-53); the triangle-only settings are
    accepted for call compatibility and ignored by the point splat."""
    output_res: Union[int, Tuple[int, int]] = 512
    blur_radius: float = 0
    faces_per_pixel: int = 1
    perspective_correct: bool = True
    cull_backfaces: bool = False
    clip_barycentric_coords: bool = True
    bin_size: Optional[int] = None
    blend_type: str = "hard"
    blend_sigma: float = 1e-4
    blend_gamma: float = 1e-4
    background_color: Tuple[float, float, float] = (0.0, 0.0, 0.0)


def fov_scales(k11: torch.Tensor, H: int, W: int, znear: float = 0.1):
    """NDC scales of the FoVPerspectiveCameras the reference builds (pytorch3d_renderer.py:890-913): fov from the
    intrinsics (of the larger image side), aspect ratio 1, evaluated with torch fp32 ops like the reference."""
    k = k11.detach().to("cpu", torch.float32)
    if H != W:
        k = k * (H / max(H, W))
    fov = 2 * torch.rad2deg(torch.atan(1 / k))
    fov = (3.141592653589793 / 180) * fov
    tan_half = torch.tan(fov / 2)
    max_y = tan_half * znear
    min_y = -max_y
    max_x = max_y * 1.0
    min_x = -max_x
    return float(2.0 * znear / (max_x - min_x)), float(2.0 * znear / (max_y - min_y))


class MeshRenderer(Renderer):
    """Hard z-buffer TRIANGLE rasteriser behind the reference's renderer interface (mesh mode).  Follows the published
    semantics of pytorch3d's MeshRasterizer for faces_per_pixel = 1; pytorch3d itself is not
    available, so parity with it is unpinned.  Layers: world_position, camera_position, flat_vertex_color, depth."""
    SUPPORTED_OUTPUTS = ("world_position", "camera_position", "flat_vertex_color", "depth")
    SUPPORTED_SCENE_ELEMENTS = ("meshes", "cameras")
    z_near = 0.1

    def __init__(self, output_names: Optional[List[str]] = None, args: "SplatRendererArgs" = None):
        args = SplatRendererArgs() if args is None else args
        super().__init__(args=args)
        self._scene: Dict[str, Any] = {}
        self.fragments = None
        self.set_output_layers(["world_position"] if output_names is None else output_names)

    def update_scene(self, scene_elements: Dict[str, Any], ignore_unsupported_elements: bool = False):
        if not ignore_unsupported_elements:
            unsupported = set(scene_elements.keys()) - set(self.SUPPORTED_SCENE_ELEMENTS)
            if len(unsupported) > 0:
                raise RuntimeError(f"Unsupported scene elements for the mesh renderer: {unsupported}")
        if "meshes" in scene_elements:
            meshes = scene_elements["meshes"]
            if not isinstance(meshes, list) or not all(isinstance(m, Mesh) for m in meshes):
                raise RuntimeError("Provided meshes must be given as list of geometry.mesh.Mesh")
        self._scene.update({k: v for k, v in scene_elements.items() if k in self.SUPPORTED_SCENE_ELEMENTS})

    def set_output_layers(self, output_names: List[str]):
        unsupported = set(output_names) - set(self.SUPPORTED_OUTPUTS)
        if len(unsupported) > 0:
            raise RuntimeError(f"Unsupported output channels: {unsupported}.")
        if self.args.faces_per_pixel != 1 or self.args.blend_type != "hard":
            raise RuntimeError("Only faces_per_pixel = 1 with hard blending is implemented")
        self.output_names = list(output_names)

    def render(self) -> Dict[str, torch.Tensor]:
        if "meshes" not in self._scene or "cameras" not in self._scene:
            raise RuntimeError("The scene needs 'meshes' and 'cameras' before rendering.")
        lib = N.load()
        res = self.args.output_res
        H, W = (res, res) if isinstance(res, int) else (int(res[0]), int(res[1]))
        meshes: List[Mesh] = self._scene["meshes"]
        dev = meshes[0].verts.device
        verts = torch.cat([m.verts.detach().to(torch.float32) for m in meshes], dim=0).contiguous()
        offs, faces = 0, []
        for m in meshes:        # join_meshes_as_scene: faces of later meshes are offset; earlier meshes win exact z ties
            faces.append(m.faces.to(torch.int64) + offs)
            offs += m.verts.shape[0]
        faces = torch.cat(faces, dim=0).to(torch.int32).contiguous()
        V, F, P = verts.shape[0], faces.shape[0], H * W
        colors = None
        if "flat_vertex_color" in self.output_names:
            dims = {m.vert_attributes["color"].shape[-1] for m in meshes if m.has_vert_attribute("color")}
            if len(dims) != 1:
                raise RuntimeError("Vertex attribute color has different dimensions across meshes.")
            d = dims.pop()
            colors = torch.cat([m.vert_attributes["color"].detach().to(torch.float32) if m.has_vert_attribute("color")
                                else torch.zeros((m.verts.shape[0], d), device=dev) for m in meshes], dim=0).contiguous()
        st = N.stream_handle(dev)
        ws_bytes = int(lib.dh_raster_workspace_bytes(V, H, W))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        layers = {name: [] for name in self.output_names}
        frags = []
        for cam in self._scene["cameras"]:
            sx, sy = fov_scales(cam.intrinsics[1, 1], H, W, self.z_near)
            R = (C.c_float * 9)(*cam.extrinsics_R.detach().to("cpu", torch.float32).reshape(-1).tolist())
            T = (C.c_float * 3)(*cam.extrinsics_t.detach().to("cpu", torch.float32).reshape(-1).tolist())
            p2f = torch.empty((H, W), dtype=torch.int32, device=dev)
            zbuf = torch.empty((H, W), dtype=torch.float32, device=dev)
            bary = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
            N.check(lib.dh_rasterize_meshes(N.ptr(verts), V, N.ptr(faces), F, H, W, R, T, sx, sy, float(self.args.blur_radius),
                                            int(bool(self.args.cull_backfaces)), int(bool(self.args.perspective_correct)),
                                            int(bool(self.args.clip_barycentric_coords)), N.ptr(p2f), N.ptr(zbuf), N.ptr(bary),
                                            N.ptr(ws), ws_bytes, st), "dh_rasterize_meshes")
            frags.append((p2f, zbuf, bary))

            def interp(attr):
                out = torch.empty((H, W, attr.shape[1] + 1), dtype=torch.float32, device=dev)
                N.check(lib.dh_interpolate_face_attributes(N.ptr(attr), attr.shape[1], N.ptr(faces), N.ptr(p2f), N.ptr(bary), H, W,
                                                           N.ptr(out), st), "dh_interpolate_face_attributes")
                return out
            if "world_position" in layers:
                layers["world_position"].append(interp(verts))
            if "camera_position" in layers or "depth" in layers:
                vcam = (verts @ cam.extrinsics_R.to(dev, torch.float32) + cam.extrinsics_t.to(dev, torch.float32)).contiguous()
                cp = interp(vcam)
                if "camera_position" in layers:
                    layers["camera_position"].append(cp)
                if "depth" in layers:
                    layers["depth"].append(cp[..., [2, 3]].contiguous())
            if "flat_vertex_color" in layers:
                layers["flat_vertex_color"].append(interp(colors))
        self.fragments = frags
        return {k: torch.stack(v, dim=0) for k, v in layers.items()}


class SplatRenderer(Renderer):
    SUPPORTED_OUTPUTS = ("world_position", "flat_vertex_color", "depth")
    SUPPORTED_SCENE_ELEMENTS = ("meshes", "cameras")

    def __init__(self, output_names: Optional[List[str]] = None, args: SplatRendererArgs = SplatRendererArgs()):
        super().__init__(args=args)
        self._scene: Dict[str, Any] = {}
        self.set_output_layers(["world_position"] if output_names is None else output_names)

    def update_scene(self, scene_elements: Dict[str, Any], ignore_unsupported_elements: bool = False):
        if not ignore_unsupported_elements:
            unsupported = set(scene_elements.keys()) - set(self.SUPPORTED_SCENE_ELEMENTS)
            if len(unsupported) > 0:
                raise RuntimeError(f"Unsupported scene elements for the splat renderer: {unsupported}")
        if "meshes" in scene_elements:
            meshes = scene_elements["meshes"]
            if not isinstance(meshes, list) or not all(isinstance(m, Mesh) for m in meshes):
                raise RuntimeError("Provided meshes must be given as list of geometry.mesh.Mesh")
        self._scene.update({k: v for k, v in scene_elements.items() if k in self.SUPPORTED_SCENE_ELEMENTS})

    def set_output_layers(self, output_names: List[str]):
        unsupported = set(output_names) - set(self.SUPPORTED_OUTPUTS)
        if len(unsupported) > 0:
            raise RuntimeError(f"Unsupported output channels: {unsupported}.")
        self.output_names = list(output_names)

    def render(self) -> Dict[str, torch.Tensor]:
        if "meshes" not in self._scene or "cameras" not in self._scene:
            raise RuntimeError("The scene needs 'meshes' and 'cameras' before rendering.")
        lib = N.load()
        res = self.args.output_res
        H, W = (res, res) if isinstance(res, int) else (int(res[0]), int(res[1]))
        meshes: List[Mesh] = self._scene["meshes"]
        dev = meshes[0].verts.device
        verts = torch.cat([m.verts.detach().to(torch.float32) for m in meshes], dim=0)
        colors = None
        if "flat_vertex_color" in self.output_names:
            if not all(m.has_vert_attribute("color") for m in meshes):
                raise RuntimeError("flat_vertex_color needs a 'color' vertex attribute on every mesh")
            colors = torch.cat([m.vert_attributes["color"].detach().to(torch.float32) for m in meshes], dim=0)
        n, P = verts.shape[0], H * W
        st = N.stream_handle(dev)
        layers = {name: [] for name in self.output_names}
        for cam in self._scene["cameras"]:
            # world -> camera: R p + t (inverse of depth_transform.py:639); identity for every reference caller
            p = verts @ cam.extrinsics_R.to(dev, torch.float32).T + cam.extrinsics_t.to(dev, torch.float32)
            pts = p.to(torch.float64).contiguous()
            camera = N.make_camera(cam.intrinsics)
            pix = torch.empty(n, dtype=torch.int32, device=dev)
            zkey = torch.empty(n, dtype=torch.int64, device=dev)
            zbuf = torch.empty(P, dtype=torch.int64, device=dev)
            winner = torch.empty(P, dtype=torch.int32, device=dev)
            N.check(lib.dh_project_points(N.ptr(pts), n, H, W, C.byref(camera), N.ptr(pix), N.ptr(zkey), None, None, st))
            N.check(lib.dh_splat_zbuffer(N.ptr(pix), N.ptr(zkey), None, n, n, n, 1, P, N.ptr(zbuf), N.ptr(winner), st))
            w = winner.view(H, W).long()
            hit = w >= 0
            idx = w.clamp(min=0)
            alpha = hit.to(torch.float32)[..., None]
            if "world_position" in layers:
                layers["world_position"].append(torch.cat([verts[idx] * alpha, alpha], dim=-1))
            if "depth" in layers:
                layers["depth"].append(torch.cat([p[idx][..., 2:3] * alpha, alpha], dim=-1))
            if "flat_vertex_color" in layers:
                layers["flat_vertex_color"].append(torch.cat([colors[idx] * alpha, alpha], dim=-1))
        return {k: torch.stack(v, dim=0) for k, v in layers.items()}

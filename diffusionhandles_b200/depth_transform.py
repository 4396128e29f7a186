"""Drop-in mirror of the reference's ``diffhandles/depth_transform.py`` (same names, argument meaning and
error behaviour) backed by libdiffhandles_b200 (sm_100a CUDA kernels, C ABI in include/dh_b200.h).

There is no CPU implementation here: tensors must live on a CUDA device, otherwise the call raises
``NativeLibraryError``.  Reference lines are cited per function.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as N
from .engine import get_engine, make_rigid, pixel_grid
from .utils import pack_correspondences  # noqa: F401  (re-exported: the reference's depth_transform imports it too)


def normalize_depth(depth, bounds=None, return_bounds=False):
    """depth_transform.py:15-28: maps each batch element linearly onto [0, 255]; ``bounds`` = (min, max) tensors reuse another
    image's range (use_input_depth_normalization).  Elementwise torch ops in the reference's operation order
    (255 * (x - min), then the division), on whatever device ``depth`` lives."""
    if depth.dim() != 4:
        raise RuntimeError(f"normalize_depth expects a (B,1,H,W) tensor, got {depth.dim()} dimensions")
    if bounds is not None:
        lo, hi = bounds
    else:
        lo, hi = torch.aminmax(depth.flatten(start_dim=1), dim=1)
        lo, hi = lo.reshape(-1, 1, 1, 1), hi.reshape(-1, 1, 1, 1)
    scaled = (depth - lo) * 255 / (hi - lo)
    return (scaled, (lo, hi)) if return_bounds else scaled


def depth_to_mesh(depth: torch.Tensor, intrinsics: torch.Tensor, extrinsics_R: torch.Tensor = None,
                  extrinsics_t: torch.Tensor = None, mask: torch.Tensor = None):
    """depth_transform.py:30-71: a depth map as a triangle mesh.  One vertex per pixel inside ``mask`` (all pixels without
    one), unprojected on the device; every 2x2 block of pixels whose corners are all vertices gives two counter-clockwise
    triangles (bottom-left, top-right, top-left) and (bottom-left, bottom-right, top-right), blocks in raster order; the
    per-vertex ``color`` attribute is (x/(W-1), y/(H-1), 1 if a mask was given else 0), which the renderer carries to the
    target image as the source coordinate of each rendered pixel."""
    from .mesh import Mesh
    H, W = depth.shape[-2], depth.shape[-1]
    dev = depth.device
    inside = torch.ones((H, W), dtype=torch.bool, device=dev) if mask is None else mask.reshape(H, W).to(dev).bool()
    rows, cols = torch.nonzero(inside, as_tuple=True)                       # raster order = vertex order
    world = depth_to_world_coords(depth, intrinsics=intrinsics, extrinsics_R=extrinsics_R, extrinsics_t=extrinsics_t)
    verts = world[rows, cols].contiguous()
    # vertex number of every pixel (-1 outside the mask), then the four corners of every 2x2 block
    number = torch.full((H, W), -1, dtype=torch.int64, device=dev)
    number[rows, cols] = torch.arange(rows.numel(), dtype=torch.int64, device=dev)
    tl, tr, bl, br = number[:-1, :-1], number[:-1, 1:], number[1:, :-1], number[1:, 1:]
    quads = torch.stack([torch.stack([bl, tr, tl], dim=-1), torch.stack([bl, br, tr], dim=-1)], dim=-2)   # (H-1, W-1, 2, 3)
    faces = quads.reshape(-1, 3)
    faces = faces[(faces >= 0).all(dim=-1)].contiguous()
    mesh = Mesh(verts=verts, faces=faces)
    xs, ys = torch.linspace(0, 1, W, device=dev), torch.linspace(0, 1, H, device=dev)
    flag = torch.full((rows.numel(),), 0.0 if mask is None else 1.0, dtype=xs.dtype, device=dev)
    mesh.add_vert_attribute("color", torch.stack([xs[cols], ys[rows], flag], dim=-1))
    return mesh


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise N.NativeLibraryError(
            f"{name} is on {t.device}; diffusionhandles_b200 runs on CUDA (sm_100a) only and has no CPU fallback")


def depth_to_world_coords(depth: torch.Tensor, intrinsics: torch.Tensor, extrinsics_R: torch.Tensor = None,
                          extrinsics_t: torch.Tensor = None):
    """depth_transform.py:589-641.  depth (1,1,H,W) fp32 -> (H,W,3) fp32 points in pytorch3d axes."""
    if depth.shape[0] != 1:
        raise ValueError("Only batch size 1 is supported")
    height, width = depth.shape[-2:]
    if height < 2 or width < 2:
        raise RuntimeError(f'Expected depth to have at least 2 pixels in each dimension, got {height} x {width}.')
    _require_cuda(depth, "depth")
    lib = N.load()
    d = depth.reshape(1, height, width).to(torch.float32).contiguous()
    xs, ys = pixel_grid(height, width, d.device)
    out = torch.empty((height, width, 3), dtype=torch.float32, device=d.device)
    cam = N.make_camera(intrinsics)
    N.check(lib.dh_unproject(N.ptr(d), 1, height, width, C.byref(cam), N.ptr(xs), N.ptr(ys), N.ptr(out),
                             N.stream_handle(d.device)), "dh_unproject")
    if extrinsics_R is not None or extrinsics_t is not None:
        # camera -> world: R^T (p - t) (depth_transform.py:639); general extrinsics are outside the bit-exact contract
        R = torch.eye(3, device=d.device) if extrinsics_R is None else extrinsics_R.to(d.device, torch.float32)
        t = torch.zeros(3, device=d.device) if extrinsics_t is None else extrinsics_t.to(d.device, torch.float32)
        out = (out - t) @ R
    return out


def transform_points(points: torch.Tensor, rot_angle: torch.Tensor = None, rot_axis: torch.Tensor = None,
                     translation: torch.Tensor = None):
    """depth_transform.py:439-459 (torch fp32 variant used by mesh mode and the webapp): Rodrigues rotation
    about the mean of the given points + translation.  Angle in degrees."""
    _require_cuda(points, "points")
    lib = N.load()
    p = points.to(torch.float32).contiguous()
    n = p.shape[0]
    out = torch.empty_like(p)
    ws_bytes = int(lib.dh_transform_points_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=p.device)
    axis = (C.c_float * 3)(*[float(v) for v in rot_axis.detach().cpu().reshape(-1)[:3]])
    tr = (C.c_float * 3)(*[float(v) for v in translation.detach().cpu().reshape(-1)[:3]])
    N.check(lib.dh_transform_points(N.ptr(p), n, float(rot_angle), axis, tr, N.ptr(out), N.ptr(ws), ws_bytes,
                                    N.stream_handle(p.device)), "dh_transform_points")
    return out


def points_to_depth(points: torch.Tensor, intrinsics: torch.Tensor, output_size: Tuple[int, int],
                    extrinsics_R: torch.Tensor = None, extrinsics_t: torch.Tensor = None, point_mask: torch.Tensor = None):
    """depth_transform.py:643-747.  Z-buffered point splat: returns, like the reference,
    ``(depth_map (1,1,H,W) fp32 on points.device, depth_mask (H,W) bool ndarray, u[visible], v[visible] int64
    ndarrays, visible (N,) bool ndarray)``.  Only identity extrinsics are implemented on the device."""
    if extrinsics_R is not None or extrinsics_t is not None:
        raise NotImplementedError("points_to_depth: only identity extrinsics are supported (all reference callers)")
    _require_cuda(points, "points")
    lib = N.load()
    dev = points.device
    H, W = int(output_size[0]), int(output_size[1])
    pts = points.to(torch.float64).contiguous()
    n = pts.shape[0]
    P = H * W
    st = N.stream_handle(dev)
    cam = N.make_camera(intrinsics)
    i32 = torch.int32
    pix = torch.empty(max(n, 1), dtype=i32, device=dev)
    zkey = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    u = torch.empty(max(n, 1), dtype=i32, device=dev)
    v = torch.empty(max(n, 1), dtype=i32, device=dev)
    zbuf = torch.empty(P, dtype=torch.int64, device=dev)
    winner = torch.empty(P, dtype=i32, device=dev)
    depth_map = torch.empty((H, W), dtype=torch.float32, device=dev)
    target = torch.empty((H, W), dtype=torch.uint8, device=dev)
    visible = torch.zeros(max(n, 1), dtype=torch.uint8, device=dev)
    pm = (torch.zeros(max(n, 1), dtype=torch.uint8, device=dev) if point_mask is None
          else (point_mask.to(dev) != 0).to(torch.uint8).contiguous())
    N.check(lib.dh_project_points(N.ptr(pts), n, H, W, C.byref(cam), N.ptr(pix), N.ptr(zkey), N.ptr(u), N.ptr(v), st),
            "dh_project_points")
    N.check(lib.dh_splat_zbuffer(N.ptr(pix), N.ptr(zkey), None, n, n, max(n, 1), 1, P, N.ptr(zbuf), N.ptr(winner), st),
            "dh_splat_zbuffer")
    N.check(lib.dh_splat_resolve(N.ptr(zbuf), N.ptr(winner), 1, H, W, 0, N.ptr(pm), None, max(n, 1), N.ptr(depth_map),
                                 N.ptr(target), None, None, None, st), "dh_splat_resolve")
    N.check(lib.dh_splat_visible(N.ptr(pix), N.ptr(winner), None, n, n, max(n, 1), 1, P, 0, N.ptr(pm), N.ptr(visible), st),
            "dh_splat_visible")
    vis = visible[:n].bool()
    uv = torch.stack([u[:n][vis], v[:n][vis]]).to(torch.int64).cpu().numpy()
    return (depth_map[None, None], target.bool().cpu().numpy(), uv[0], uv[1], vis.cpu().numpy())


def _pack_mask(mask_f32: torch.Tensor) -> torch.Tensor:
    """(B,H,W) float mask (non-zero = set) -> row-padded bit planes (B,H,ceil(W/32)) int32."""
    lib = N.load()
    B, H, W = mask_f32.shape
    bits = torch.empty((B, H, (W + 31) // 32), dtype=torch.int32, device=mask_f32.device)
    N.check(lib.dh_pack_mask_bits(N.ptr(mask_f32, torch.float32, "mask"), B, H, W, N.ptr(bits), N.stream_handle(mask_f32.device)),
            "dh_pack_mask_bits")
    return bits


def _poisson_device(image: torch.Tensor, mask_bits: torch.Tensor, lap_source: Optional[torch.Tensor] = None,
                    check: bool = False, what: str = "Poisson fill") -> torch.Tensor:
    """fp64 CG fill on the device.  ``check=True`` reads the iteration counts back (one small synchronising copy) and warns
    when a system stopped without reaching the tolerance - the reference's direct solver cannot fail that way."""
    lib = N.load()
    B, H, W = image.shape
    out = torch.empty_like(image)
    ws_bytes = int(lib.dh_poisson_workspace_bytes(B, H, W))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=image.device)
    iters = torch.zeros(B, dtype=torch.int32, device=image.device)
    N.check(lib.dh_poisson_fill_source(N.ptr(image, torch.float32, "image"), N.ptr(mask_bits), None,
                                       N.ptr(lap_source, torch.float32, "lap_source") if lap_source is not None else None,
                                       B, H, W, N.ptr(out), 0, 1e-13, N.ptr(iters), N.ptr(ws), ws_bytes, N.stream_handle(image.device)),
            "dh_poisson_fill_source")
    _poisson_device.last_iters = iters
    if check:
        warn_if_not_converged(iters, what)
    return out


def warn_if_not_converged(iters: torch.Tensor, what: str) -> None:
    """iters: the solver's per-system iteration counts (negative = stopped without reaching the tolerance)."""
    it = iters.cpu().numpy()
    bad = it < 0
    if bad.any():
        import warnings
        warnings.warn(f"{what}: the conjugate-gradient solver did not reach its tolerance for {int(bad.sum())} of {len(it)} system(s) "
                      f"(stopped after {int((-it[bad] - 1).max())} iterations); the filled values are approximate", RuntimeWarning)


def poisson_solve(input_image, mask):
    """depth_transform.py:535-587 - masked Poisson fill (fp64 CG on the device; SuperLU in the reference).
    NumPy in / NumPy out like the reference; the image is processed on the current CUDA device."""
    dev = torch.device("cuda", torch.cuda.current_device())
    arr = np.asarray(input_image)
    img = torch.as_tensor(arr, dtype=torch.float32, device=dev)[None].contiguous()
    m = torch.as_tensor(np.asarray(mask) != 0, device=dev).to(torch.float32)[None].contiguous()
    out = _poisson_device(img, _pack_mask(m), check=True, what="poisson_solve")
    return out[0].cpu().numpy().astype(arr.dtype)


def transform_point_cloud(points, axis, angle_degrees, x, y, z, mask):
    """depth_transform.py:461-533 - rotate a point image about the centroid of the masked points and translate it.
    points (S,S,3) fp32 (NumPy or torch), mask (S,S); returns ``(rotated_points (S,S,3) float64 ndarray,
    modified_indices (S*S,) bool ndarray)`` like the reference (which hard-codes S = 512; any S works here)."""
    lib = N.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    pts = torch.as_tensor(np.asarray(points) if not isinstance(points, torch.Tensor) else points).to(device=dev, dtype=torch.float32)
    shape = tuple(pts.shape)
    pts = pts.reshape(-1, 3).contiguous()
    n = pts.shape[0]
    mk = torch.as_tensor(np.asarray(mask) if not isinstance(mask, torch.Tensor) else mask).to(device=dev)
    mk_flat = (mk.reshape(-1) != 0)
    if mk_flat.numel() != n:
        raise ValueError("mask must have one entry per point")
    rg = make_rigid(angle_degrees, axis, [x, y, z])
    rg.t[:] = [float(x), float(y), float(z)]          # the reference adds the Python floats as given
    out = torch.empty((n, 3), dtype=torch.float64, device=dev)
    cen = torch.empty(3, dtype=torch.float32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.dh_transform_point_cloud_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    N.check(lib.dh_transform_point_cloud(N.ptr(pts), N.ptr(mk_flat.to(torch.float32).contiguous()), n, C.byref(rg), N.ptr(out), N.ptr(cen),
                                         N.ptr(cnt), N.ptr(ws), ws_bytes, N.stream_handle(dev)), "dh_transform_point_cloud")
    return out.reshape(shape).cpu().numpy(), mk_flat.cpu().numpy()


def _empty_mask_result(depth, use_input_depth_normalization):
    """No foreground at all (depth_transform.py:203-216): the normalised input disparity and an empty (0,4) int64 list."""
    disparity = 1.0 / depth
    bounds = normalize_depth(disparity, return_bounds=True)[1] if use_input_depth_normalization else None
    return normalize_depth(disparity, bounds=bounds), torch.empty((0, 4), dtype=torch.int64)


def transform_depth_pc(depth: torch.Tensor, bg_depth: torch.Tensor, fg_mask: torch.Tensor, intrinsics: torch.Tensor,
                       rot_angle: float = None, rot_axis: torch.Tensor = None, translation: torch.Tensor = None,
                       use_input_depth_normalization=False, return_device_result: bool = False):
    """depth_transform.py:198-363 - the default ('pc') depth transform.

    Returns ``(edited disparity (1,1,S,S) fp32 on depth.device in [0,255], correspondences (N,4) int64 CPU)``.
    With ``return_device_result=True`` a third value, the device-side ``EditResult`` (winner indices, masks,
    device copy of the correspondences), is appended for the warp / loss kernels.
    """
    if not fg_mask.any():
        res = _empty_mask_result(depth, use_input_depth_normalization)
        return (*res, None) if return_device_result else res
    S = fg_mask.shape[-1]
    if fg_mask.shape[-2] != S:
        raise RuntimeError(f"the 'pc' depth transform needs a square mask, got {fg_mask.shape[-2]} x {S}")
    if depth.shape[0] != 1:
        raise ValueError(f"one image per call (batch size 1), got a batch of {depth.shape[0]}")
    _require_cuda(depth, "depth")
    # defaults of the reference: no rotation about +y, no translation (depth_transform.py:219-224)
    rot_angle = 0.0 if rot_angle is None else rot_angle
    rot_axis = torch.tensor([0.0, 1.0, 0.0]) if rot_axis is None else rot_axis
    translation = torch.zeros(3) if translation is None else translation
    dev = depth.device
    eng = get_engine(dev, 1, S, S)
    f32 = torch.float32
    d = depth.reshape(1, S, S).to(f32).contiguous()
    b = bg_depth.reshape(1, S, S).to(device=dev, dtype=f32).contiguous()
    m = fg_mask.reshape(1, S, S).to(device=dev, dtype=f32).contiguous()
    res = eng.run(d, b, m, intrinsics, [make_rigid(rot_angle, rot_axis, translation)],
                  use_input_depth_normalization=use_input_depth_normalization, poisson=True)
    corr = res.correspondences(0).cpu()
    disparity = res.disparity[0][None, None].clone()
    return (disparity, corr, res) if return_device_result else (disparity, corr)


def transform_depth_mesh(depth: torch.Tensor, bg_depth: torch.Tensor, fg_mask: torch.Tensor, intrinsics: torch.Tensor,
                         rot_angle: float = None, rot_axis: torch.Tensor = None, translation: torch.Tensor = None,
                         use_input_depth_normalization=False):
    """depth_transform.py:91-195 - opt-in mesh mode: triangulate both depth maps, move the foreground mesh, rasterise the
    scene (hard z-buffer, back faces culled, blur 1e-5) and read depth and source coordinates off the render layers.
    Correspondences are enumerated over TARGET pixels in raster order; no mask cleaning and no hole fill (as in the
    reference).  The rasteriser follows pytorch3d's published semantics; bit parity with pytorch3d is unpinned."""
    from .pytorch3d_renderer import PyTorch3DRenderer, PyTorch3DRendererArgs
    from .renderer import Camera
    if not fg_mask.any():
        return _empty_mask_result(depth, use_input_depth_normalization)
    _require_cuda(depth, "depth")
    dev = depth.device
    H, W = depth.shape[-2], depth.shape[-1]
    angle = torch.tensor(0.0 if rot_angle is None else float(rot_angle), dtype=torch.float32)
    axis = torch.tensor([0.0, 1.0, 0.0], device=dev) if rot_axis is None else rot_axis
    shift = torch.zeros(3, device=dev) if translation is None else translation
    background = depth_to_mesh(depth=bg_depth, intrinsics=intrinsics)
    foreground = depth_to_mesh(depth=depth, intrinsics=intrinsics, mask=fg_mask[0, 0] > 0.5)
    foreground.verts = transform_points(points=foreground.verts, rot_angle=angle, rot_axis=axis, translation=shift)
    renderer = PyTorch3DRenderer(output_names=['world_position', 'flat_vertex_color'],
                                 args=PyTorch3DRendererArgs(device=dev, output_res=(H, W), cull_backfaces=True, blur_radius=1e-5))
    renderer.update_scene(scene_elements={'meshes': [background, foreground], 'cameras': [Camera(intrinsics=intrinsics)]})
    layers = renderer.render()
    target_depth = layers['world_position'][None, ..., 2]                   # (1,1,H,W): world z of the visible surface
    carried = layers['flat_vertex_color'][0]                                # (H,W,4): source x/(W-1), y/(H-1), fg flag, alpha
    # one correspondence per target pixel that shows the foreground mesh, in target raster order; the destination is the
    # pixel itself, the source the carried normalised coordinate scaled back to pixels and rounded half-to-even
    ty, tx = torch.nonzero(carried[..., 2] > 0.5, as_tuple=True)
    source = torch.round(carried[ty, tx, :2] * carried.new_tensor([W - 1, H - 1])).to(torch.int64)
    correspondences = torch.stack([source[:, 0], source[:, 1], tx, ty], dim=-1).to(device='cpu')
    bounds = normalize_depth(1.0 / depth, return_bounds=True)[1] if use_input_depth_normalization else None
    return normalize_depth(1.0 / target_depth, bounds=bounds), correspondences


def transform_depth(depth: torch.Tensor, bg_depth: torch.Tensor, fg_mask: torch.Tensor, intrinsics: torch.Tensor,
                    rot_angle: float = None, rot_axis: torch.Tensor = None, translation: torch.Tensor = None,
                    use_input_depth_normalization=False, depth_transform_mode: str = "pc"):
    """depth_transform.py:73-89."""
    if depth_transform_mode == "mesh":
        return transform_depth_mesh(depth=depth, bg_depth=bg_depth, fg_mask=fg_mask, intrinsics=intrinsics,
                                    rot_angle=rot_angle, rot_axis=rot_axis, translation=translation,
                                    use_input_depth_normalization=use_input_depth_normalization)
    elif depth_transform_mode == "pc":
        return transform_depth_pc(depth=depth, bg_depth=bg_depth, fg_mask=fg_mask, intrinsics=intrinsics,
                                  rot_angle=rot_angle, rot_axis=rot_axis, translation=translation,
                                  use_input_depth_normalization=use_input_depth_normalization)
    else:
        raise ValueError(f"Unknown depth transform mode '{depth_transform_mode}'.")

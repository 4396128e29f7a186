"""Mirror of the hot-path pieces of the reference's ``diffhandles/guided_stable_diffuser.py``:
``get_depth_intrinsics`` (:129-153), ``process_correspondences`` (:490-584) and the guidance weight
schedules (:336-373, :612-665).  The U-Net, VAE, text encoder and scheduler stay stock PyTorch/diffusers and
are outside this build (north star), so ``GuidedStableDiffuser`` here carries no model weights.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import _native as N
from .utils import unpack_correspondences

LATENT_GRID = 64   # guided_stable_diffuser.py:526-529: the loss grid is hard-wired to 64 x 64


class ProcessedCorrespondences(dict):
    """The reference's dict of ten NumPy int64 index arrays, plus the device-side int32 cell lists the CUDA
    loss kernels consume (``.device_lists``), so no index array is re-uploaded per loss call."""
    device_lists: Dict[str, torch.Tensor]
    grid: int


def process_correspondences_device(corr: torch.Tensor, img_res: int, bg_erosion: int = 0, grid: int = LATENT_GRID,
                                   device: torch.device = None) -> ProcessedCorrespondences:
    """corr (N,4) int64 (CPU or CUDA).  One kernel launch + one small device->host copy."""
    lib = N.load()
    if device is None:
        device = corr.device if corr.is_cuda else torch.device("cuda", torch.cuda.current_device())
    c = corr.to(device=device, dtype=torch.int64).reshape(-1, 4).contiguous()
    n = c.shape[0]
    cells = grid * grid
    i32 = torch.int32
    fg_src = torch.empty(max(n, 1), dtype=i32, device=device)
    fg_dst = torch.empty(max(n, 1), dtype=i32, device=device)
    bg = torch.empty(cells, dtype=i32, device=device)
    bg_o = torch.empty(cells, dtype=i32, device=device)
    bg_t = torch.empty(cells, dtype=i32, device=device)
    counts = torch.zeros(8, dtype=i32, device=device)
    N.check(lib.dh_process_correspondences(N.ptr(c) if n else None, n, int(img_res), grid, int(bg_erosion), N.ptr(fg_src),
                                           N.ptr(fg_dst), N.ptr(bg), N.ptr(bg_o), N.ptr(bg_t), N.ptr(counts),
                                           N.stream_handle(device)), "dh_process_correspondences")
    nv, nb, nbo, nbt = counts[:4].tolist()
    dl = {"fg_src": fg_src[:nv], "fg_dst": fg_dst[:nv], "bg": bg[:nb], "bg_orig": bg_o[:nbo], "bg_trans": bg_t[:nbt]}
    host = {k: v.cpu().numpy().astype(np.int64) for k, v in dl.items()}
    pc = ProcessedCorrespondences({
        'original_x': host["fg_src"] % grid, 'original_y': host["fg_src"] // grid,
        'transformed_x': host["fg_dst"] % grid, 'transformed_y': host["fg_dst"] // grid,
        'background_x': host["bg"] % grid, 'background_y': host["bg"] // grid,
        'background_x_orig': host["bg_orig"] % grid, 'background_y_orig': host["bg_orig"] // grid,
        'background_x_trans': host["bg_trans"] % grid, 'background_y_trans': host["bg_trans"] // grid,
    })
    pc.device_lists = dl
    pc.grid = grid
    return pc


class GuidanceWeightSchedule:
    """guided_stable_diffuser.py:612-620"""

    def __call__(self, denoising_step: int, optimization_step: int):
        return [1.0] * 3, [1.0] * 3


class StepGuidanceWeightSchedule(GuidanceWeightSchedule):
    """guided_stable_diffuser.py:622-665"""

    def __init__(self, denoising_steps, optimization_steps):
        super().__init__()
        if not all(len(fg) == len(bg) for _, fg, bg in denoising_steps):
            raise ValueError("Number of foreground and background weights do not match.")
        if not all(len(fg) == len(bg) for _, fg, bg in optimization_steps):
            raise ValueError("Number of foreground and background weights do not match.")
        if len(denoising_steps[0][1]) != len(optimization_steps[0][1]):
            raise ValueError("Number of denoising and optimization weights do not match.")
        self.denoising_steps = sorted(denoising_steps, key=lambda step: step[0])
        self.optimization_steps = sorted(optimization_steps, key=lambda step: step[0])

    def __call__(self, denoising_step: int, optimization_step: int):
        d = o = None
        for step, fg, bg in reversed(self.denoising_steps):
            if denoising_step >= step:
                d = (fg, bg)
                break
        for step, fg, bg in reversed(self.optimization_steps):
            if optimization_step >= step:
                o = (fg, bg)
                break
        if d is None or o is None:
            raise ValueError(f"Could not find weights for denoising step {denoising_step} and optimization step {optimization_step}.")
        return [a * b for a, b in zip(d[0], o[0])], [a * b for a, b in zip(d[1], o[1])]


def make_guidance_weight_schedule(fg_weight: float, bg_weight: float, guidance_max_step: int = 38,
                                  guidance_schedule_type: str = "constant") -> StepGuidanceWeightSchedule:
    """guided_stable_diffuser.py:336-373."""
    fg_weight = fg_weight * 30
    bg_weight = bg_weight * 30
    if guidance_schedule_type == "constant":
        ff = np.linspace(fg_weight, fg_weight, guidance_max_step)
        bf = np.linspace(bg_weight, bg_weight, guidance_max_step)
    elif guidance_schedule_type == "linear":
        ff = np.linspace(fg_weight, 0.0, guidance_max_step)
        bf = np.linspace(bg_weight, 0.0, guidance_max_step)
    elif guidance_schedule_type == "quadratic":
        ff = np.linspace(np.sqrt(fg_weight), 0.0, guidance_max_step) ** 2
        bf = np.linspace(np.sqrt(bg_weight), 0.0, guidance_max_step) ** 2
    else:
        raise ValueError(f"Unknown guidance schedule type: {guidance_schedule_type}")
    den = []
    for t_idx in range(guidance_max_step):
        if t_idx % 3 == 0:
            fw, bw = [0.0, 0.0, 7.5], [0.0, 0.0, 1.5]
        elif t_idx % 3 == 1:
            fw, bw = [0.0, 5.0, 0.0], [0.0, 1.5, 0.0]
        else:
            fw, bw = [0.0, 5.0, 7.5], [0.0, 1.5, 1.5]
        den.append((t_idx, (np.array(fw) * ff[t_idx]).tolist(), (np.array(bw) * bf[t_idx]).tolist()))
    den.append((guidance_max_step, [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]))
    opt = [(0, [2.5] * 3, [1.25] * 3), (1, [1.25] * 3, [2.5] * 3), (2, [1.25] * 3, [1.25] * 3), (3, [2.5] * 3, [2.5] * 3)]
    return StepGuidanceWeightSchedule(denoising_steps=den, optimization_steps=opt)


class GuidedStableDiffuser:
    """Carrier of the two hot-path methods of the reference class (same names / signatures)."""

    def __init__(self, conf=None):
        self.conf = conf
        self.device = torch.device("cpu")

    def to(self, device: torch.device = None):
        self.device = device
        return self

    @staticmethod
    def get_depth_intrinsics(device: torch.device = None):
        """guided_stable_diffuser.py:129-153: 55 degree FoV pinhole, principal point 0, image plane [-1,1]^2."""
        fov = 55.0
        f = 1.0 / np.tan(0.5 * fov * (np.pi / 180.0))
        return torch.tensor([[f, 0, 0.0], [0, f, 0.0], [0, 0, 1]], dtype=torch.float32, device=device)

    def process_correspondences(self, correspondences, img_res, bg_erosion=0):
        """guided_stable_diffuser.py:490-584 - returns the reference's dict of NumPy int64 arrays (a dict
        subclass that also keeps the device-side lists for the loss kernels)."""
        ox, oy, tx, ty = unpack_correspondences(correspondences)   # keeps the reference's (N,4) contract
        del ox, oy, tx, ty
        return process_correspondences_device(correspondences, img_res, bg_erosion)

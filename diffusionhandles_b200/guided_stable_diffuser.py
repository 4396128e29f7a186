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
    """Base schedule (guided_stable_diffuser.py:612-620): unit weights for the three guided layers at every step."""

    def __call__(self, denoising_step: int, optimization_step: int):
        return [1.0, 1.0, 1.0], [1.0, 1.0, 1.0]


class _StepTable:
    """Piecewise-constant weight table: rows (first_step, fg_weights, bg_weights); a step uses the row with the largest
    first_step that does not exceed it."""

    def __init__(self, rows, what: str):
        rows = sorted(rows, key=lambda r: r[0])
        for _, fg, bg in rows:
            if len(fg) != len(bg):
                raise ValueError("Number of foreground and background weights do not match.")
        self.what = what
        self.starts = [r[0] for r in rows]
        self.weights = [(list(r[1]), list(r[2])) for r in rows]
        self.width = len(rows[0][1])

    def at(self, step: int):
        import bisect
        i = bisect.bisect_right(self.starts, step) - 1
        return self.weights[i] if i >= 0 else None


class StepGuidanceWeightSchedule(GuidanceWeightSchedule):
    """guided_stable_diffuser.py:622-665: per-layer weights = (weights of the denoising step) x (weights of the optimisation
    iteration), each looked up in a piecewise-constant table.  Same constructor arguments and errors as the reference."""

    def __init__(self, denoising_steps, optimization_steps):
        self._denoising = _StepTable(denoising_steps, "denoising")
        self._optimization = _StepTable(optimization_steps, "optimization")
        if self._denoising.width != self._optimization.width:
            raise ValueError("Number of denoising and optimization weights do not match.")
        self.denoising_steps = list(zip(self._denoising.starts, *zip(*self._denoising.weights)))
        self.optimization_steps = list(zip(self._optimization.starts, *zip(*self._optimization.weights)))

    def __call__(self, denoising_step: int, optimization_step: int):
        den, opt = self._denoising.at(denoising_step), self._optimization.at(optimization_step)
        if den is None or opt is None:
            raise ValueError(f"Could not find weights for denoising step {denoising_step} and optimization step {optimization_step}.")
        return [d * o for d, o in zip(den[0], opt[0])], [d * o for d, o in zip(den[1], opt[1])]


# guided_stable_diffuser.py:336-373 - which of the three guided layers is active in denoising step t (t mod 3), and the
# multipliers of the optimisation iterations 0..3
_LAYER_PATTERN = (((0.0, 0.0, 7.5), (0.0, 0.0, 1.5)),
                  ((0.0, 5.0, 0.0), (0.0, 1.5, 0.0)),
                  ((0.0, 5.0, 7.5), (0.0, 1.5, 1.5)))
_ITERATION_GAIN = ((2.5, 1.25), (1.25, 2.5), (1.25, 1.25), (2.5, 2.5))


def _fade(weight: float, steps: int, kind: str) -> np.ndarray:
    """Weight over the guided denoising steps: constant, linear to zero, or quadratic to zero (np.linspace like the reference)."""
    if kind == "constant":
        return np.linspace(weight, weight, steps)
    if kind == "linear":
        return np.linspace(weight, 0.0, steps)
    if kind == "quadratic":
        return np.linspace(np.sqrt(weight), 0.0, steps) ** 2
    raise ValueError(f"Unknown guidance schedule type: {kind}")


def make_guidance_weight_schedule(fg_weight: float, bg_weight: float, guidance_max_step: int = 38,
                                  guidance_schedule_type: str = "constant") -> StepGuidanceWeightSchedule:
    """The schedule ``guided_inference`` builds (guided_stable_diffuser.py:336-373): user weights x 30, faded over the first
    ``guidance_max_step`` denoising steps, distributed over the layers by ``t mod 3``, zero afterwards."""
    fg_fade = _fade(fg_weight * 30, guidance_max_step, guidance_schedule_type)
    bg_fade = _fade(bg_weight * 30, guidance_max_step, guidance_schedule_type)
    denoising = []
    for t in range(guidance_max_step):
        fg_pat, bg_pat = _LAYER_PATTERN[t % 3]
        denoising.append((t, (np.array(fg_pat) * fg_fade[t]).tolist(), (np.array(bg_pat) * bg_fade[t]).tolist()))
    denoising.append((guidance_max_step, [0.0] * 3, [0.0] * 3))
    optimization = [(i, [f] * 3, [b] * 3) for i, (f, b) in enumerate(_ITERATION_GAIN)]
    return StepGuidanceWeightSchedule(denoising_steps=denoising, optimization_steps=optimization)


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

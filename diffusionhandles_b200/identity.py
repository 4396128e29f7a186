"""Input-image identity: the recorded activation stacks and the inversion results that the
reference's drivers cache as ``input_image_identity.npz`` (test/test_diffusion_handles.py:85-114,
webapp/webapps/diffhandles_webapp.py:82-96): keys ``null_text_emb``, ``init_noise``, ``activations1..3``,
``latent_image``.  The stacks are ~1 GB fp32 per image; they are loaded once through a pinned staging buffer and stay
resident on the device, so every guidance step reads them from HBM.
"""
from __future__ import annotations

import struct
import zipfile
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

KEYS = ("null_text_emb", "init_noise", "activations1", "activations2", "activations3", "latent_image")


@dataclass
class InputImageIdentity:
    null_text_emb: torch.Tensor
    init_noise: torch.Tensor
    activations: List[torch.Tensor]      # three (T, C, h, w) fp32 stacks, device resident
    latent_image: torch.Tensor

    def recorded(self, t_idx: int) -> List[torch.Tensor]:
        """The three recorded (C,h,w) activation maps of denoising step ``t_idx`` (contiguous views, no copy)."""
        return [a[t_idx] for a in self.activations]

    def nbytes(self) -> int:
        return sum(a.numel() * a.element_size() for a in self.activations)

    def prewarp(self, correspondences: torch.Tensor, img_res: int, winner_src: torch.Tensor = None) -> List[torch.Tensor]:
        """Warps ALL recorded timesteps of the three stacks through one edit's correspondences with K3 (one persistent TMA-staged
        launch: 1.05 GB read + 1.05 GB written for SD2-depth, 0.32 ms at the HBM peak): ``out[l][t, c, q] = A_l[t, c, src_l(q)]`` with
        ``src_l(q)`` the source cell of the first correspondence whose destination cell is q at level l (0 where there is none;
        with ``winner_src`` - (img_res^2,) int32 from the splat - cells without a correspondence fall back to their splat winner).
        The per-step guidance then compares aligned tensors instead of gathering (dense form)."""
        from . import warp
        dev = self.activations[0].device
        corr = correspondences.to(device=dev, dtype=torch.int64).reshape(1, -1, 4).contiguous()
        n = torch.tensor([corr.shape[1]], dtype=torch.int32, device=dev)
        if corr.shape[1] == 0:
            corr = torch.zeros((1, 1, 4), dtype=torch.int64, device=dev)
        sides = [int(a.shape[-1]) for a in self.activations]
        ws = winner_src.reshape(1, -1).to(device=dev, dtype=torch.int32).contiguous() if winner_src is not None else None
        maps1 = warp.dense_source_maps(corr, n, img_res, sides, ws)
        T = self.activations[0].shape[0]
        maps = [m.expand(T, -1).contiguous() for m in maps1]          # the same edit at every timestep: T "edits" for K3
        return warp.warp_stacks(self.activations, maps)


class ActivationRecorder:
    """Recording hook of the generation pass (guided_stable_diffuser.py:207-239, :269-272): the reference appends the three guided
    activations of every timestep to Python lists and ``torch.stack``s them at the end (the 1.05 GB exists twice at that point).
    Here the (T,C,h,w) stacks are allocated once, on the device the first activation lives on, and ``record(t_idx, acts)`` copies
    each (C,h,w) map into its slot on the current stream - the layout ``prewarp`` / K3 / K4 read, with no further copy."""

    def __init__(self, num_timesteps: int, dtype: torch.dtype = torch.float32):
        if num_timesteps < 1:
            raise ValueError("num_timesteps must be positive")
        self.num_timesteps = int(num_timesteps)
        self.dtype = dtype
        self._stacks: List[torch.Tensor] = []
        self._seen = [False] * self.num_timesteps

    def record(self, t_idx: int, activations) -> None:
        """activations: the three (C,h,w) maps of timestep ``t_idx`` (a leading batch dimension of 1 is accepted)."""
        if not 0 <= t_idx < self.num_timesteps:
            raise IndexError(f"timestep index {t_idx} outside the {self.num_timesteps} recorded steps")
        acts = [a[0] if a.dim() == 4 else a for a in activations]
        if not self._stacks:
            self._stacks = [torch.empty((self.num_timesteps, *a.shape), dtype=self.dtype, device=a.device) for a in acts]
        if len(acts) != len(self._stacks):
            raise ValueError(f"expected {len(self._stacks)} activation maps, got {len(acts)}")
        for stack, a in zip(self._stacks, acts):
            if tuple(a.shape) != tuple(stack.shape[1:]):
                raise ValueError(f"activation shape {tuple(a.shape)} differs from the recorded {tuple(stack.shape[1:])}")
            stack[t_idx].copy_(a.detach(), non_blocking=True)
        self._seen[t_idx] = True

    def stacks(self) -> List[torch.Tensor]:
        """The recorded (T,C,h,w) stacks; every timestep must have been recorded."""
        if not all(self._seen):
            raise RuntimeError(f"timesteps {[i for i, s in enumerate(self._seen) if not s]} were never recorded")
        return self._stacks

    def identity(self, null_text_emb: torch.Tensor, init_noise: torch.Tensor, latent_image: torch.Tensor) -> "InputImageIdentity":
        return InputImageIdentity(null_text_emb=null_text_emb, init_noise=init_noise, activations=self.stacks(), latent_image=latent_image)


def save_identity(path: str, identity: InputImageIdentity) -> None:
    """Writes the reference's ``.npz`` layout (np.savez, uncompressed)."""
    arrays = {"null_text_emb": identity.null_text_emb, "init_noise": identity.init_noise, "latent_image": identity.latent_image}
    arrays.update({f"activations{i + 1}": a for i, a in enumerate(identity.activations)})
    np.savez(path, **{k: v.detach().cpu().numpy() for k, v in arrays.items()})


def npz_member_layout(path: str) -> Optional[Dict[str, Tuple[int, tuple, np.dtype]]]:
    """Where the arrays of an UNCOMPRESSED ``.npz`` (np.savez: ZIP_STORED members, each a ``.npy``) lie in the file:
    ``{name: (byte offset of the raw data, shape, dtype)}``.  None when a member is compressed, Fortran-ordered or an object
    array - the caller then falls back to ``np.load``."""
    out = {}
    with zipfile.ZipFile(path) as zf, open(path, "rb") as f:
        for info in zf.infolist():
            if not info.filename.endswith(".npy"):
                continue
            if info.compress_type != zipfile.ZIP_STORED:
                return None
            f.seek(info.header_offset)
            local = f.read(30)                       # local file header: signature, ..., name length @26, extra length @28
            if local[:4] != b"PK\x03\x04":
                return None
            name_len, extra_len = struct.unpack("<HH", local[26:30])
            f.seek(info.header_offset + 30 + name_len + extra_len)
            version = np.lib.format.read_magic(f)
            if version == (1, 0):
                shape, fortran, dtype = np.lib.format.read_array_header_1_0(f)
            elif version == (2, 0):
                shape, fortran, dtype = np.lib.format.read_array_header_2_0(f)
            else:
                return None
            if fortran or dtype.hasobject:
                return None
            out[info.filename[:-4]] = (f.tell(), tuple(shape), dtype)
    return out


class _PinnedRing:
    """Two pinned staging buffers; a buffer is refilled only after the device copy that read it has completed."""

    def __init__(self, device: torch.device, nbytes: int):
        self.bufs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event() for _ in range(2)]
        self.used = [False, False]
        self.i = 0
        self.device = device

    def upload(self, src: np.ndarray, dst: torch.Tensor) -> None:
        """src: contiguous 1-D uint8 view (<= buffer size) of host or memory-mapped bytes; dst: 1-D uint8 device view."""
        k, n = self.i, src.shape[0]
        if self.used[k]:
            self.events[k].synchronize()
        stage = self.bufs[k][:n]
        np.copyto(stage.numpy(), src)                # the disk read (page cache -> pinned memory) happens here
        dst.copy_(stage, non_blocking=True)
        self.events[k].record(torch.cuda.current_stream(self.device))
        self.used[k] = True
        self.i = 1 - k


def load_identity(path: str, device: torch.device, staging_bytes: int = 64 << 20) -> InputImageIdentity:
    """Reads the reference's ``.npz`` layout and places everything on ``device``.  For an uncompressed file (what ``np.savez``
    and ``save_identity`` write) the arrays are memory-mapped where they lie in the archive and streamed through two pinned
    staging buffers of ``staging_bytes``: reading chunk i+1 from disk overlaps the host->device copy of chunk i, and the 1.05 GB
    of stacks never exist as a pageable or pinned host copy.  Compressed archives take the ``np.load`` path."""
    device = torch.device(device)
    layout = npz_member_layout(path)
    names = set(layout) if layout is not None else None
    if names is None:
        with np.load(path) as z:
            names = set(z.files)
    missing = [k for k in KEYS if k not in names]
    if missing:
        raise KeyError(f"{path} is not an input-image identity file (missing {missing})")
    out = {}
    if layout is None or device.type != "cuda":
        with np.load(path) as z:
            for k in KEYS:
                host = torch.from_numpy(np.ascontiguousarray(z[k]))
                if device.type == "cuda":
                    host = host.pin_memory()
                out[k] = host.to(device, non_blocking=True)
    else:
        if staging_bytes < 4096:
            raise ValueError("staging_bytes must be at least 4096")
        with torch.cuda.device(device):
            ring = _PinnedRing(device, staging_bytes)
            for k in KEYS:
                offset, shape, dtype = layout[k]
                nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
                dst = torch.empty(shape, dtype=_torch_dtype(dtype), device=device)
                if nbytes:
                    raw = np.memmap(path, dtype=np.uint8, mode="r", offset=offset, shape=(nbytes,))
                    flat = dst.view(-1).view(torch.uint8)
                    for a in range(0, nbytes, staging_bytes):
                        b = min(a + staging_bytes, nbytes)
                        ring.upload(raw[a:b], flat[a:b])
                    del raw
                out[k] = dst
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()
    return InputImageIdentity(null_text_emb=out["null_text_emb"], init_noise=out["init_noise"],
                              activations=[out["activations1"], out["activations2"], out["activations3"]],
                              latent_image=out["latent_image"])


def _torch_dtype(dtype: np.dtype) -> torch.dtype:
    if dtype.byteorder == ">":
        raise TypeError(f"big-endian arrays ({dtype}) are not supported")
    return torch.from_numpy(np.empty(0, dtype=dtype)).dtype

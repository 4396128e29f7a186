"""Shared helpers for the parity tests."""
import hashlib

import numpy as np


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def f32_translation(t):
    """The reference feeds ``translation[i].item()`` of an fp32 tensor (depth_transform.py:241-243)."""
    return tuple(float(np.float32(v)) for v in t)


def load_photogen():
    """The reference's 20 bundled photogen scenes (tests/golden/photogen_inputs.npz, packed by oracle/make_golden_photogen.py)
    and the reference outputs pinned on them.  Returns (meta, get) with get(scene) -> (depth, bg_depth, mask) fp32 (512,512)."""
    import json
    import os
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(golden, "photogen_ref.json")) as f:
        meta = json.load(f)
    g = np.load(os.path.join(golden, "photogen_inputs.npz"))

    def get(scene):
        depth = g[f"{scene}/depth"].astype(np.float32)
        bg = g[f"{scene}/bg_depth"].astype(np.float32)
        mask = np.unpackbits(g[f"{scene}/mask"])[: 512 * 512].reshape(512, 512).astype(np.float32)
        return depth, bg, mask
    get.rows = lambda scene: (g[f"{scene}/set_foreground_rows_idx"].astype(np.int64), g[f"{scene}/set_foreground_rows"])
    return meta, get


_FILLED_BG = {}


def photogen_filled_bg(scene, get):
    """set_foreground of the scene by the ORACLE (SuperLU, bit-identical to the reference's - pinned by SHA in the CPU suite)."""
    if scene not in _FILLED_BG:
        import scipy.ndimage
        from oracle import dh_oracle as O
        depth, bg, mask = get(scene)
        _FILLED_BG[scene] = O.solve_laplacian_depth(depth, bg, scipy.ndimage.binary_dilation(mask, iterations=15))
    return _FILLED_BG[scene]

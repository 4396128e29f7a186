"""Shared helpers for the parity tests."""
import hashlib

import numpy as np


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def f32_translation(t):
    """The reference feeds ``translation[i].item()`` of an fp32 tensor (depth_transform.py:241-243)."""
    return tuple(float(np.float32(v)) for v in t)

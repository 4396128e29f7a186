"""diffusionhandles_b200 - B200-native (sm_100a) implementation of the DiffusionHandles activation-lifting and
3D-warp hot path behind the reference's Python API.  See DESIGN.md and include/dh_b200.h.

``install_as_diffhandles()`` registers this package's modules under the reference's import names
(``diffhandles.depth_transform``, ``diffhandles.losses`` ...) so existing callers switch over without edits.
"""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"

_MIRRORED = ("depth_transform", "losses", "renderer", "pytorch3d_renderer", "mesh", "utils", "guided_stable_diffuser",
             "diffusion_handles")


def install_as_diffhandles() -> None:
    """Alias ``diffhandles`` and its hot-path sub-modules to this package in ``sys.modules``."""
    pkg = importlib.import_module(__name__)
    sys.modules["diffhandles"] = pkg
    for name in _MIRRORED:
        sys.modules[f"diffhandles.{name}"] = importlib.import_module(f"{__name__}.{name}")


def __getattr__(name):
    if name == "DiffusionHandles":
        from .diffusion_handles import DiffusionHandles
        return DiffusionHandles
    raise AttributeError(name)

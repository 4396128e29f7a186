"""TEST INFRASTRUCTURE ONLY - loads the *real* reference (read-only, /root/reference) as ground truth.

The reference is pure Python (``diffhandles/*.py``).  Its default ``'pc'`` depth transform, the losses and
``process_correspondences`` run in this container once the third-party modules that are not installed
(``pytorch3d``, ``diffusers``, ``omegaconf``) are stubbed and the package ``__init__`` is bypassed
(SURVEY.md Appendix C).  The reference modules are registered under the alias ``refdh`` so that they
never collide with the product package.

``/root/reference`` does not exist on the GPU box: everything that needs this loader is either a
golden-vector generator (``oracle/make_golden.py``, run here, output committed under ``tests/golden/``)
or a test that is skipped when the tree is absent.  No reference source is copied into this repository.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("DH_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "pytorch3d", "pytorch3d.renderer", "pytorch3d.renderer.mesh", "pytorch3d.renderer.mesh.shader",
    "pytorch3d.renderer.blending", "pytorch3d.structures", "pytorch3d.structures.meshes",
    "diffusers", "diffusers.image_processor", "diffusers.configuration_utils", "diffusers.utils",
    "diffusers.utils.torch_utils", "omegaconf",
]

_cache = None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "diffhandles"))


class _Ref(types.SimpleNamespace):
    """Namespace with the reference modules: ``depth_transform``, ``losses``, ``utils``, ``gsd``."""


def load_reference() -> _Ref:
    """Import the reference's hot-path modules from ``REFERENCE_ROOT`` and return them.

    After the import every ``sys.modules['diffhandles*']`` entry is moved to ``refdh*`` so a later
    ``import diffhandles`` (the drop-in alias of the product) cannot pick up the reference.
    """
    global _cache
    if _cache is not None:
        return _cache
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    saved = {k: v for k, v in sys.modules.items() if k == "diffhandles" or k.startswith("diffhandles.")}
    for k in saved:
        del sys.modules[k]
    stubbed = []
    for n in _STUBS:
        if n not in sys.modules:
            sys.modules[n] = MagicMock()
            stubbed.append(n)
    sys.modules["pytorch3d.renderer.mesh.shader"].ShaderBase = type("ShaderBase", (), {})
    pkg = types.ModuleType("diffhandles")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "diffhandles")]
    sys.modules["diffhandles"] = pkg
    m = types.ModuleType("diffhandles.model")
    m.__path__ = []
    sys.modules["diffhandles.model"] = m
    sys.modules["diffhandles.model.unet_2d_condition"] = MagicMock()
    try:
        dt = importlib.import_module("diffhandles.depth_transform")
        ls = importlib.import_module("diffhandles.losses")
        ut = importlib.import_module("diffhandles.utils")
        gsd = importlib.import_module("diffhandles.guided_stable_diffuser")
    finally:
        for k in [k for k in sys.modules if k == "diffhandles" or k.startswith("diffhandles.")]:
            sys.modules["refdh" + k[len("diffhandles"):]] = sys.modules.pop(k)
        for n in stubbed:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    diffuser = gsd.GuidedStableDiffuser.__new__(gsd.GuidedStableDiffuser)
    _cache = _Ref(depth_transform=dt, losses=ls, utils=ut, gsd=gsd,
                  process_correspondences=diffuser.process_correspondences,
                  get_depth_intrinsics=gsd.GuidedStableDiffuser.get_depth_intrinsics)
    return _cache


def load_exr(path: str):
    """Read a single-channel fp32 EXR fixture of the reference (``test/data/photogen/*/depth.exr``)."""
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3:
        img = img[..., 0]
    return img

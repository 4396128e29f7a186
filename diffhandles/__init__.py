"""Drop-in import name of the reference package (``diffhandles/__init__.py:1``): ``import diffhandles`` and
``from diffhandles.depth_transform import transform_depth`` resolve to the sm_100a implementation in ``diffusionhandles_b200``
with no install call.  Every sub-module here IS the corresponding ``diffusionhandles_b200`` module (same object)."""
from diffusionhandles_b200.diffusion_handles import DiffusionHandles  # noqa: F401

__all__ = ["DiffusionHandles"]

"""Names of the reference's ``diffhandles/pytorch3d_renderer.py`` kept importable.  The classes are the vertex
splat renderer (``renderer.SplatRenderer``); a triangle rasteriser with pytorch3d semantics is not part of this
build (parity unpinned)."""
from .renderer import SplatRenderer as PyTorch3DRenderer, SplatRendererArgs as PyTorch3DRendererArgs  # noqa: F401

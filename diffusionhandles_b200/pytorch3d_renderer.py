"""Names of the reference's ``diffhandles/pytorch3d_renderer.py`` kept importable.  ``PyTorch3DRenderer`` is the sm_100a
hard z-buffer triangle rasteriser (``renderer.MeshRenderer``) that follows the published semantics of pytorch3d's
MeshRasterizer for the settings the hot path uses (faces_per_pixel = 1, hard blend); pytorch3d itself is not a dependency,
so bit parity with it is unpinned."""
from .renderer import MeshRenderer as PyTorch3DRenderer, SplatRendererArgs as PyTorch3DRendererArgs  # noqa: F401

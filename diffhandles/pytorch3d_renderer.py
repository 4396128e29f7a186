"""``diffhandles.pytorch3d_renderer`` -> ``diffusionhandles_b200.pytorch3d_renderer`` (the module object itself)."""
import sys

from diffusionhandles_b200 import pytorch3d_renderer as _impl

sys.modules[__name__] = _impl

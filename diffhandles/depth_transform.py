"""``diffhandles.depth_transform`` -> ``diffusionhandles_b200.depth_transform`` (the module object itself)."""
import sys

from diffusionhandles_b200 import depth_transform as _impl

sys.modules[__name__] = _impl

"""``diffhandles.diffusion_handles`` -> ``diffusionhandles_b200.diffusion_handles`` (the module object itself)."""
import sys

from diffusionhandles_b200 import diffusion_handles as _impl

sys.modules[__name__] = _impl

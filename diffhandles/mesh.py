"""``diffhandles.mesh`` -> ``diffusionhandles_b200.mesh`` (the module object itself)."""
import sys

from diffusionhandles_b200 import mesh as _impl

sys.modules[__name__] = _impl

"""``diffhandles.renderer`` -> ``diffusionhandles_b200.renderer`` (the module object itself)."""
import sys

from diffusionhandles_b200 import renderer as _impl

sys.modules[__name__] = _impl

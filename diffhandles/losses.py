"""``diffhandles.losses`` -> ``diffusionhandles_b200.losses`` (the module object itself)."""
import sys

from diffusionhandles_b200 import losses as _impl

sys.modules[__name__] = _impl

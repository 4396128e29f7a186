"""``diffhandles.utils`` -> ``diffusionhandles_b200.utils`` (the module object itself)."""
import sys

from diffusionhandles_b200 import utils as _impl

sys.modules[__name__] = _impl

"""``diffhandles.guided_stable_diffuser`` -> ``diffusionhandles_b200.guided_stable_diffuser`` (the module object itself)."""
import sys

from diffusionhandles_b200 import guided_stable_diffuser as _impl

sys.modules[__name__] = _impl

"""Alias of pangenomix_b200.plot (same module object)."""
import sys

import pangenomix_b200.plot as _impl

sys.modules[__name__] = _impl

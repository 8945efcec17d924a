"""Alias of pangenomix_b200.core_genome (same module object)."""
import sys

import pangenomix_b200.core_genome as _impl

sys.modules[__name__] = _impl

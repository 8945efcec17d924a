"""Alias of pangenomix_b200.sparse_utils (same module object)."""
import sys

import pangenomix_b200.sparse_utils as _impl

sys.modules[__name__] = _impl

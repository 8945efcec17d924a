"""Alias of pangenomix_b200.pangenome_analysis (same module object)."""
import sys

import pangenomix_b200.pangenome_analysis as _impl

sys.modules[__name__] = _impl

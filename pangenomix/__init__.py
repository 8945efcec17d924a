"""Import shim: ``import pangenomix.sparse_utils`` / ``pangenomix.pangenome_analysis`` /
``pangenomix.plot`` resolve to the B200 implementation, so scripts written against the
reference package layout (README.md:146, pangenome_analysis.py:22) run unchanged."""

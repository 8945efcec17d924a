"""pangenomix-b200: the pan/core rarefaction + Bernoulli-grid hot path of pangenomix on B200.

Drop-in modules (same names and call signatures as /root/reference/pangenomix):
``sparse_utils`` (read_lsdf, LightSparseDataFrame), ``pangenome_analysis``
(estimate_pan_core_size, fit_heaps_by_iteration, compute_bernoulli_grid_core_genome)
and ``plot`` (calculate_mean).  The compute runs in hand-written sm_100a CUDA kernels
behind the C-ABI declared in ``include/pgx.h``; there is no CPU fallback.
"""
__version__ = "0.1.0"

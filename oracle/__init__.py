"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's pan/core hot path.

Nothing under ``pangenomix_b200/`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import, link or execute it, and there only as the checker / the CPU arm --
never as the thing measured as the product or shipped.

Parity pin: the reference (AnnaLew/pangenomix) has NO tests and NO golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself: ``tests/golden/make_golden.py`` imports /root/reference in the build
container (statsmodels stubbed, see that script) and writes the fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function here
against them bit-for-bit (curves) / to 1e-12 relative (fp64 likelihoods).

Third-party arithmetic on the path that is not under /root/reference:
  * numpy legacy ``RandomState.shuffle`` (MT19937 + masked rejection), call site
    pangenome_analysis.py:84-85 -- no version pinned by the reference; restated in
    ``legacy_shuffle`` / ``pancore_ref.c`` and checked against numpy itself.
  * scipy ``minimize(L-BFGS-B)`` (pangenome_analysis.py:159-160) and ``curve_fit``
    (:46-47) stay on the host on both sides and are not restated.
"""
from .pancore_np import (  # noqa: F401
    bernoulli_grad,
    bernoulli_ll,
    calculate_mean,
    estimate_pan_core_size_direct,
    estimate_pan_core_size_minrank,
    fit_heaps_by_iteration,
    legacy_shuffle,
    legacy_shuffle_stream,
    pan_core_curves_direct,
    pan_core_curves_minrank,
)

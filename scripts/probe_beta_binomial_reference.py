"""can the reference's compute_beta_binomial_core_genome (pangenome_analysis.py:295-400)
be pinned on the section-8(d) tables?  Runs the LIVE reference in the build container (statsmodels is absent: a stub with
the one function it uses, durbin_watson, the textbook formula) and prints what it returns.

    python scripts/probe_beta_binomial_reference.py > profiles/r02/probe_beta_binomial_reference.log 2>&1
"""
import contextlib, io, os, sys, types, warnings
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference")
import numpy as np
import pandas as pd

stub = types.ModuleType("statsmodels"); stats = types.ModuleType("statsmodels.stats"); tools = types.ModuleType("statsmodels.stats.stattools")
tools.durbin_watson = lambda r: float(np.sum(np.diff(r) ** 2) / np.sum(np.asarray(r) ** 2))
stub.stats = stats; stats.stattools = tools
sys.modules.update({"statsmodels": stub, "statsmodels.stats": stats, "statsmodels.stats.stattools": tools})
import pangenomix.pangenome_analysis as ref_pa
from pangenomix_b200 import synth

for name in ("c1", "c2"):
    coo = synth.config_matrix(name)
    n = coo.shape[1]
    counts = pd.Series(np.bincount(np.bincount(coo.row, minlength=coo.shape[0]), minlength=n + 1)[1:], index=np.arange(1, n + 1))
    counts = counts[counts > 0]
    print("== %s: %d genes x %d genomes; genes present in all genomes: %d, in all but one: %d" % (
        name, coo.shape[0], n, int(counts.get(n, 0)), int(counts.get(n - 1, 0))), flush=True)
    for num_points in (10, 25):
        if num_points >= n:
            continue
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            try:
                out = ref_pa.compute_beta_binomial_core_genome(None, df_counts=counts, num_points=num_points, ks_iter=50)
                print("num_points %d ->\n%s" % (num_points, out.to_string()), flush=True)
            except Exception as exc:                       # noqa: BLE001
                print("num_points %d -> %s: %s" % (num_points, type(exc).__name__, exc), flush=True)
            for w in caught[:4]:
                print("   warning: %s" % str(w.message).splitlines()[0])

"""Where the time of estimate_pan_core_size goes (development aid)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pangenomix_b200 import _native, engine, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
import bench
coo = bench.load_matrix(name, 0, lambda: None)       # cached in /tmp per box
eng = engine.PanCoreEngine(coo)
n = eng.n_genomes
np.random.seed(1)
eng.estimate(64)
for rep in range(3):
    np.random.seed(1)
    t0 = time.perf_counter(); out = eng.estimate(iters); t1 = time.perf_counter()
    print("estimate(%d): %.1f ms (%.0f perms/s)" % (iters, (t1 - t0) * 1e3, iters / (t1 - t0)))
    del out
# pieces
np.random.seed(1)
t0 = time.perf_counter(); perms = engine.draw_legacy_permutations(n, iters); t1 = time.perf_counter()
print("rng alone: %.1f ms" % ((t1 - t0) * 1e3))
t0 = time.perf_counter(); out = np.empty((iters, 2 * n)); t1 = time.perf_counter(); out[:] = 1.0; t2 = time.perf_counter()
print("np.empty %.2f ms, first touch fill %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
t0 = time.perf_counter(); out[:] = 2.0; t1 = time.perf_counter()
print("second fill %.1f ms" % ((t1 - t0) * 1e3))
t0 = time.perf_counter(); got = eng.curves_host(perms, out_f64=True, out=out); t1 = time.perf_counter()
print("curves_host f64 into touched pageable memory: %.1f ms" % ((t1 - t0) * 1e3))
hp, owner = engine.pinned_empty((iters, 2 * n), np.float64)
t0 = time.perf_counter(); got = eng.curves_host(perms, out_f64=True, out=hp); t1 = time.perf_counter()
print("curves_host f64 into pinned memory: %.1f ms" % ((t1 - t0) * 1e3))

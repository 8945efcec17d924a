"""Timeline of pgx_estimate_pan_core (PGX_ESTIMATE_TRACE=1) for one call on a cached workload (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pangenomix_b200 import engine

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
coo = bench.load_matrix(name, 0, lambda: None)
eng = engine.PanCoreEngine(coo)
np.random.seed(1)
eng.estimate(64)
for rep in range(3):
    np.random.seed(1)
    t0 = time.perf_counter(); out = eng.estimate(iters); t1 = time.perf_counter()
    print("estimate(%d): %.1f ms" % (iters, (t1 - t0) * 1e3), flush=True)
    del out
for block in (0, 104, 416):
    os.environ["PGX_ESTIMATE_TRACE"] = "1"
    np.random.seed(1)
    t0 = time.perf_counter(); out = eng.estimate(iters, block=block); t1 = time.perf_counter()
    del os.environ["PGX_ESTIMATE_TRACE"]
    print("traced estimate(%d, block=%d): %.1f ms" % (iters, block, (t1 - t0) * 1e3), flush=True)
    del out

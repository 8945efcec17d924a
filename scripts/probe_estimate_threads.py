"""estimate(2000) on a cached workload against the host thread counts of the pipeline (development aid):
PGX_RNG_THREADS (Fisher-Yates workers of the shuffle stream) x PGX_COPY_THREADS (staging -> result copy)."""
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
print("host cores:", os.cpu_count(), flush=True)
os.system("lscpu | grep -E 'Model name|Thread|Core|Socket'")
for rng_threads in (os.environ.get("SWEEP_RNG", ",1,2,3,4,6").split(",")):
    for copy_threads in (os.environ.get("SWEEP_COPY", ",1,2,4").split(",")):
        for key, val in (("PGX_RNG_THREADS", rng_threads), ("PGX_COPY_THREADS", copy_threads)):
            if val:
                os.environ[key] = val
            else:
                os.environ.pop(key, None)
        best = 1e9
        for rep in range(4):
            np.random.seed(1)
            t0 = time.perf_counter(); out = eng.estimate(iters); best = min(best, time.perf_counter() - t0)
            del out
        print("rng workers %-2s copy threads %-2s: %.1f ms" % (rng_threads or "d", copy_threads or "d", best * 1e3), flush=True)

"""The host half of the host-buffer call alone (no GPU work): pgx_expand_deltas -- uint16 curve steps -> int32 curves,
the rebuild that bounds bench.py's e2e -- against the number of host threads, with the host it ran on.

    python scripts/probe_host_rebuild.py [n_perm] [n_genomes] > profiles/r02/probe_host_rebuild.log
"""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pangenomix_b200 import _native  # noqa: E402

n_perm = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
lib = _native.load()
print("host: %d logical CPUs" % os.cpu_count())
try:
    for line in subprocess.run(["lscpu"], capture_output=True, text=True).stdout.splitlines():
        if any(k in line for k in ("Model name", "Socket(s)", "Core(s) per socket", "Thread(s) per core", "NUMA node(s)", "L3 cache")):
            print("  " + " ".join(line.split()))
except OSError:
    pass
rng = np.random.RandomState(0)
deltas = rng.randint(0, 40, size=(n_perm, 2 * n), dtype=np.uint16)
out = np.empty((n_perm, 2 * n), dtype=np.int32)
out[:] = 0                                                      # pages touched before timing
moved = deltas.nbytes + out.nbytes
for threads in (1, 2, 4, 8, 12, 16, 24, 32, 48, 64):
    if threads > 2 * os.cpu_count():
        break
    best = None
    for _ in range(3):
        t = time.perf_counter()
        _native.check(lib.pgx_expand_deltas(deltas.ctypes.data, n_perm, n, out.ctypes.data, 0, threads))
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    print("%2d threads: %.2f ms per %d x %d curves = %.1f GB/s read + written (%.1f GB/s written)" % (
        threads, best * 1e3, n_perm, 2 * n, moved / best / 1e9, out.nbytes / best / 1e9), flush=True)
want = np.cumsum(deltas[:3, :n].astype(np.int64), axis=1)
assert np.array_equal(out[:3, :n], want)

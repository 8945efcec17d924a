"""Host-buffer path (pgx_pan_core_curves_host) on C4: permutations per second against the block size of
its three-slot pipeline, plus the planner / upload times of the table on this host (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from pangenomix_b200 import engine, plan

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
perms_n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
coo = bench.load_matrix(name, 0, lambda: None)
for rep in range(2):
    t = time.perf_counter(); hp = plan.build_host_plan(coo); t_plan = time.perf_counter() - t
    print("plan %s: %.3f s (%d host cores)" % (name, t_plan, os.cpu_count()))
t = time.perf_counter(); eng = engine.PanCoreEngine(coo, host_plan=hp); torch.cuda.synchronize()
print("upload: %.3f s" % (time.perf_counter() - t))
n = eng.n_genomes
h_perms, o1 = engine.pinned_empty((perms_n, n), np.uint16)
h_out, o2 = engine.pinned_empty((perms_n, 2 * n), np.int32)
np.random.seed(12345)
engine.draw_legacy_permutations(n, perms_n, out=h_perms)
ref = eng.curves_host(h_perms, out=h_out).copy()
for block in (0, 200, 400, 800, 1600, 3200):
    eng.curves_host(h_perms, out=h_out, perms_per_block=block)
    torch.cuda.synchronize()
    t = time.perf_counter()
    reps = 5
    for _ in range(reps):
        eng.curves_host(h_perms, out=h_out, perms_per_block=block)
    dt = (time.perf_counter() - t) / reps
    pass
    print("perms_per_block %5d: %.2f ms per %d perms = %.0f perms/s, %.1f GB/s over PCIe (both directions)" % (
        block, dt * 1e3, perms_n, perms_n / dt, (h_perms.nbytes + h_out.nbytes) / dt / 1e9))

"""Step time against the list-or-bitmap threshold for one table, in one process (development aid).

    python scripts/probe_threshold.py c5 256 100 143 200 286
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from pangenomix_b200 import _native, engine

name, n_perm = sys.argv[1], int(sys.argv[2])
thresholds = [int(v) for v in sys.argv[3:]]
coo = bench.load_matrix(name, 0, lambda: None)
n = coo.shape[1]
np.random.seed(12345)
perms = engine.draw_legacy_permutations(n, n_perm)
d_perms = torch.from_numpy(perms.view(np.int16)).cuda()
out = torch.empty((n_perm, 2 * n), dtype=torch.int32, device="cuda")
ref = None
for thr in thresholds:
    t = time.time()
    eng = engine.PanCoreEngine(coo, long_threshold=thr)
    plan_s = time.time() - t
    hp = eng.host_plan
    for _ in range(2):
        eng.curves_device(d_perms, out=out)
    torch.cuda.synchronize()
    steps = 6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.curves_device(d_perms, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    _native.profile_read()
    _native.profile_enable(True)
    for _ in range(2):
        eng.curves_device(d_perms, out=out)
    torch.cuda.synchronize()
    a, b, c, calls = _native.profile_read()
    _native.profile_enable(False)
    got = out[:4].cpu().numpy()
    if ref is None:
        ref = got
    same = bool(np.array_equal(ref, got))
    print("%s threshold %d: %d list rows, %d bitmap rows (%.0f MB), plan %.1f s; step %.3f ms per %d perms (%.0f perms/s); "
          "serialised: list %.3f probe %.3f prep+scan %.3f ms; curves equal to the first threshold's: %s" % (
              name, thr, hp.n_rows, hp.n_long, hp.bits.nbytes / 1e6, plan_s, ms, n_perm, n_perm / ms * 1e3,
              a / calls, b / calls, c / calls, same), flush=True)
    eng.close()
    del eng

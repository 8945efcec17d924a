"""Back-to-back steps of the device-resident call, timed as bench.py times them, and isolated calls (development aid).

    [PGX_... env] python scripts/probe_step.py c4 10000 [steps]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from pangenomix_b200 import _native, engine

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
n_perm = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
coo = bench.load_matrix(name, 0, lambda: None)
eng = engine.PanCoreEngine(coo)
n = eng.n_genomes
np.random.seed(12345)
perms = engine.draw_legacy_permutations(n, n_perm)
d_perms = torch.from_numpy(perms.view(np.int16)).cuda()
out = torch.empty((n_perm, 2 * n), dtype=torch.int32, device="cuda")
for _ in range(3):
    eng.curves_device(d_perms, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    eng.curves_device(d_perms, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
# isolated calls: the device idle before and after every call
isolated = []
for _ in range(9):
    torch.cuda.synchronize()
    e0.record()
    eng.curves_device(d_perms, out=out)
    e1.record()
    torch.cuda.synchronize()
    isolated.append(e0.elapsed_time(e1))
iso = sorted(isolated)[len(isolated) // 2]
_native.profile_read()
_native.profile_enable(True)
for _ in range(3):
    eng.curves_device(d_perms, out=out)
torch.cuda.synchronize()
a, b, c, calls = _native.profile_read()
_native.profile_enable(False)
env = " ".join("%s=%s" % kv for kv in sorted(os.environ.items()) if kv[0].startswith("PGX_"))
print("%s %d perms [%s]: step %.3f ms (%.0f perms/s), isolated call %.3f ms; serialised: list %.3f probe %.3f prep+scan %.3f ms" % (
    name, n_perm, env, ms, n_perm / ms * 1e3, iso, a / calls, b / calls, c / calls), flush=True)

"""How the list kernel and the probe kernel share the SMs inside ONE production step (two streams, no profiler):
both kernels append per-CTA / per-warp {kind, SM, start, end} records in %globaltimer nanoseconds (pgx_set_trace).

    python scripts/overlap_trace.py c4 10000 profiles/r02_overlap_timeline_c4.json
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from pangenomix_b200 import _native, engine

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
n_perm = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
out_path = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/overlap_timeline_%s.json" % name
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
ref = out[:64].cpu().numpy()
n_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1          # back-to-back steps; the LAST one is analysed
capacity = n_steps * (n_perm * max(1, eng.host_plan.n_superblocks) + 4096)
trace = torch.zeros(1 + 3 * capacity, dtype=torch.int64, device="cuda")
lib = _native.load()
_native.check(lib.pgx_set_trace(trace.data_ptr(), capacity))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n_steps):
    if i == n_steps - 1:
        e0.record()
    eng.curves_device(d_perms, out=out)
e1.record()
torch.cuda.synchronize()
_native.check(lib.pgx_set_trace(None, 0))
assert np.array_equal(out[:64].cpu().numpy(), ref)
t = trace.cpu().numpy()
count = int(t[0])
rec = t[1:1 + 3 * min(count, capacity)].reshape(-1, 3)
if n_steps > 1:
    # the list kernel leaves exactly one record per CTA and step: the last step starts with the last 148-CTA batch
    lst_all = np.flatnonzero((rec[:, 0] & 0xff) == 1)
    per_step = lst_all.size // n_steps
    last_start = rec[lst_all[np.argsort(rec[lst_all, 1])][-per_step:], 1].min()
    prev_end = np.sort(rec[lst_all, 2])[-per_step - 1] if lst_all.size > per_step else 0
    keep = rec[:, 1] > prev_end if prev_end else np.ones(rec.shape[0], dtype=bool)
    # probe warps of the last step: those that started after the previous step's scan, i.e. after prev_end
    rec = rec[keep]
    count = int(rec.shape[0])
kind, sm = rec[:, 0] & 0xff, rec[:, 0] >> 8
t0 = int(rec[:, 1].min())
start, end = (rec[:, 1] - t0) / 1e6, (rec[:, 2] - t0) / 1e6           # ms
lst, prb = kind == 1, kind == 2
list_span = (float(start[lst].min()), float(end[lst].max()))
probe_span = (float(start[prb].min()), float(end[prb].max()))
# probe warps alive over time, and the share of probe warp-time spent while list CTAs are resident
grid = np.linspace(0, max(list_span[1], probe_span[1]), 201)
alive_probe = [(int(((start[prb] <= x) & (end[prb] > x)).sum())) for x in grid]
alive_list = [(int(((start[lst] <= x) & (end[lst] > x)).sum())) for x in grid]
inside = np.clip(np.minimum(end[prb], list_span[1]) - np.maximum(start[prb], list_span[0]), 0, None).sum()
total = (end[prb] - start[prb]).sum()
per_sm_probe = np.bincount(sm[prb].astype(np.int64), weights=(end[prb] - start[prb]), minlength=int(sm.max()) + 1)
doc = {
    "workload": name, "perms": n_perm, "records": count, "step_ms_cuda_events": e0.elapsed_time(e1),
    "back_to_back_steps": n_steps, "analysed": "the last step",
    "list_ctas": int(lst.sum()), "probe_warps": int(prb.sum()),
    "list_kernel_span_ms": list_span, "probe_kernel_span_ms": probe_span,
    "probe_warp_time_inside_list_span_frac": float(inside / total),
    "probe_warps_finished_inside_list_span_frac": float((end[prb] <= list_span[1]).mean()),
    "mean_probe_warps_alive_per_sm_while_list_runs": float(np.mean([a for a, x in zip(alive_probe, grid) if list_span[0] <= x <= list_span[1]]) / (int(sm.max()) + 1)),
    "mean_probe_warps_alive_per_sm_after_list_ends": float(np.mean([a for a, x in zip(alive_probe, grid) if x > list_span[1]] or [0]) / (int(sm.max()) + 1)),
    "timeline_ms": [float(x) for x in grid], "list_ctas_alive": alive_list, "probe_warps_alive": alive_probe,
    "probe_warp_ms_per_sm_min_max": [float(per_sm_probe.min()), float(per_sm_probe.max())],
}
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
with open(out_path, "w") as f:
    json.dump(doc, f, indent=1)
print(json.dumps({k: v for k, v in doc.items() if not isinstance(v, list) or len(v) < 4}, indent=1))

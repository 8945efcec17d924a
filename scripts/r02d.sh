#!/bin/bash
# Round 2, fourth GPU call (1 GPU): per-kernel times of the new pipeline (ncu launch list), the host-buffer call
# against block size / copy threads with its per-block timeline, then suite + bench.
set -u
OUT=gpurun_out/r02d
mkdir -p "$OUT"
step() { echo "== $* ($(date +%T))" | tee -a "$OUT/steps.log"; }
step "probe_step"
{
python scripts/probe_step.py c4 10000
PGX_LIST_VARIANT=2 python scripts/probe_step.py c4 10000
PGX_PROBE_STRAGGLERS=32 python scripts/probe_step.py c4 10000
python scripts/probe_step.py c4 1250 40
python scripts/probe_step.py c2 1000 100
} > "$OUT/probe_step.log" 2>&1
step "ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
      --log-file "$OUT/r02d_launches_c4_10000perms.csv" python scripts/probe_step.py c4 10000 2 > "$OUT/ncu_launch.log" 2>&1
step "probe_e2e (block sizes)"
timeout 600 python scripts/probe_e2e.py c4 10000 > "$OUT/probe_e2e.log" 2>&1
for t in 4 8 16; do
  PGX_COPY_THREADS=$t timeout 600 python scripts/probe_e2e.py c4 10000 2>&1 | grep "perms_per_block     0" | sed "s/^/copy threads $t: /" >> "$OUT/probe_e2e.log"
done
step "e2e trace"
PGX_ESTIMATE_TRACE=1 timeout 300 python - > "$OUT/e2e_trace.log" 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench
from pangenomix_b200 import engine
coo = bench.load_matrix("c4", 0, lambda: None)
eng = engine.PanCoreEngine(coo)
n = eng.n_genomes
h_perms, o1 = engine.pinned_empty((10000, n), np.uint16)
h_out, o2 = engine.pinned_empty((10000, 2 * n), np.int32)
np.random.seed(12345)
engine.draw_legacy_permutations(n, 10000, out=h_perms)
os.environ.pop("PGX_ESTIMATE_TRACE")
eng.curves_host(h_perms, out=h_out)
os.environ["PGX_ESTIMATE_TRACE"] = "1"
eng.curves_host(h_perms, out=h_out)
PY
step "probe_api c4"
timeout 300 python scripts/probe_api.py c4 2000 > "$OUT/probe_api_c4.log" 2>&1
step "pytest -m gpu (defaults)"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_default.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "pytest -m gpu with PGX_WIDE_BINS=1 (int32 bins everywhere), without the C5 test"
PGX_WIDE_BINS=1 timeout 900 python -m pytest tests -m gpu -x -q -k "not c5" > "$OUT/pytest_wide.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "bench c4 defaults"
timeout 600 python bench.py > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
echo "rc=$?" | tee -a "$OUT/steps.log"
step "done"

#!/bin/bash
# Round 2, sixth GPU call (1 GPU): scatter-based prep kernel; suite; launch list; e2e; bench.
set -u
OUT=gpurun_out/r02f
mkdir -p "$OUT"
step() { echo "== $* ($(date +%T))" | tee -a "$OUT/steps.log"; }
step "pytest subset"
timeout 900 python -m pytest tests -m gpu -x -q -k "not c5 and not c4_full" > "$OUT/pytest_subset.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "probe_step"
{
python scripts/probe_step.py c4 10000
PGX_PREP_GATHER=1 python scripts/probe_step.py c4 10000
PGX_LIST_THREADS=896 python scripts/probe_step.py c4 10000
python scripts/probe_step.py c4 1250 40
python scripts/probe_step.py c2 1000 100
} > "$OUT/probe_step.log" 2>&1
step "ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
      --log-file "$OUT/r02f_launches_c4_10000perms.csv" python scripts/probe_step.py c4 10000 2 > "$OUT/ncu_launch.log" 2>&1
step "probe_e2e"
timeout 600 python scripts/probe_e2e.py c4 10000 > "$OUT/probe_e2e.log" 2>&1
step "probe_api c4"
timeout 300 python scripts/probe_api.py c4 2000 > "$OUT/probe_api_c4.log" 2>&1
step "bench c4"
timeout 600 python bench.py > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
echo "rc=$?" | tee -a "$OUT/steps.log"
step "wide subset"
PGX_WIDE_BINS=1 timeout 900 python -m pytest tests -m gpu -x -q -k "fixtures or int32 or permutations or degenerate or random or launch_shape" > "$OUT/pytest_wide.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "done"

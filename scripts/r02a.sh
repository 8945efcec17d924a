#!/bin/bash
# Round 2, first GPU call (1 GPU): the suite at HEAD (row balance on by default, never run on a GPU before),
# the A/B of the row balance and of the vectorised scan, the bench line, and fresh ncu evidence.
#   gpurun --timeout 1500 -- 'bash scripts/r02a.sh'
set -u
OUT=gpurun_out/r02a
mkdir -p "$OUT"
step() { echo "== $* ($(date +%T))" | tee -a "$OUT/steps.log"; }

step "pytest -m gpu (defaults)"
timeout 600 python -m pytest tests -m gpu -x -q > "$OUT/pytest_default.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"

step "probe_rarefy c4 row balance ON (thresholds 128,160,192)"
timeout 400 python scripts/probe_rarefy.py c4 10000 --thresholds 128,160,192 > "$OUT/probe_c4_balance_on.log" 2>&1
step "probe_rarefy c4 row balance OFF"
PGX_NO_ROW_BALANCE=1 timeout 300 python scripts/probe_rarefy.py c4 10000 --thresholds 128 > "$OUT/probe_c4_balance_off.log" 2>&1
step "probe_rarefy c4 scan v8"
PGX_SCAN_V8=1 timeout 300 python scripts/probe_rarefy.py c4 10000 --thresholds 128 > "$OUT/probe_c4_scan_v8.log" 2>&1

step "pytest -m gpu with PGX_SCAN_V8=1"
PGX_SCAN_V8=1 timeout 600 python -m pytest tests -m gpu -x -q > "$OUT/pytest_scan_v8.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"

step "bench c4 defaults"
timeout 600 python bench.py > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
echo "rc=$?" | tee -a "$OUT/steps.log"

step "probe_api c4"
timeout 300 python scripts/probe_api.py c4 2000 > "$OUT/probe_api_c4.log" 2>&1
PGX_ESTIMATE_SPIN=1 timeout 300 python scripts/probe_api.py c4 2000 > "$OUT/probe_api_c4_estimate_spin.log" 2>&1

step "ncu launch list + full capture (r02a)"
CMD="python bench.py --perms 10000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > "$OUT/plain_r02a.log" 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file "$OUT/r02a_launches_c4_10000perms.csv" $CMD > "$OUT/ncu_launch_r02a.log" 2>&1
echo "launch list rc=$?" | tee -a "$OUT/steps.log"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'list_kernel|probe_kernel|scan_kernel' -s 9 -c 3 \
      -o "$OUT/prof_r02a" $CMD > "$OUT/ncu_full_r02a.log" 2>&1
echo "full capture rc=$?" | tee -a "$OUT/steps.log"
step "done"

#!/bin/bash
# Round 2, second GPU call (1 GPU): A/B of the list-kernel variants (0 = round 1, 1 = distinct prefetch buffers at
# 48 registers, 2 = the same at 64 registers x 768 threads) and of the probe kernel's straggler mode; then the suite.
set -u
OUT=gpurun_out/r02b
mkdir -p "$OUT"
step() { echo "== $* ($(date +%T))" | tee -a "$OUT/steps.log"; }
for v in 0 1 2; do
  for st in 0 64; do
    step "probe_rarefy c4 variant $v stragglers $st"
    PGX_LIST_VARIANT=$v PGX_PROBE_STRAGGLERS=$st timeout 300 python scripts/probe_rarefy.py c4 10000 --thresholds 128 > "$OUT/probe_c4_v${v}_s${st}.log" 2>&1
  done
done
for st in 16 32 128 256; do
  step "probe_rarefy c4 variant 1 stragglers $st"
  PGX_LIST_VARIANT=1 PGX_PROBE_STRAGGLERS=$st timeout 300 python scripts/probe_rarefy.py c4 10000 --thresholds 96,128 > "$OUT/probe_c4_v1_s${st}.log" 2>&1
done
step "probe_rarefy c4 variant 1 threads 768 / 896"
PGX_LIST_THREADS=768 timeout 300 python scripts/probe_rarefy.py c4 10000 --thresholds 128 > "$OUT/probe_c4_v1_t768.log" 2>&1
PGX_LIST_THREADS=896 timeout 300 python scripts/probe_rarefy.py c4 10000 --thresholds 128 > "$OUT/probe_c4_v1_t896.log" 2>&1
step "probe_rarefy c2"
timeout 300 python scripts/probe_rarefy.py c2 1000 > "$OUT/probe_c2_default.log" 2>&1
PGX_LIST_VARIANT=0 PGX_PROBE_STRAGGLERS=0 timeout 300 python scripts/probe_rarefy.py c2 1000 > "$OUT/probe_c2_round1.log" 2>&1
step "pytest -m gpu (defaults)"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_default.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "done"

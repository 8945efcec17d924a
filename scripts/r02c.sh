#!/bin/bash
# Round 2, third GPU call (1 GPU): the restructured kernels (prep kernel + rank rows, packed uint16 bins, compact
# host transfer): suite, A/Bs of the switches as back-to-back steps, bench line.
set -u
OUT=gpurun_out/r02c
mkdir -p "$OUT"
step() { echo "== $* ($(date +%T))" | tee -a "$OUT/steps.log"; }
step "pytest -m gpu (defaults)"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_default.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "probe_step A/B"
{
python scripts/probe_step.py c4 10000
PGX_WIDE_BINS=1 python scripts/probe_step.py c4 10000
PGX_LIST_VARIANT=1 python scripts/probe_step.py c4 10000
PGX_PROBE_STRAGGLERS=0 python scripts/probe_step.py c4 10000
PGX_PROBE_STRAGGLERS=8 python scripts/probe_step.py c4 10000
PGX_PROBE_STRAGGLERS=32 python scripts/probe_step.py c4 10000
PGX_NO_OVERLAP=1 python scripts/probe_step.py c4 10000
PGX_LONG_THRESHOLD=96 python scripts/probe_step.py c4 10000
PGX_LONG_THRESHOLD=160 python scripts/probe_step.py c4 10000
PGX_LIST_THREADS=640 python scripts/probe_step.py c4 10000
PGX_LIST_THREADS=896 python scripts/probe_step.py c4 10000
python scripts/probe_step.py c4 1250 40
python scripts/probe_step.py c4 2500 40
python scripts/probe_step.py c2 1000 100
} > "$OUT/probe_step.log" 2>&1
step "pytest -m gpu with PGX_WIDE_BINS=1 (int32 bins everywhere)"
PGX_WIDE_BINS=1 timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_wide.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "bench c4 defaults"
timeout 600 python bench.py > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
echo "rc=$?" | tee -a "$OUT/steps.log"
step "probe_api c4"
timeout 300 python scripts/probe_api.py c4 2000 > "$OUT/probe_api_c4.log" 2>&1
PGX_ESTIMATE_TRACE=1 timeout 300 python scripts/probe_estimate_trace.py > "$OUT/estimate_trace_c4.log" 2>&1
step "done"

#!/bin/bash
# Round 2 (1 GPU): beta-binomial pieces after the run-length count kernel / pinned lanes / streaming raw words;
# the probe gate (PGX_PROBE_GATE_US) on back-to-back steps, isolated calls, the host-buffer call and the API.
set -u
OUT=gpurun_out/r02o
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_betabin.py -m gpu -x -q > "$OUT/pytest_betabin.log" 2>&1
echo "pytest betabin rc=$?" > "$OUT/steps.log"
timeout 600 python scripts/probe_betabin.py c2 c4 > "$OUT/probe_betabin.log" 2>&1
echo "probe betabin rc=$?" >> "$OUT/steps.log"
{
for g in 0 300 2000; do
  PGX_PROBE_GATE_US=$g python scripts/probe_step.py c4 10000 2>&1 | grep perms
  PGX_PROBE_GATE_US=$g python scripts/probe_step.py c4 1250 40 2>&1 | grep perms
done
PGX_PROBE_GATE_US=0 python scripts/probe_step.py c2 1000 40 2>&1 | grep perms
PGX_PROBE_GATE_US=2000 python scripts/probe_step.py c2 1000 40 2>&1 | grep perms
} > "$OUT/probe_gate.log" 2>&1
echo "probe gate rc=$?" >> "$OUT/steps.log"
{
for g in 0 2000; do
  echo "== PGX_PROBE_GATE_US=$g"
  PGX_PROBE_GATE_US=$g python scripts/probe_e2e.py c4 10000 2>&1 | grep -E "perms_per_block +0:"
  PGX_PROBE_GATE_US=$g python scripts/probe_api.py c4 2000 2>&1 | head -4
done
} > "$OUT/probe_gate_e2e.log" 2>&1
PGX_PROBE_GATE_US=2000 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixtures or int32_bins or permutations or estimate or c_host or ragged" > "$OUT/pytest_gate_subset.log" 2>&1
echo "pytest gate subset rc=$?" >> "$OUT/steps.log"
timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
echo "bench rc=$?" >> "$OUT/steps.log"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke rc=$?" >> "$OUT/steps.log"

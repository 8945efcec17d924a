#!/bin/bash
# Round 2 (1 GPU): the probe gate as a stream-level wait (cuStreamWaitValue32), on (default) against off
# (PGX_PROBE_GATE=0): back-to-back steps, isolated calls, the host-buffer call, the API; parity subset with the gate on.
set -u
OUT=gpurun_out/r02p
mkdir -p "$OUT"
{
for g in 0 1; do
  PGX_PROBE_GATE=$g python scripts/probe_step.py c4 10000 2>&1 | grep perms
  PGX_PROBE_GATE=$g python scripts/probe_step.py c4 1250 40 2>&1 | grep perms
  PGX_PROBE_GATE=$g python scripts/probe_step.py c2 1000 40 2>&1 | grep perms
  PGX_PROBE_GATE=$g python scripts/probe_step.py c1 100 40 2>&1 | grep perms
done
PGX_PROBE_GATE=0 python scripts/probe_step.py c5 256 6 2>&1 | grep perms
PGX_PROBE_GATE=1 python scripts/probe_step.py c5 256 6 2>&1 | grep perms
} > "$OUT/probe_gate.log" 2>&1
echo "probe gate rc=$?" > "$OUT/steps.log"
{
for g in 0 1; do
  echo "== PGX_PROBE_GATE=$g"
  PGX_PROBE_GATE=$g python scripts/probe_e2e.py c4 10000 2>&1 | grep -E "perms_per_block +(0|400|800):"
  PGX_PROBE_GATE=$g python scripts/probe_api.py c4 2000 2>&1 | head -4
done
} > "$OUT/probe_gate_e2e.log" 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_betabin.py -m gpu -x -q -k "not c5 and not c4" > "$OUT/pytest_gate_subset.log" 2>&1
echo "pytest gate subset rc=$?" >> "$OUT/steps.log"
timeout 300 python scripts/probe_betabin.py c4 2>&1 | grep marginals > "$OUT/probe_marginals.log"
timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
echo "bench rc=$?" >> "$OUT/steps.log"

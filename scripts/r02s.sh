#!/bin/bash
# Round 2 (1 GPU): the whole GPU suite at HEAD after the last changes (KS kernel: shared-memory instance + four loads
# in flight; staging floors), the beta-binomial probe, the C4 bench line, the host rebuild alone.
set -u
OUT=gpurun_out/r02s
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" > "$OUT/steps.log"
timeout 600 python scripts/probe_betabin.py c4 > "$OUT/probe_betabin.log" 2>&1
echo "probe betabin rc=$?" >> "$OUT/steps.log"
timeout 600 python bench.py > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
echo "bench rc=$?" >> "$OUT/steps.log"
timeout 200 python scripts/probe_host_rebuild.py 10000 10000 > "$OUT/probe_host_rebuild.log" 2>&1
echo "host rebuild rc=$?" >> "$OUT/steps.log"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke rc=$?" >> "$OUT/steps.log"

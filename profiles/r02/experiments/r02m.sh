#!/bin/bash
# Round 2, experiment (1 GPU): share of every block of the host-buffer call that travels as int32 curves copied by
# the device instead of uint16 steps rebuilt by host threads.
set -u
OUT=gpurun_out/r02m
mkdir -p "$OUT"
{
for f in 0 0.2 0.3 0.4 0.5; do echo "== PGX_DIRECT_FRACTION=$f"; PGX_DIRECT_FRACTION=$f python scripts/probe_e2e.py c4 10000 2>&1 | grep -E "perms_per_block +(0|400|800):"; done
echo "== PGX_DIRECT_FRACTION=0.3 PGX_COPY_THREADS=8"
PGX_DIRECT_FRACTION=0.3 PGX_COPY_THREADS=8 python scripts/probe_e2e.py c4 10000 2>&1 | grep -E "perms_per_block +0:"
} > "$OUT/probe_direct.log" 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pinned_int32 or fixtures or int32_bins or permutations or estimate or c_host" > "$OUT/pytest_subset.log" 2>&1
echo "rc=$?" >> "$OUT/probe_direct.log"
timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
echo "bench rc=$?" >> "$OUT/probe_direct.log"

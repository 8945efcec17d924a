#!/bin/bash
# Round 2 (1 GPU): the split transfer format of the host-buffer calls (uint16 heads + uint8 tails): parity tests,
# then the host-buffer call and the bench line of C4 with it, and the host-buffer call without it (PGX_SPLIT_HEAD=0).
set -u
OUT=gpurun_out/r02x
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_transfer or curves_match_reference_fixtures or estimate or not_permutations or ragged" > "$OUT/pytest_split.log" 2>&1
echo "pytest rc=$?" > "$OUT/steps.log"
{
echo "== split (default)"; python scripts/probe_e2e.py c4 10000 2>&1 | grep -E "perms_per_block +(0|400|800):"
echo "== PGX_SPLIT_HEAD=0"; PGX_SPLIT_HEAD=0 python scripts/probe_e2e.py c4 10000 2>&1 | grep -E "perms_per_block +(0|400|800):"
} > "$OUT/probe_e2e_split.log" 2>&1
echo "probe rc=$?" >> "$OUT/steps.log"
timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
echo "bench rc=$?" >> "$OUT/steps.log"

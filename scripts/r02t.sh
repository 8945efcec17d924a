#!/bin/bash
# Round 2 (1 GPU): a SUSTAINED timed region for the record (300 back-to-back steps of C4 = 2.3 s, against the 0.15 s of
# the default bench line), and the C5 full-size test with the marginals check added.
set -u
OUT=gpurun_out/r02t
mkdir -p "$OUT"
timeout 600 python bench.py --steps 300 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_c4_n1_sustained_300_steps.json" 2> "$OUT/bench_sustained.err"
echo "sustained rc=$?" > "$OUT/steps.log"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c5_full_size" > "$OUT/pytest_c5.log" 2>&1
echo "pytest c5 rc=$?" >> "$OUT/steps.log"

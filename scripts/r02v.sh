#!/bin/bash
# Round 2 (1 GPU): list-or-bitmap threshold on C5 (2 permutations per list CTA make list rows four times as costly as
# on C4, where the default 1.28 sqrt(N) was tuned).
set -u
OUT=gpurun_out/r02v
mkdir -p "$OUT"
timeout 900 python scripts/probe_threshold.py c5 256 286 200 143 100 > "$OUT/probe_threshold_c5.log" 2>&1
echo "rc=$?" > "$OUT/steps.log"

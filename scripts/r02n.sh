#!/bin/bash
# Round 2 (1 GPU): first run of the beta-binomial pieces: GPU tests + timing probe.
set -u
OUT=gpurun_out/r02n
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_betabin.py -m gpu -x -q > "$OUT/pytest_betabin.log" 2>&1
echo "pytest rc=$?" > "$OUT/steps.log"
timeout 600 python scripts/probe_betabin.py c2 c4 > "$OUT/probe_betabin.log" 2>&1
echo "probe rc=$?" >> "$OUT/steps.log"

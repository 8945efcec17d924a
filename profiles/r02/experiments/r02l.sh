#!/bin/bash
# Round 2, experiment (1 GPU): blocks in flight of the host-buffer call (PGX_SLOTS), and the same call without the
# host rebuild (PGX_SKIP_REBUILD: wrong results, timing only) to see which side bounds it.
set -u
OUT=gpurun_out/r02l
mkdir -p "$OUT"
{
for s in 3 4 5 6; do echo "== PGX_SLOTS=$s"; PGX_SLOTS=$s python scripts/probe_e2e.py c4 10000 2>&1 | grep perms_per_block; done
echo "== PGX_SLOTS=4 PGX_SKIP_REBUILD=1 (timing only)"
PGX_SLOTS=4 PGX_SKIP_REBUILD=1 python scripts/probe_e2e_noassert.py c4 10000 2>&1 | grep perms_per_block
echo "== PGX_SLOTS=6 PGX_COPY_THREADS=8"
PGX_SLOTS=6 PGX_COPY_THREADS=8 python scripts/probe_e2e.py c4 10000 2>&1 | grep "perms_per_block     0"
echo "== api"
PGX_SLOTS=3 python scripts/probe_api.py c4 2000 2>&1 | head -4
PGX_SLOTS=4 python scripts/probe_api.py c4 2000 2>&1 | head -4
PGX_SLOTS=6 python scripts/probe_api.py c4 2000 2>&1 | head -4
} > "$OUT/probe_slots.log" 2>&1

#!/bin/bash
# Round 2, 4-GPU call with the probe gate: C4 weak and strong scaling at 4 ranks.
#   gpurun --gpus 4 --timeout 600 -- 'bash scripts/r02u.sh'
set -u
OUT=gpurun_out/r02u
mkdir -p "$OUT"
run() {  # name, ranks, bench args...
  local name=$1 ranks=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$ranks" --master-addr 127.0.0.1 \
      --master-port 29541 bench.py --gpus "$ranks" "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  echo "$name rc=$?" >> "$OUT/steps.log"
}
run bench_c4_weak_n4 4 --no-cpu-baseline
run bench_c4_strong_n4 4 --scaling strong --no-cpu-baseline --no-e2e

#!/bin/bash
# Round 2, 2-GPU call after the probe gate: the sharded API test on real GPUs, C4 weak and strong scaling at 2 ranks.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/r02r.sh'
set -u
OUT=gpurun_out/r02r
mkdir -p "$OUT"
step() { echo "== $* ($(date +%T))" | tee -a "$OUT/steps.log"; }
run() {  # name, ranks, bench args...
  local name=$1 ranks=$2; shift 2
  step "$name"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$ranks" --master-addr 127.0.0.1 \
      --master-port 29541 bench.py --gpus "$ranks" "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  echo "rc=$?" | tee -a "$OUT/steps.log"
}
step "pytest tests/test_distributed_gpu.py"
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q -s > "$OUT/pytest_distributed.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
run bench_c4_weak_n2 2 --no-cpu-baseline
run bench_c4_strong_n2 2 --scaling strong --no-cpu-baseline --no-e2e
step "host rebuild alone"
timeout 200 python scripts/probe_host_rebuild.py 10000 10000 > "$OUT/probe_host_rebuild.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "done"

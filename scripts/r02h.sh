#!/bin/bash
# Round 2, the 8-GPU call: sharded API at 8 ranks, C4 weak scaling at 8, C4 strong scaling at 1/2/4/8 (10,000
# permutations split over the GPUs, BASELINE.json's wording), C5 at its full size on 8 GPUs with the oracle check on.
#   gpurun --gpus 8 --timeout 1800 -- 'bash scripts/r02h.sh'
set -u
OUT=gpurun_out/r02h
mkdir -p "$OUT"
step() { echo "== $* ($(date +%T))" | tee -a "$OUT/steps.log"; }
run() {  # name, ranks, bench args...
  local name=$1 ranks=$2; shift 2
  step "$name"
  if [ "$ranks" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$ranks" --master-addr 127.0.0.1 \
        --master-port 29541 bench.py --gpus "$ranks" "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  fi
  echo "rc=$?" | tee -a "$OUT/steps.log"
}
nvidia-smi topo -m > "$OUT/topology.txt" 2>&1
run bench_c4_weak_n8 8 --no-cpu-baseline
for n in 1 2 4 8; do
  run bench_c4_strong_n${n} $n --scaling strong --no-cpu-baseline --no-e2e
done
run bench_c4_strong_n8_e2e 8 --scaling strong --no-cpu-baseline
run bench_c5_n8 8 --workload c5 --steps 5 --warmup 3 --cpu-check --cpu-seconds 8 --no-e2e
step "pytest tests/test_distributed_gpu.py"
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q -s > "$OUT/pytest_distributed.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "done"

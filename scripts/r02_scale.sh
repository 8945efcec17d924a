#!/bin/bash
# The multi-GPU call of the next round, to be run AFTER scripts/r02_first.sh is green:
#
#   gpurun --gpus 8 --timeout 1800 -- 'bash scripts/r02_scale.sh 8'        (charged 8 x the box time: about 12 min)
#
# C4 at N GPUs with the peer-memory CurveGather and with the NCCL gather, then C5 at its full 1,000 permutations
# (generating the 50,000 x 2,000,000 table takes about 200 s per box, once; it is cached in /tmp for the second run).
set -u
N=${1:-8}
OUT=gpurun_out/r02_scale
mkdir -p "$OUT"
run() {  # name, extra env, bench args...
  local name=$1 envs=$2; shift 2
  echo "== $name ($(date +%T))" | tee -a "$OUT/steps.log"
  env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29531 bench.py --gpus "$N" "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  echo "rc=$?" | tee -a "$OUT/steps.log"
}
run bench_c4_n${N} "PGX_X=1" --no-cpu-baseline
run bench_c4_n${N}_nccl_gather "PGX_GATHER=nccl" --no-cpu-baseline --no-e2e --steps 10
run bench_c5_n${N} "PGX_X=1" --workload c5 --no-cpu-baseline --no-e2e --steps 5 --warmup 3
echo "== done ($(date +%T))" | tee -a "$OUT/steps.log"

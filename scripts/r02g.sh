#!/bin/bash
# Round 2, multi-GPU validation at 2 GPUs: the sharded API against the single-GPU table, then bench weak + strong.
#   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/r02g.sh 2'
set -u
N=${1:-2}
OUT=gpurun_out/r02g
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
step "1-GPU A/B of the list kernel's thread count and subset of the suite (GPU 0)"
{
CUDA_VISIBLE_DEVICES=0 python scripts/probe_step.py c4 10000
CUDA_VISIBLE_DEVICES=0 PGX_LIST_THREADS=832 python scripts/probe_step.py c4 10000
CUDA_VISIBLE_DEVICES=0 PGX_LIST_THREADS=960 python scripts/probe_step.py c4 10000
CUDA_VISIBLE_DEVICES=0 PGX_LIST_THREADS=1024 python scripts/probe_step.py c4 10000
} > "$OUT/probe_step.log" 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixtures or int32 or permutations or degenerate or random or launch_shape or wide_tables or estimate" > "$OUT/pytest_subset.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 8 --csv \
      --log-file "$OUT/r02g_launches_c4_10000perms.csv" python scripts/probe_step.py c4 10000 2 > "$OUT/ncu_launch.log" 2>&1
step "pytest tests/test_distributed_gpu.py"
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q -s > "$OUT/pytest_distributed.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
run bench_c4_weak_n${N} "$N" --no-cpu-baseline
run bench_c4_strong_n${N} "$N" --scaling strong --no-cpu-baseline
step "done"

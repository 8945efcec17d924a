#!/bin/bash
# ncu evidence for the Bernoulli-grid kernel (1 GPU); ncu follows a plain run of the same command.
TAG=${1:-r01}
CMD="python bench.py --workload c3 --steps 200 --warmup 3"
$CMD > gpurun_out/plain_bern_${TAG}.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'grid_kernel|finish_kernel' -s 6 -c 2 -o gpurun_out/prof_bern_${TAG} $CMD > gpurun_out/ncu_bern_${TAG}.log 2>&1
echo full rc=$?

#!/bin/bash
# ncu evidence for the row kernels (1 GPU).  Each ncu run follows a plain run of the same command.
#   bash scripts/ncu_capture.sh <tag> [extra bench args]
TAG=${1:-r01}
shift
CMD="python bench.py --perms 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $*"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo launchlist rc=$?
$CMD > gpurun_out/plain_${TAG}b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'list_kernel|probe_kernel' -s 6 -c 2 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo full rc=$?
tail -2 gpurun_out/ncu_full_${TAG}.log

CMD="python bench.py --perms 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_r1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo launchlist rc=$?
$CMD > gpurun_out/plain_r1b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:minrank -s 3 -c 2 -o gpurun_out/prof_minrank_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo full rc=$?
tail -3 gpurun_out/ncu_full.log

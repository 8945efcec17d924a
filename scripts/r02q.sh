#!/bin/bash
# Round 2, final evidence call (1 GPU): the whole suite at HEAD, the bench lines of every config, ncu launch list + full
# captures of every kernel, CTA timelines, the RNG-ceiling table.  No number printed under ncu is a bench value.
set -u
OUT=gpurun_out/r02q
mkdir -p "$OUT"
step() { echo "== $* ($(date +%T))" | tee -a "$OUT/steps.log"; }
step "pytest -m gpu (whole suite, defaults)"
timeout 1200 python -m pytest tests -m gpu -x -q -s > "$OUT/pytest_gpu.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"
step "bench c4"
timeout 600 python bench.py > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
echo "rc=$?" | tee -a "$OUT/steps.log"
step "bench --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref_c4.json" 2> "$OUT/bench_ref_c4.err"
echo "rc=$?" | tee -a "$OUT/steps.log"
step "bench c2 / c1 / c3 / c5"
timeout 600 python bench.py --workload c2 > "$OUT/bench_c2_n1.json" 2> "$OUT/bench_c2_n1.err"
timeout 600 python bench.py --workload c1 > "$OUT/bench_c1_n1.json" 2> "$OUT/bench_c1_n1.err"
timeout 600 python bench.py --workload c3 > "$OUT/bench_c3_40000_n1.json" 2> "$OUT/bench_c3_40000_n1.err"
timeout 600 python bench.py --workload c3 --genes 4000 > "$OUT/bench_c3_4000_n1.json" 2> "$OUT/bench_c3_4000_n1.err"
timeout 900 python bench.py --workload c5 --steps 5 --cpu-seconds 8 --no-e2e > "$OUT/bench_c5_n1.json" 2> "$OUT/bench_c5_n1.err"
echo "rc=$?" | tee -a "$OUT/steps.log"
step "ncu launch list + full captures"
CMD="python bench.py --perms 10000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > "$OUT/plain_r02q.log" 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file "$OUT/r02q_launches_c4_10000perms.csv" $CMD > "$OUT/ncu_launch.log" 2>&1
echo "launch list rc=$?" | tee -a "$OUT/steps.log"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'prep_|list_kernel|probe_kernel|scan_kernel' -s 12 -c 4 \
      -o "$OUT/prof_r02q_c4" $CMD > "$OUT/ncu_full_c4.log" 2>&1
echo "full c4 rc=$?" | tee -a "$OUT/steps.log"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'grid_kernel|finish_kernel' -s 20 -c 2 \
      -o "$OUT/prof_r02q_c3" python bench.py --workload c3 --steps 5 > "$OUT/ncu_full_c3.log" 2>&1
echo "full c3 rc=$?" | tee -a "$OUT/steps.log"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'heaps_kernel' -s 1 -c 1 \
      -o "$OUT/prof_r02q_heaps" python bench.py --perms 10000 --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full_heaps.log" 2>&1
echo "full heaps rc=$?" | tee -a "$OUT/steps.log"
step "beta-binomial pieces: probe + ncu"
timeout 600 python scripts/probe_betabin.py c2 c4 > "$OUT/probe_betabin.log" 2>&1
echo "probe rc=$?" | tee -a "$OUT/steps.log"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ks_kernel|coo_count_kernel|spectrum_kernel' -s 6 -c 6 \
      -o "$OUT/prof_r02q_betabin" python scripts/probe_betabin.py c4 > "$OUT/ncu_full_betabin.log" 2>&1
echo "full betabin rc=$?" | tee -a "$OUT/steps.log"
step "probe step (gate on / off)"
PGX_PROBE_GATE=1 python scripts/probe_step.py c4 10000 2>&1 | grep perms > "$OUT/probe_step_gate.log"
PGX_PROBE_GATE=0 python scripts/probe_step.py c4 10000 2>&1 | grep perms >> "$OUT/probe_step_gate.log"
step "CTA timelines"
timeout 300 python scripts/overlap_trace.py c4 10000 "$OUT/overlap_timeline_c4_isolated_call.json" 1 > "$OUT/overlap_isolated.log" 2>&1
timeout 300 python scripts/overlap_trace.py c4 10000 "$OUT/overlap_timeline_c4_back_to_back.json" 4 > "$OUT/overlap_back_to_back.log" 2>&1
step "rng ceiling"
timeout 900 python scripts/rng_ceiling.py c1 c2 c4 c5 > "$OUT/rng_ceiling.json" 2> "$OUT/rng_ceiling.err"
echo "rc=$?" | tee -a "$OUT/steps.log"
step "done"

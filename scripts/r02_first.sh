#!/bin/bash
# The first GPU call of the next round, as ONE command: everything round 1 left unmeasured
# when its GPU budget ran out, ordered so that a call cut short still leaves the most useful files behind.
#
#   gpurun --timeout 3000 -- 'bash scripts/r02_first.sh'          (1 GPU, about 35 box-minutes)
#
# Every log lands under gpurun_out/r02_first/.  Nothing here reads /root/reference or the oracle except the
# test suite (its checker) and bench.py's CPU arm.  No number printed by a run under ncu is a bench value.
set -u
OUT=gpurun_out/r02_first
mkdir -p "$OUT"
step() { echo "== $* ($(date +%T))" | tee -a "$OUT/steps.log"; }

# 1. the suite as committed (row balance ON is the default since the end of round 1, never run on a GPU before)
step "pytest -m gpu (defaults)"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_default.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"

# 2. A/B of the residue-balanced row grouping on the two sizes it was modelled on (list / probe / scan per call
#    and the overlapped step; thresholds around the default 128 because cheaper gathers move the break-even up)
step "probe_rarefy c4 row balance ON"
timeout 600 python scripts/probe_rarefy.py c4 10000 --thresholds 112,128,144,160,192 > "$OUT/probe_c4_balance_on.log" 2>&1
step "probe_rarefy c4 row balance OFF"
PGX_NO_ROW_BALANCE=1 timeout 600 python scripts/probe_rarefy.py c4 10000 --thresholds 128 > "$OUT/probe_c4_balance_off.log" 2>&1
step "probe_rarefy c2 row balance ON / OFF"
timeout 300 python scripts/probe_rarefy.py c2 1000 > "$OUT/probe_c2_balance_on.log" 2>&1
PGX_NO_ROW_BALANCE=1 timeout 300 python scripts/probe_rarefy.py c2 1000 > "$OUT/probe_c2_balance_off.log" 2>&1

# 3. the vectorised scan under the whole suite (six parity tests passed with it in round 1); default if green
step "pytest -m gpu with PGX_SCAN_V8=1"
PGX_SCAN_V8=1 timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_scan_v8.log" 2>&1
echo "rc=$?" | tee -a "$OUT/steps.log"

# 4. the bench line of the committed defaults, then the same with the old row order (the A/B as bench.py sees it)
step "bench c4 defaults"
timeout 600 python bench.py > "$OUT/bench_c4_n1.json" 2> "$OUT/bench_c4_n1.err"
step "bench c4 PGX_NO_ROW_BALANCE=1"
PGX_NO_ROW_BALANCE=1 timeout 600 python bench.py --no-cpu-baseline --steps 10 > "$OUT/bench_c4_n1_no_balance.json" 2> "$OUT/bench_c4_n1_no_balance.err"

# 4b. the reference-facing call after the CPU-only changes of round 1's last session (shuffle stream: 32-draw steps,
#     PDEP regime ends, persistent spinning worker pool): estimate(2000) on C4 was 29 ms; then the opt-in spin
#     hand-off of pgx_estimate_pan_core (PGX_ESTIMATE_SPIN=1, never run on a GPU before) and the old wake-up
#     behaviour of the pool (PGX_RNG_SPIN_MS=0) for the A/B; the per-block timeline of the default
step "probe_api c4 (default / PGX_ESTIMATE_SPIN=1 / PGX_RNG_SPIN_MS=0 / trace)"
timeout 300 python scripts/probe_api.py c4 2000 > "$OUT/probe_api_c4.log" 2>&1
PGX_ESTIMATE_SPIN=1 timeout 300 python scripts/probe_api.py c4 2000 > "$OUT/probe_api_c4_estimate_spin.log" 2>&1
PGX_RNG_SPIN_MS=0 timeout 300 python scripts/probe_api.py c4 2000 > "$OUT/probe_api_c4_pool_sleeps.log" 2>&1
PGX_ESTIMATE_TRACE=1 timeout 300 python scripts/probe_estimate_trace.py > "$OUT/estimate_trace_c4.log" 2>&1

# 5. ncu: launch list, then one full capture of the two row kernels (each after a plain run of the same command).
#    The counter that decides the A/B: l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld per LDS.128 of
#    list_kernel<8> (5.27 in profiles/r01f_*; the layout model says 4.08 x measured/model ratio of round 1).
step "ncu launch list + full capture (r02a)"
CMD="python bench.py --perms 10000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > "$OUT/plain_r02a.log" 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file "$OUT/r02a_launches_c4_10000perms.csv" $CMD > "$OUT/ncu_launch_r02a.log" 2>&1
echo "launch list rc=$?" | tee -a "$OUT/steps.log"
$CMD > "$OUT/plain_r02a_b.log" 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'list_kernel|probe_kernel' -s 6 -c 2 \
      -o "$OUT/prof_r02a" $CMD > "$OUT/ncu_full_r02a.log" 2>&1
echo "full capture rc=$?" | tee -a "$OUT/steps.log"
step "done"

#!/bin/bash
# Round 2 (1 GPU), last call: smoke() and the beta-binomial tests at HEAD, ncu capture of the final ks_kernel and of the
# marginals kernels (the r02q capture shows the first version of ks_kernel).
set -u
OUT=gpurun_out/r02w
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke rc=$?" > "$OUT/steps.log"
timeout 600 python -m pytest tests/test_betabin.py -m gpu -x -q -k "not c4" > "$OUT/pytest_betabin.log" 2>&1
echo "pytest betabin rc=$?" >> "$OUT/steps.log"
timeout 300 python scripts/probe_betabin.py c2 > "$OUT/probe_betabin_plain.log" 2>&1
echo "probe rc=$?" >> "$OUT/steps.log"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ks_kernel' -s 4 -c 3 \
      -o "$OUT/prof_r02w_ks" python scripts/probe_betabin.py c2 > "$OUT/ncu_ks.log" 2>&1
echo "ncu ks rc=$?" >> "$OUT/steps.log"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'coo_count_kernel|spectrum_kernel' -c 3 \
      -o "$OUT/prof_r02w_marginals" python scripts/probe_betabin.py c2 > "$OUT/ncu_marginals.log" 2>&1
echo "ncu marginals rc=$?" >> "$OUT/steps.log"

#!/usr/bin/env python
"""Turns an ncu report (``ncu --set full``) into the two files profiles/ keeps: a CSV of the selected raw counters per
kernel, and the per-permutation figures bench.py scales into its ``roofline`` object (profiles/roofline_traffic.json).

    python scripts/ncu_summary.py gpurun_out/r02d/prof_r02d.ncu-rep r02d c4 10000
"""
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "sm__sass_inst_executed_op_shared_ld.sum", "sm__sass_inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
    "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", "sm__cycles_elapsed.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sectors.sum",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__cycles_active.avg", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def short_kernel_name(name):
    """``void pgx::<unnamed>::list_kernel<(int)8, (int)48, (bool)1>(pgx_plan, ...)`` -> ``list_kernel<(int)8, (int)48, (bool)1>``."""
    import re
    m = re.search(r"(\w*kernel\w*)", name)
    if not m:
        return name.split("(")[0].split("::")[-1]
    end = m.end()
    if end < len(name) and name[end] == "<":
        depth = 0
        for i in range(end, len(name)):
            depth += name[i] == "<"
            depth -= name[i] == ">"
            if depth == 0:
                return name[m.start():i + 1]
    return m.group(1)


def main():
    rep, tag, workload, perms = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    keep = [k for k in KEEP if k in idx]
    out_csv = os.path.join(REPO, "profiles", "%s_%s_ncu_raw_selected.csv" % (tag, workload))
    summary = {}
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for r in data:
            name = r[idx["Kernel Name"]]
            short = short_kernel_name(name)
            vals = {}
            for k in keep:
                w.writerow([short, k, units[idx[k]], r[idx[k]]])
                try:
                    vals[k] = float(r[idx[k]].replace(",", "")) * SCALE.get(units[idx[k]], 1.0)
                except ValueError:
                    pass
            kind = short
            for key, tag_ in (("list_kernel", "list"), ("probe_kernel", "probe"), ("scan_kernel", "scan"), ("prep_scatter_kernel", "prep"),
                              ("prep_kernel", "prep"), ("grid_kernel", "grid"), ("finish_kernel", "finish"), ("heaps_kernel", "heaps"),
                              ("ks_kernel", "ks"), ("coo_count_kernel", "coo_count"), ("spectrum_kernel", "spectrum")):
                if key in name:
                    kind = tag_
                    break
            lds = vals.get("smsp__sass_inst_executed_op_shared_ld.sum") or vals.get("smsp__inst_executed_op_shared_ld.sum")
            sms = (vals.get("sm__cycles_elapsed.sum", 0.0) / vals["sm__cycles_elapsed.avg"]) if vals.get("sm__cycles_elapsed.avg") else 148.0
            wavefronts = vals.get("l1tex__data_pipe_lsu_wavefronts.sum") or \
                (vals.get("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", 0.0) * sms)
            summary[kind] = {
                "kernel": short, "ms_under_ncu": vals.get("gpu__time_duration.sum"),
                "dram_bytes_per_perm": (vals.get("dram__bytes_read.sum", 0.0) + vals.get("dram__bytes_write.sum", 0.0)) / perms,
                "lsu_wavefronts_per_perm": wavefronts / perms or None,
                "lsu_wavefronts_shared_per_perm": vals.get("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", 0.0) * sms / perms,
                "lsu_wavefronts_global_per_perm": vals.get("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg", 0.0) * sms / perms,
                "lsu_pipe_pct_of_peak": vals.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                "shared_ld_wavefronts_per_perm": vals.get("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", 0.0) / perms,
                "wavefronts_per_shared_load": (vals.get("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", 0.0) / lds) if lds else None,
                "issue_active_pct": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active")
                or vals.get("smsp__issue_active.avg.pct_of_peak_sustained_elapsed"),
                "alu_pipe_pct": vals.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "fp64_pipe_pct": vals.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
                or vals.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                "active_lanes_per_instruction": vals.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                "l2_bytes_per_perm": (vals.get("lts__t_bytes.sum") or 32.0 * vals.get("lts__t_sectors.sum", 0.0)) / perms,
                "source": "profiles/%s_%s_ncu_raw_selected.csv (%d permutations per launch)" % (tag, workload, perms),
            }
    path = os.path.join(REPO, "profiles", "roofline_traffic.json")
    try:
        with open(path) as f:
            doc = json.load(f)
    except (OSError, ValueError):
        doc = {}
    doc.setdefault(workload, {}).update(summary)
    with open(path, "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    details = subprocess.run(["ncu", "-i", rep, "--page", "details"], check=True, capture_output=True, text=True).stdout
    with open(os.path.join(REPO, "profiles", "%s_%s_ncu_details.txt" % (tag, workload)), "w") as f:
        f.write(details)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()

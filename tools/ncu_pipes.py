#!/usr/bin/env python
"""Pipe utilisation, issue and stall breakdown per profiled launch from an .ncu-rep (raw page). usage: ncu_pipes.py rep [name-regex]"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]; pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else ".")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[0]
keys = [k for k in hdr if ("pipe" in k and ".avg.pct_of_peak_sustained_active" in k and k.startswith("sm__inst_executed")) or (k.startswith("l1tex__") and ("pct" in k or "wavefront" in k or "bank" in k)) or (k.startswith("smsp__inst_executed_op")) or k in (
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__time_duration.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors.sum", "sm__cycles_elapsed.avg",
    "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed")]
stall = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and "not_issued" not in k]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if not pat.search(name): continue
    print("==", name[:60])
    for k in keys:
        print(f"   {k:75s} {r[hdr.index(k)]:>14s} {rows[1][hdr.index(k)]}")
    st = sorted(((k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[hdr.index(k)] or 0)) for k in stall), key=lambda x: -x[1])
    print("   stalls (warps per issue):", " ".join(f"{k}={v:.2f}" for k, v in st[:8]))

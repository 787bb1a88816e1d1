#!/usr/bin/env python
"""One line per profiled launch from an .ncu-rep (raw page): time, SIMT efficiency, issue %, occupancy, caches, DRAM."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[0]
def g(r, k):
    try: return float(r[hdr.index(k)])
    except Exception: return float("nan")
print(f"{'kernel':18s} {'ms':>7s} {'thr/inst':>8s} {'issue%':>6s} {'warps%':>6s} {'Minst':>7s} {'regs':>4s} {'L1hit':>5s} {'L2hit':>5s} {'dramRdMB':>8s} {'dramWrMB':>8s} {'dram%':>5s}  top stalls")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0][:18]
    ur = rows[1][hdr.index("dram__bytes_read.sum")]; uw = rows[1][hdr.index("dram__bytes_write.sum")]
    sc = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}
    st = [(hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i])) for i in range(len(hdr))
          if hdr[i].startswith("smsp__average_warps_issue_stalled") and hdr[i].endswith("per_issue_active.ratio") and "not_issued" not in hdr[i] and r[i]]
    tops = " ".join(f"{k}={v:.1f}" for k, v in sorted(st, key=lambda x: -x[1])[:4])
    tu = rows[1][hdr.index("gpu__time_duration.sum")]; tms = g(r, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(tu, 1)
    print(f"{name:18s} {tms:7.3f} {g(r,'smsp__thread_inst_executed_per_inst_executed.ratio'):8.2f} {g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {g(r,'smsp__inst_executed.sum')/1e6:7.1f} {g(r,'launch__registers_per_thread'):4.0f} "
          f"{g(r,'l1tex__t_sector_hit_rate.pct'):5.1f} {g(r,'lts__t_sector_hit_rate.pct'):5.1f} {g(r,'dram__bytes_read.sum')*sc.get(ur,1):8.1f} {g(r,'dram__bytes_write.sum')*sc.get(uw,1):8.1f} "
          f"{g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}  {tops}")

#!/usr/bin/env python
"""Average DRAM traffic per launch of each kernel in an `ncu --set full` capture -> profiles/traffic.json entry.
usage: ncu_traffic.py report.ncu-rep <workload>/<integrator> [bench-log-of-the-same-run]
With the log (the JSON line bench.py printed under ncu: its device times are NOT bench values, but its ray / path
counts are exact) the entry also records the units each profiled launch processed, so bench.py can scale the
measured bytes to the launch size of the run it reports: bytes_per_unit = bytes_per_launch / units_per_launch."""
import csv, io, json, os, subprocess, sys, collections
rep, key = sys.argv[1], sys.argv[2]
units = {}
if len(sys.argv) > 3:
    for line in open(sys.argv[3]):
        if line.startswith("{") and '"kernels"' in line:
            units = {"k_" + k: v["units_per_launch"] for k, v in json.loads(line)["kernels"].items()}
            if key.endswith("/realtime"):      # timing buckets "bounce" / "film" are k_bounce_rt / k_rt_denoise in the realtime profile
                units["k_bounce_rt"] = units.get("k_bounce"); units["k_rt_denoise"] = units.get("k_film")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, un = rows[0], rows[1]
sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])     # launches, dram bytes, warp insts, thread insts, l2 hit x sectors, issue-active x time
def col(r, name, default=0.0):
    try: return float(r[hdr.index(name)])
    except Exception: return default
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
    b = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(m); b += float(r[i]) * sc.get(un[i], 1)
    a = agg[name]; a[0] += 1; a[1] += b
    wi = col(r, "smsp__inst_executed.sum"); a[2] += wi; a[3] += wi * col(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
    a[4] += col(r, "lts__t_sector_hit_rate.pct") * wi; a[5] += col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") * wi
path = os.environ.get("OHB_TRAFFIC_JSON") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[key] = {k: {"launches_profiled": n, "bytes_per_launch": tot / n, "units_per_launch": units.get(k), "source": os.path.basename(rep),
                 "inst_per_launch": wi / n, "thr_per_inst": (ti / wi) if wi else None, "l2_hit_pct": (l2 / wi) if wi else None, "issue_active_pct": (ia / wi) if wi else None}
             for k, (n, tot, wi, ti, l2, ia) in agg.items()}
for k, e in data[key].items():
    if e["units_per_launch"]:
        e["bytes_per_unit"] = e["bytes_per_launch"] / e["units_per_launch"]; e["warp_inst_per_unit"] = e["inst_per_launch"] / e["units_per_launch"]
json.dump(data, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(data[key], indent=1))

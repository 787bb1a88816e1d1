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
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, un = rows[0], rows[1]
sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
    b = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(m); b += float(r[i]) * sc.get(un[i], 1)
    agg[name][0] += 1; agg[name][1] += b
path = os.environ.get("OHB_TRAFFIC_JSON") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[key] = {k: {"launches_profiled": n, "bytes_per_launch": tot / n, "units_per_launch": units.get(k), "source": os.path.basename(rep)} for k, (n, tot) in agg.items()}
json.dump(data, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(data[key], indent=1))

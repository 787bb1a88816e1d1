# r2bc: ncu launch list (gpu__time_duration) of the default bench command with the final kernels
O=gpurun_out; TAG=r2bc
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_launches_helmet.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > $O/${TAG}_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2bc_launches_helmet.csv")) if len(r) > 10]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value"); u = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: t = float(r[v].replace(",", ""))
    except ValueError: continue
    t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[u], 1e-6)
    n = r[k].split("(")[0]; agg[n][0] += 1; agg[n][1] += t
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:12]: print("%-60s launches %4d  %8.3f ms  %5.1f %%" % (n[:60], a[0], a[1], 100 * a[1] / tot))
PY

# r2y: the driver's two bench commands with the final bench.py (issue roofline from the committed capture)
set -x
O=gpurun_out; TAG=r2y
( time timeout 900 python bench.py ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -3 $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2y_bench.json") if l.startswith("{")][0])
print("headline", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("frac_measured_bytes"), "cpu", d["cpu_baseline"]["value"])
for k, v in d["kernels"].items(): print("  ", k, v["bound"], "share", v["share_of_step"], "frac", v["frac"], "Grays/s", v.get("grays_per_s"), "winst/ray profiled", v.get("warp_inst_per_ray_profiled"), "implied live", v.get("warp_inst_per_ray_implied_live"), "B/unit", v.get("measured_dram_bytes_per_unit"))
for w in d.get("workloads", []):
    print(w["name"], w["integrator"], w["resolution"], round(w["value"], 1), "e2e", round(w["e2e"]["value"], 1), "SBE", round(w["rays"]["gsamples_sbe_per_s"], 3), "fps", w["frames_per_s"], "roofline", {k: w["roofline"].get(k) for k in ("kernel", "bound", "frac", "grays_per_s", "warp_inst_per_ray_profiled", "warp_inst_per_ray_implied_live", "thr_per_inst")}, "cpu", (w.get("cpu_baseline") or {}).get("value"))
PY

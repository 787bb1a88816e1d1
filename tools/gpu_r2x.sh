# r2x: what the driver runs at round end (full GPU suite, smoke, default bench, reference arm) + the TLAS stress A/B with concurrent BLAS builds
set -x
O=gpurun_out; TAG=r2x
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_tests.log 2>&1; tail -4 $O/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
( time timeout 900 python bench.py ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -3 $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
timeout 900 python tools/tlas_ab.py $O/${TAG}_tlas_ab.json > $O/${TAG}_tlas_ab.log 2>&1; grep -E "build_ms|update_ms|render_msamples" $O/${TAG}_tlas_ab.log
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2x_bench.json") if l.startswith("{")][0])
print("headline", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("frac_measured_bytes"))
for k, v in d["kernels"].items(): print("  ", k, v["bound"], "share", v["share_of_step"], "frac", v["frac"], "winst/ray profiled", v.get("warp_inst_per_ray_profiled"), "implied live", v.get("warp_inst_per_ray_implied_live"), "B/unit", v.get("measured_dram_bytes_per_unit"))
for w in d.get("workloads", []):
    print(w["name"], w["integrator"], w["resolution"], round(w["value"], 1), "SBE", round(w["rays"]["gsamples_sbe_per_s"], 3), "roofline", w["roofline"]["kernel"], w["roofline"]["bound"], w["roofline"]["frac"], "cpu", (w.get("cpu_baseline") or {}).get("value"))
    for k, v in w["kernels"].items(): print("     ", k, v)
PY

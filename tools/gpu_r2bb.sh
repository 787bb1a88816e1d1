# r2bb: what the driver runs at round end, with the final kernels: full GPU suite, smoke, default bench, reference arm
O=gpurun_out; TAG=r2bb
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_tests.log 2>&1; tail -5 $O/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
( time timeout 900 python bench.py ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -4 $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2bb_bench.json") if l.startswith("{")][0])
print("headline", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), d["roofline"]["kernel"], d["roofline"]["frac"], "launches", d["gpu_launches"], d["rays"])
for k, v in d["kernels"].items(): print("  ", k, v["bound"], "ms", v["ms"], "share", v["share_of_step"], "frac", round(v["frac"] or 0, 3), v.get("frac_measured_bytes"), "live winst/ray", v.get("warp_inst_per_ray_implied_live"))
for w in d.get("workloads", []):
    print(w["name"], w["integrator"], w["resolution"], round(w["value"], 1), "fps", w.get("frames_per_s"), "e2e", round(w["e2e"]["value"], 1), "SBE", round(w["rays"]["gsamples_sbe_per_s"], 3), "Mrays/s", round(w["rays"]["mrays_per_s"]), w["roofline"]["kernel"], w["roofline"]["frac"], "cpu", (w.get("cpu_baseline") or {}).get("value"))
print(d["cpu_baseline"]); print(d["clocks"])
PY

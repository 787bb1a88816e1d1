# r2n: full GPU test suite (with the BASELINE-scale parity tests) + the driver's bench line with the `workloads` array
set -x
O=gpurun_out; TAG=r2n
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > $O/${TAG}_tests.log 2>&1
tail -30 $O/${TAG}_tests.log
( time timeout 600 python bench.py ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -3 $O/${TAG}_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2n_bench.json") if l.startswith("{")][0])
print("headline", d["value"], "e2e", d["e2e"]["value"], d["roofline"])
for w in d.get("workloads", []): print(w["name"], w["integrator"], w["resolution"], round(w["value"], 1), "SBE", round(w["rays"]["gsamples_sbe_per_s"], 3), w["roofline"]["kernel"], w["roofline"]["frac"])
PY

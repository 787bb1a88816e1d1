# One GPU-box pass: tests, bench lines, ncu launch list, ncu --set full summaries (post-processed ON the box:
# the .ncu-rep files are too large for gpurun_out's 64 MiB limit, so only the text summaries come back).
# usage: bash tools/gpu_round_run.sh <tag> [notests]
TAG=${1:-r1}; set -x
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
if [ "$2" != "notests" ]; then
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_tests.log 2>&1
tail -5 $O/${TAG}_tests.log
fi
timeout 600 python bench.py > $O/${TAG}_bench_helmet.json 2> $O/${TAG}_bench_helmet.err
timeout 600 python bench.py --workload synthetic2m --no-cpu-baseline > $O/${TAG}_bench_s2m.json 2> $O/${TAG}_bench_s2m.err
timeout 600 python bench.py --workload synthetic2m --integrator realtime --steps 60 --no-cpu-baseline > $O/${TAG}_bench_s2m_rt.json 2> $O/${TAG}_bench_s2m_rt.err
timeout 600 python bench.py --workload cornell --no-cpu-baseline > $O/${TAG}_bench_cornell.json 2> $O/${TAG}_bench_cornell.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_helmet.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_b.log 2>&1
for WL in helmet synthetic2m; do
  REP=/tmp/${TAG}_full_$WL
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_(trace|surface|bounce|shade|film)' --launch-skip 0 -c 36 -o $REP python bench.py --workload $WL --steps 1 --warmup 3 --spp-step 4 --no-cpu-baseline > $O/${TAG}_ncu_full_$WL.log 2>&1
  python tools/ncu_summary.py $REP.ncu-rep > $O/${TAG}_ncu_full_$WL.txt 2>&1
  OHB_TRAFFIC_JSON=$O/${TAG}_traffic.json python tools/ncu_traffic.py $REP.ncu-rep $WL/offline $O/${TAG}_ncu_full_$WL.log > /dev/null 2>&1
  for K in k_trace_closest k_trace_shadow k_bounce k_surface k_shade; do
    python tools/ncu_lines.py $REP.ncu-rep "regex:^$K" 2 30 > $O/${TAG}_${WL}_${K}_lines.txt 2>&1
  done
done
ls -la $O /tmp/*.ncu-rep

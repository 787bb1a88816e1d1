# r2am (as r2v, after the shading-side changes; ncu runs with OHB_LANES=1 so that launch sizes match the serial kernel pass): the round's evidence pass on one GPU: driver-style bench line (+ workloads), reference arm, ncu launch list, ncu --set full
# captures of the four single-GPU configurations -> per-kernel DRAM bytes / warp instructions per unit (profiles/traffic.json)
TAG=${1:-r2am}
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_tests.log 2>&1; tail -6 $O/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
OHB_LANES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_launches_helmet.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > $O/${TAG}_ncu_launches.log 2>&1
run_full() {   # name, key, extra bench args
  REP=/tmp/${TAG}_full_$1
  OHB_LANES=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'^k_(trace|surface|bounce|shade|film|sort_hits|rt_pixel|rt_denoise)' --launch-skip 0 -c 44 -o $REP python bench.py $3 --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > $O/${TAG}_ncu_full_$1.log 2>&1
  python tools/ncu_summary.py $REP.ncu-rep > $O/${TAG}_ncu_full_$1.txt 2>&1
  OHB_TRAFFIC_JSON=$O/${TAG}_traffic.json python tools/ncu_traffic.py $REP.ncu-rep $2 $O/${TAG}_ncu_full_$1.log > /dev/null 2>> $O/${TAG}_traffic.err
}
run_full helmet helmet/offline "--workload helmet --spp-step 4"
run_full synthetic2m synthetic2m/offline "--workload synthetic2m --spp-step 4"
run_full synthetic2m_rt synthetic2m/realtime "--workload synthetic2m --integrator realtime"
run_full cornell512 cornell/offline "--workload cornell --width 512 --height 512 --spp-step 32"
for K in k_trace_closest k_shade; do python tools/ncu_lines.py /tmp/${TAG}_full_synthetic2m.ncu-rep "regex:^$K" 2 30 > $O/${TAG}_synthetic2m_${K}_lines.txt 2>&1; done
python tools/ncu_pipes.py /tmp/${TAG}_full_synthetic2m.ncu-rep > $O/${TAG}_ncu_pipes_synthetic2m.txt 2>&1
ls -la $O | tail -20

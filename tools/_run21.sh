O=gpurun_out; TAG=r1x; WL=synthetic2m; REP=/tmp/${TAG}_full_$WL
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_(trace|surface|bounce|shade|film)' --launch-skip 0 -c 28 -o $REP python bench.py --workload $WL --steps 1 --warmup 3 --spp-step 4 --no-cpu-baseline > $O/${TAG}_ncu_full_$WL.log 2>&1
python tools/ncu_summary.py $REP.ncu-rep > $O/${TAG}_ncu_full_$WL.txt 2>&1
OHB_TRAFFIC_JSON=$O/${TAG}_traffic.json python tools/ncu_traffic.py $REP.ncu-rep $WL/offline $O/${TAG}_ncu_full_$WL.log > /dev/null 2>&1
python tools/ncu_lines.py $REP.ncu-rep "regex:^k_shade" 1 30 > $O/${TAG}_${WL}_k_shade_lines.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_tests.log 2>&1; tail -3 $O/${TAG}_tests.log

set -x
O=gpurun_out; TAG=r2c
REP=/tmp/${TAG}_full
OHB_TRACE_OCC=7 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_trace' --launch-skip 2 -c 4 -o $REP python bench.py --workload synthetic2m --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline > $O/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py $REP.ncu-rep > $O/${TAG}_ncu_summary.txt 2>&1
python tools/ncu_pipes.py $REP.ncu-rep > $O/${TAG}_ncu_pipes.txt 2>&1
for K in k_trace_closest k_trace_shadow; do python tools/ncu_lines.py $REP.ncu-rep "regex:^$K" 2 40 > $O/${TAG}_${K}_lines.txt 2>&1; done
cat $O/${TAG}_ncu_summary.txt

set -x
O=gpurun_out; TAG=r2m
REP=/tmp/${TAG}_full
OHB_TRACE_OCC=7 timeout 900 ncu --set full --clock-control none -k regex:'^k_(trace|shade|sort_hits|film|raygen|advance)' --launch-skip 0 -c 40 -o $REP python bench.py --workload synthetic2m --steps 1 --warmup 0 --spp-step 4 --no-cpu-baseline > $O/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py $REP.ncu-rep > $O/${TAG}_ncu_summary.txt 2>&1
cat $O/${TAG}_ncu_summary.txt

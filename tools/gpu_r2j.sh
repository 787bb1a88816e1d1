set -x
O=gpurun_out; TAG=r2j
REP=/tmp/${TAG}_full
OHB_TRACE_OCC=7 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_trace_closest' --launch-skip 1 -c 1 -o $REP python bench.py --workload synthetic2m --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline > $O/${TAG}_ncu.log 2>&1
python tools/ncu_lines.py $REP.ncu-rep "regex:k_trace_closest" 0 50 > $O/${TAG}_lines.txt 2>&1
python tools/ncu_sass.py $REP.ncu-rep "regex:k_trace_closest" 0 90 > $O/${TAG}_sass.txt 2>&1

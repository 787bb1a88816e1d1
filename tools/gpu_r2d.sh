set -x
O=gpurun_out; TAG=r2d
( cd tools/ubench && ./l1 ) > $O/${TAG}_l1.txt 2>&1
REP=/tmp/${TAG}_full
OHB_TRACE_OCC=7 timeout 900 ncu --set full --clock-control none -k regex:'^k_trace' --launch-skip 2 -c 2 -o $REP python bench.py --workload synthetic2m --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline > $O/${TAG}_ncu.log 2>&1
python tools/ncu_pipes.py $REP.ncu-rep > $O/${TAG}_ncu_pipes.txt 2>&1
cat $O/${TAG}_l1.txt

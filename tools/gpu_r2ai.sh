# r2ai: where k_shade's instructions go on the helmet workload (source-line attribution, sorted by executed instructions)
O=gpurun_out; TAG=r2ai; REP=/tmp/${TAG}_shade
OHB_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_shade' --launch-skip 9 -c 3 -o $REP python bench.py --workload helmet --spp-step 4 --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > $O/${TAG}_ncu.log 2>&1
tail -3 $O/${TAG}_ncu.log
python tools/ncu_summary.py $REP.ncu-rep > $O/${TAG}_ncu_summary.txt 2>&1; cat $O/${TAG}_ncu_summary.txt
for L in 0 1; do python tools/ncu_lines.py $REP.ncu-rep "regex:^k_shade" $L 70 inst > $O/${TAG}_helmet_k_shade_launch${L}_lines_by_inst.txt 2>&1; done
head -5 $O/${TAG}_helmet_k_shade_launch0_lines_by_inst.txt

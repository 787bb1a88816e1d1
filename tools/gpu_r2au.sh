# r2au: instruction-level profile of k_shade (primary + first bounce iteration) on the helmet scene; report comes back for local reading
O=gpurun_out; TAG=r2au; REP=$O/${TAG}_shade
OHB_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_shade' --launch-skip 9 -c 2 -o $REP python bench.py --workload helmet --spp-step 4 --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > $O/${TAG}_ncu.log 2>&1
ls -la $REP.ncu-rep; python tools/ncu_summary.py $REP.ncu-rep

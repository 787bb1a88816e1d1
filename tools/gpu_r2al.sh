# r2al: instruction-level profile of k_trace_closest on the 2 M scene; the report comes back for local reading
O=gpurun_out; TAG=r2al; REP=$O/${TAG}_trace
OHB_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_trace_closest' --launch-skip 10 -c 1 -o $REP python bench.py --workload synthetic2m --spp-step 4 --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > $O/${TAG}_ncu.log 2>&1
tail -2 $O/${TAG}_ncu.log | cut -c1-200
ls -la $REP.ncu-rep
python tools/ncu_summary.py $REP.ncu-rep
python tools/ncu_lines.py $REP.ncu-rep "regex:^k_trace_closest" 0 80 inst > $O/${TAG}_s2m_trace_lines_by_inst.txt 2>&1
head -30 $O/${TAG}_s2m_trace_lines_by_inst.txt | cut -c1-180

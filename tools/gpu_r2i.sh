set -x
O=gpurun_out; TAG=r2i
REP=/tmp/${TAG}_full
OHB_TRACE_OCC=7 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_trace' --launch-skip 2 -c 2 -o $REP python bench.py --workload synthetic2m --steps 1 --warmup 1 --spp-step 4 --no-cpu-baseline > $O/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py $REP.ncu-rep > $O/${TAG}_ncu_summary.txt 2>&1
ncu -i $REP.ncu-rep --page raw --csv > /tmp/raw.csv
python - <<'PY' > gpurun_out/r2i_raw_l1.txt
import csv
rows = list(csv.reader(open('/tmp/raw.csv'))); hdr = rows[0]
for r in rows[2:3]:
    for i, k in enumerate(hdr):
        if k.startswith(('l1tex__', 'lts__', 'smsp__inst_executed', 'sm__inst_executed', 'smsp__warp', 'smsp__average', 'sm__cycles', 'smsp__cycles', 'smsp__issue', 'smsp__thread', 'smsp__pcsamp')) and r[i] not in ('0', '', 'n/a'):
            print(f"{k:95s} {r[i]:>18s} {rows[1][i]}")
PY
python tools/ncu_lines.py $REP.ncu-rep "regex:k_trace_closest" 2 50 > $O/${TAG}_lines.txt 2>&1

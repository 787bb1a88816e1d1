# A/B sweep of the traversal scheduling knobs (env overrides, ohb_kernels.cu traceKnobs) on the GPU box.
# usage: bash tools/gpu_sweep.sh <tag> "<env settings>;<env settings>;..." [workloads]
TAG=$1; IFS=';' read -ra CFGS <<< "$2"; WLS=${3:-"helmet synthetic2m"}
O=gpurun_out; : > $O/${TAG}_sweep.txt
for WL in $WLS; do for C in "${CFGS[@]}"; do
  L=$(env $C timeout 300 python bench.py --workload $WL --steps 6 --warmup 3 --no-cpu-baseline --no-workloads 2>>$O/${TAG}_sweep.err)
  echo "$L" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernels']
print('%-12s %-58s value %7.1f e2e %7.1f | ' % ('$WL', '$C', d['value'], d['e2e']['value']) + ' '.join('%s %.1f' % (n, k[n]['ms']) for n in sorted(k)))
" >> $O/${TAG}_sweep.txt 2>>$O/${TAG}_sweep.err
done; done
cat $O/${TAG}_sweep.txt

# r2ae: k_rt_denoise variants (identity passes skipped, table / squaring weights), A/B + parity

O=gpurun_out; TAG=r2ae
( timeout 900 python -m pytest tests/test_gpu_realtime.py -m gpu -x -q ) 2>&1 | tail -3
: > $O/${TAG}_sweep.txt
for C in "OHB_RT_DENOISE_VAR=1" "OHB_RT_DENOISE_VAR=2" "OHB_RT_DENOISE_VAR=3" "OHB_RT_DENOISE_VAR=0"; do
  L=$(env $C timeout 300 python bench.py --workload synthetic2m --integrator realtime --steps 60 --warmup 8 --no-cpu-baseline --no-workloads 2>>$O/${TAG}_sweep.err)
  echo "$L" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernels']
print('%-12s %-22s value %7.1f fps %6.1f e2e %7.1f | ' % ('s2m realtime', '$C', d['value'], d['config']['frames_per_s'], d['e2e']['value']) + ' '.join('%s %.1f' % (n, k[n]['ms']) for n in sorted(k)))
" >> $O/${TAG}_sweep.txt
done
cat $O/${TAG}_sweep.txt

# r2ak: one atan2 per miss (lookup + pdf), paired Sobol draws (sobolDraw2)
O=gpurun_out; TAG=r2ak
bash tools/gpu_sweep.sh $TAG "OHAO_B200_LIB=ab/lib_head.so;OHB_X=new" "helmet cornell synthetic2m"
L=$(timeout 300 python bench.py --workload synthetic2m --integrator realtime --steps 60 --warmup 8 --no-cpu-baseline --no-workloads 2>>$O/${TAG}_sweep.err)
echo "$L" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernels']
print('s2m realtime new value %7.1f fps %6.1f | ' % (d['value'], d['config']['frames_per_s']) + ' '.join('%s %.1f' % (n, k[n]['ms']) for n in sorted(k)))" >> $O/${TAG}_sweep.txt
L=$(OHAO_B200_LIB=ab/lib_head.so timeout 300 python bench.py --workload synthetic2m --integrator realtime --steps 60 --warmup 8 --no-cpu-baseline --no-workloads 2>>$O/${TAG}_sweep.err)
echo "$L" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernels']
print('s2m realtime head value %7.1f fps %6.1f | ' % (d['value'], d['config']['frames_per_s']) + ' '.join('%s %.1f' % (n, k[n]['ms']) for n in sorted(k)))" >> $O/${TAG}_sweep.txt
tail -2 $O/${TAG}_sweep.txt
( timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py tests/test_gpu_realtime.py -m gpu -x -q -k "per_sample or offline_samples or golden or psnr or 4k or env or realtime or pcg" ) 2>&1 | tail -2

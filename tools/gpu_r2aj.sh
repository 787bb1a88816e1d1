# r2aj: env-CDF search inside one line (2 dependent load rounds per 32 entries) and pow(x, 2.2) = x^2 * 2^(0.2 log2 x)
O=gpurun_out; TAG=r2aj
bash tools/gpu_sweep.sh $TAG "OHAO_B200_LIB=ab/lib_base.so;OHAO_B200_LIB=ab/lib_line.so;OHAO_B200_LIB=ab/lib_pow.so;OHB_X=both" "helmet synthetic2m"
( timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py tests/test_gpu_realtime.py -m gpu -x -q -s -k "per_sample or offline_samples or golden or psnr or 4k or env or realtime" ) 2>&1 | grep -v "^$" | tail -30 > $O/${TAG}_parity.txt
cat $O/${TAG}_parity.txt

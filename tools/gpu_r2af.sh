# r2af: sin/cos evaluation in the shading kernels — sinf + cosf (0), sincosf (1, in-tree), sincospif on turns (2)
O=gpurun_out; TAG=r2af
bash tools/gpu_sweep.sh $TAG "OHAO_B200_LIB=ab/libohao_b200_trig0.so;OHB_X=1;OHAO_B200_LIB=ab/libohao_b200_trig2.so" "helmet cornell synthetic2m"
for L in ab/libohao_b200_trig2.so; do
  ( OHAO_B200_LIB=$L timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q -s -k "per_sample or offline_samples or golden or psnr or pcg or 4k" ) 2>&1 | tail -40 > $O/${TAG}_parity_trig2.txt
done
cat $O/${TAG}_parity_trig2.txt

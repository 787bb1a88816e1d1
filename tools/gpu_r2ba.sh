# r2ba: deferred emission in k_shade / k_bounce_rt (results parked in place, the barrier-separated rounds after the shading)
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_realtime.py -m gpu -x -q -k "sample or golden or lanes or sum_mode or realtime or psnr" ) 2>&1 | tail -2
bash tools/gpu_sweep.sh r2ba "OHAO_B200_LIB=ab/lib_head.so;OHB_X=new;OHAO_B200_LIB=ab/lib_head.so;OHB_X=new" "helmet cornell"

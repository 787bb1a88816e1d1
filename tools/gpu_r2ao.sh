# r2ao: default scheduling knobs baked into the trace kernels (VAR bit 2)
O=gpurun_out; TAG=r2ao
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "bit_exact or schedules" ) 2>&1 | tail -2
bash tools/gpu_sweep.sh $TAG "OHAO_B200_LIB=ab/lib_head.so;OHB_X=new;OHAO_B200_LIB=ab/lib_head.so;OHB_X=new" "helmet synthetic2m"

# r2an: traversal loop trimmed — node pieces addressed by OR into the aligned node address, prefetch variants compiled out
O=gpurun_out; TAG=r2an
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_two_level.py -m gpu -x -q -k "bit_exact or two_level or tiny or coincident or far_from or schedules or tlas or refit or instances" ) 2>&1 | tail -2
bash tools/gpu_sweep.sh $TAG "OHAO_B200_LIB=ab/lib_head.so;OHB_X=new;OHAO_B200_LIB=ab/lib_head.so;OHB_X=new" "helmet synthetic2m"

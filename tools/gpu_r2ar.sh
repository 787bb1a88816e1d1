# r2ar: triangle test — fp64 fallback returns by value (no local-memory stores per test), two unrolled copies of the test body
O=gpurun_out; TAG=r2ar
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_two_level.py -m gpu -x -q -k "bit_exact or schedules or tiny or coincident or far_from or two_level" ) 2>&1 | tail -2
bash tools/gpu_sweep.sh $TAG "OHAO_B200_LIB=ab/lib_head.so;OHAO_B200_LIB=ab/lib_nounroll.so;OHB_X=unroll;OHAO_B200_LIB=ab/lib_head.so;OHAO_B200_LIB=ab/lib_nounroll.so;OHB_X=unroll" "helmet synthetic2m"

# r2az: k_sort_hits — the tile's entries requested before the barrier-separated rounds, hit / miss emitters behind one pair of barriers
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_realtime.py -m gpu -x -q -k "sample or golden or lanes or sum_mode or realtime or tiny or empty" ) 2>&1 | tail -2
bash tools/gpu_sweep.sh r2az "OHAO_B200_LIB=ab/lib_head.so;OHB_X=new;OHAO_B200_LIB=ab/lib_head.so;OHB_X=new" "helmet cornell"

# r2bd: any-hit queries visit inner children in slot order (no octant permutation of the meta bytes)
( timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact" ) 2>&1 | tail -1
bash tools/gpu_sweep.sh r2bd "OHAO_B200_LIB=ab/lib_head.so;OHB_X=new" "synthetic2m"

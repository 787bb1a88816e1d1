# r2w: octant-binned rays + octant-specialised node test, A/B and parity
set -x
O=gpurun_out; TAG=r2w
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_realtime.py -m gpu -x -q ) 2>&1 | tail -3
bash tools/gpu_sweep.sh $TAG "OHB_OCT_BIN=0;OHB_OCT_BIN=1;OHB_OCT_BIN=1 OHB_TRACE_OCC=7;OHB_OCT_BIN=1 OHB_TRACE_OCC=9" "synthetic2m helmet cornell"

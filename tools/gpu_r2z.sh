# r2z: STABLE octant binning of the ray queues (generic node test), A/B and parity
set -x
O=gpurun_out; TAG=r2z
( OHB_OCT_BIN=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_realtime.py -m gpu -x -q ) 2>&1 | tail -3
bash tools/gpu_sweep.sh $TAG "OHB_OCT_BIN=0;OHB_OCT_BIN=1" "synthetic2m helmet cornell"

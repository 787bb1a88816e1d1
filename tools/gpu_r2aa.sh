# r2aa: stage-specialised k_shade bodies, A/B + image identity
set -x
O=gpurun_out; TAG=r2aa
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q ) 2>&1 | tail -3
bash tools/gpu_sweep.sh $TAG "OHB_SHADE_STAGES=0;OHB_SHADE_STAGES=1" "helmet synthetic2m cornell"

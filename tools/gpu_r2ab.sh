# r2ab: LEAN k_shade bodies (Sobol / no anisotropy / no SSS folded at compile time), A/B + parity
set -x
O=gpurun_out; TAG=r2ab
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q ) 2>&1 | tail -3
bash tools/gpu_sweep.sh $TAG "OHB_SHADE_LEAN=0;OHB_SHADE_LEAN=1" "helmet synthetic2m cornell"

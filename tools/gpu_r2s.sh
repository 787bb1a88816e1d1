# r2s: updated BASELINE-scale tests + NRD packing on the GPU, then the k_shade software-prefetch A/B
set -x
O=gpurun_out; TAG=r2s
( timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_nrd_packing.py -m gpu -x -q -s ) > $O/${TAG}_tests.log 2>&1
grep -E "^\[|passed|failed|Error" $O/${TAG}_tests.log | tail -20
bash tools/gpu_sweep.sh $TAG "OHB_SHADE_PREFETCH=0;OHB_SHADE_PREFETCH=1;OHB_SHADE_PREFETCH=2" "helmet synthetic2m cornell"

# r2ac: blocked env-CDF search, A/B + env parity
set -x
O=gpurun_out; TAG=r2ac
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_assets.py -m gpu -x -q ) 2>&1 | tail -3
bash tools/gpu_sweep.sh $TAG "OHB_ENV_BLOCKED=0;OHB_ENV_BLOCKED=1" "helmet synthetic2m"

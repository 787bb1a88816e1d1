# r2ag: two lanes per batch on two streams (OHB_LANES) — bit-identity test + A/B
O=gpurun_out; TAG=r2ag
( timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "two_lanes or per_sample" ) 2>&1 | tail -5 > $O/${TAG}_tests.txt
cat $O/${TAG}_tests.txt
bash tools/gpu_sweep.sh $TAG "OHB_LANES=1;OHB_LANES=2" "helmet cornell synthetic2m"

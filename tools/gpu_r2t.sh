# r2t: two-level structure + refit, hybrid techniques and NRD packing on the GPU; TLAS stress A/B; regression of the parity suite
set -x
O=gpurun_out; TAG=r2t
( timeout 900 python -m pytest tests/test_two_level.py tests/test_hybrid_rt.py tests/test_nrd_packing.py tests/test_gpu_parity.py -m gpu -x -q -s ) > $O/${TAG}_tests.log 2>&1
grep -E "^\[|passed|failed|Error|error" $O/${TAG}_tests.log | tail -20
timeout 900 python tools/tlas_ab.py $O/${TAG}_tlas_ab.json > $O/${TAG}_tlas_ab.log 2>&1; tail -40 $O/${TAG}_tlas_ab.log

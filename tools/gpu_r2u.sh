# r2u: hybrid techniques / NRD packing / parity regression on the GPU, TLAS stress A/B (fixed warm-up), NCCL-from-C++ check is in r2r
set -x
O=gpurun_out; TAG=r2u
( timeout 1200 python -m pytest tests/test_hybrid_rt.py tests/test_nrd_packing.py tests/test_gpu_parity.py tests/test_gpu_realtime.py tests/test_gpu_svgf.py tests/test_host_cpp.py -m gpu -q -s ) > $O/${TAG}_tests.log 2>&1
grep -E "^\[|passed|failed|Error|error" $O/${TAG}_tests.log | tail -20
timeout 900 python tools/tlas_ab.py $O/${TAG}_tlas_ab.json > $O/${TAG}_tlas_ab.log 2>&1; tail -32 $O/${TAG}_tlas_ab.log

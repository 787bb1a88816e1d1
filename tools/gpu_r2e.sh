set -x
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > $O/r2e_tests.log 2>&1; tail -3 $O/r2e_tests.log
bash tools/gpu_sweep.sh r2e "OHB_TRACE_OCC=6;OHB_TRACE_OCC=7;OHB_TRACE_OCC=8" "synthetic2m helmet cornell"

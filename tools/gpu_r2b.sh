set -x
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > $O/r2b_tests.log 2>&1; tail -5 $O/r2b_tests.log
bash tools/gpu_sweep.sh r2b "OHB_TRACE_OCC=6;OHB_TRACE_OCC=7;OHB_TRACE_OCC=8;OHB_TRACE_OCC=9" "synthetic2m helmet"

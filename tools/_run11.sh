( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1n_tests.log 2>&1; tail -3 gpurun_out/r1n_tests.log
bash tools/gpu_sweep.sh r1n "OHB_POSTPONE_DEN=5;OHB_POSTPONE_DEN=0;OHB_POSTPONE_DEN=0 OHB_TRACE_MIN_ACTIVE=24"

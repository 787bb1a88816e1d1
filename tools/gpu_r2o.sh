# r2o: (1) where the per-sample differences on the 2 M scene come from, (2) A/B of the shared-memory traversal variants
set -x
O=gpurun_out; TAG=r2o
timeout 600 python tools/diag_parity_2m.py > $O/${TAG}_diag_parity_2m.txt 2>&1
cat $O/${TAG}_diag_parity_2m.txt
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "prim_ids or recorded" ) 2>&1 | tail -2
bash tools/gpu_sweep.sh $TAG "OHB_TRACE_VAR=0;OHB_TRACE_VAR=1;OHB_TRACE_VAR=2;OHB_TRACE_VAR=3;OHB_TRACE_VAR=0 OHB_TRACE_OCC=7;OHB_TRACE_VAR=1 OHB_TRACE_OCC=7;OHB_TRACE_VAR=3 OHB_TRACE_OCC=7" "synthetic2m helmet"

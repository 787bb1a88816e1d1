# r2av: -prec-div=false (scalar divisions of the shading code as MUFU.RCP + multiply) — full GPU suite, then A/B
O=gpurun_out; TAG=r2av
( timeout 1800 python -m pytest tests -m gpu -q -s ) > $O/${TAG}_tests.log 2>&1; grep -E "passed|failed|beyond|FAILED|Error" $O/${TAG}_tests.log | tail -25
bash tools/gpu_sweep.sh $TAG "OHAO_B200_LIB=ab/lib_head.so;OHB_X=fastdiv;OHAO_B200_LIB=ab/lib_head.so;OHB_X=fastdiv" "helmet cornell synthetic2m"

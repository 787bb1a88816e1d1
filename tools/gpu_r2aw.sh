# r2aw: evaluation-class divisions (BSDF / pdf / MIS arithmetic) as MUFU.RCP + multiply, geometry-deciding divisions left IEEE
O=gpurun_out; TAG=r2aw
( timeout 1800 python -m pytest tests -m gpu -q -s ) > $O/${TAG}_tests.log 2>&1; grep -E "passed|failed|beyond|FAILED" $O/${TAG}_tests.log | grep -v "print\|of the samples" | tail -14
bash tools/gpu_sweep.sh $TAG "OHAO_B200_LIB=ab/lib_head.so;OHB_X=ediv;OHAO_B200_LIB=ab/lib_head.so;OHB_X=ediv" "helmet cornell synthetic2m"

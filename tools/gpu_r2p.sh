# r2p: parity diagnosis on the 2 M scene, real-asset GPU parity test + renders
set -x
O=gpurun_out; TAG=r2p
timeout 600 python tools/diag_parity_2m.py > $O/${TAG}_diag_parity_2m.txt 2>&1
cat $O/${TAG}_diag_parity_2m.txt
( timeout 600 python -m pytest tests/test_assets.py tests/test_host_cpp.py -m gpu -x -q -s ) 2>&1 | tail -8
timeout 900 python tools/render_real_assets.py 256 64 $O > $O/${TAG}_real_assets.log 2>&1
tail -40 $O/${TAG}_real_assets.log

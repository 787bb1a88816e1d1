# r2ah: full GPU suite + smoke after the r2ae..r2ag kernel changes
O=gpurun_out; TAG=r2ah
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_tests.log 2>&1; tail -6 $O/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log

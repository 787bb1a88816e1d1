# r2ap: two GPUs after the two-lane / kernel changes: the driver's weak-scaling command, strong scaling, the sharded 4K image check
O=gpurun_out; TAG=r2ap
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 6 --warmup 3 > $O/${TAG}_weak_n2.json 2> $O/${TAG}_weak_n2.err; tail -c 300 $O/${TAG}_weak_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --scaling strong --steps 6 --warmup 3 > $O/${TAG}_strong_n2.json 2> $O/${TAG}_strong_n2.err; tail -c 300 $O/${TAG}_strong_n2.err
timeout 300 python bench.py --scaling strong --steps 6 --warmup 3 --no-cpu-baseline > $O/${TAG}_strong_n1.json 2> $O/${TAG}_strong_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29554 tools/converged_4k.py --spp 256 --check > $O/${TAG}_config4_n2.json 2> $O/${TAG}_config4_n2.err; tail -c 400 $O/${TAG}_config4_n2.json
python - <<PY
import json
for n in ("weak_n2", "strong_n1", "strong_n2"):
    try: d = json.loads([l for l in open("gpurun_out/r2ap_%s.json" % n) if l.startswith("{")][0])
    except Exception as e: print(n, "missing", e); continue
    print("%s: %.1f Msamples/s  %.2f ms/step  e2e %.1f" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY

# r2aq: N GPUs of one box (default 8) with the end-of-round kernels: the driver's weak-scaling command, strong scaling (fixed 256-spp image),
# config 4 (3840x2160 x 1024 spp sharded, checked against one GPU), config 5 probes
N=${1:-8}; O=gpurun_out; TAG=r2aq
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 6 --warmup 3 > $O/${TAG}_weak_n$N.json 2> $O/${TAG}_weak_n$N.err; tail -c 200 $O/${TAG}_weak_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N --scaling strong --steps 6 --warmup 3 > $O/${TAG}_strong_n$N.json 2> $O/${TAG}_strong_n$N.err; tail -c 200 $O/${TAG}_strong_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29564 tools/converged_4k.py --spp 1024 > $O/${TAG}_config4_n$N.json 2> $O/${TAG}_config4_n$N.err; tail -c 300 $O/${TAG}_config4_n$N.json
python - <<PY
import json
for n in ("weak_n$N", "strong_n$N"):
    try: d = json.loads([l for l in open("gpurun_out/r2aq_%s.json" % n) if l.startswith("{")][0])
    except Exception as e: print(n, "missing", e); continue
    print("%s: %.1f Msamples/s  %.2f ms/step  e2e %.1f" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY

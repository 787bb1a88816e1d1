# Configs 4 and 5 of BASELINE.json on N GPUs of one box.  usage: bash tools/gpu_configs45.sh <tag> <N>
TAG=$1; N=$2; O=gpurun_out; set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/converged_4k.py --spp 1024 --check > $O/${TAG}_config4_n$N.json 2> $O/${TAG}_config4_n$N.err
tail -c 700 $O/${TAG}_config4_n$N.json; tail -3 $O/${TAG}_config4_n$N.err
timeout 300 python tools/converged_4k.py --spp 256 > $O/${TAG}_config4_n1.json 2> $O/${TAG}_config4_n1.err; tail -c 400 $O/${TAG}_config4_n1.json
make -C host all > /dev/null 2>&1
( time timeout 600 host/inverse_fit --backend pt --preset lantern --quality draft --iters 6 --gpus $N --schedule flat ) > $O/${TAG}_config5_n$N.log 2>&1; tail -8 $O/${TAG}_config5_n$N.log
( time timeout 600 host/inverse_fit --backend pt --preset lantern --quality draft --iters 6 --gpus 1 ) > $O/${TAG}_config5_n1.log 2>&1; tail -8 $O/${TAG}_config5_n1.log
# realtime config with the reworked a-trous weights, with and without the SVGF denoiser; realtime + SVGF GPU tests
timeout 300 python bench.py --workload synthetic2m --integrator realtime --steps 60 --no-cpu-baseline > $O/${TAG}_bench_s2m_rt.json 2> $O/${TAG}_bench_s2m_rt.err
timeout 300 python bench.py --workload synthetic2m --integrator realtime --steps 60 --no-cpu-baseline --denoise svgf > $O/${TAG}_bench_s2m_rt_svgf.json 2> $O/${TAG}_bench_s2m_rt_svgf.err
TAG=$TAG python - <<'PY'
import json, os
for n in ("s2m_rt", "s2m_rt_svgf"):
    d = json.load(open("gpurun_out/%s_bench_%s.json" % (os.environ["TAG"], n))); print(n, round(d["value"], 1), round(d["config"]["frames_per_s"], 1), {k: round(v["ms"] / d["steps"], 3) for k, v in d["kernels"].items()})
PY
timeout 600 python -m pytest tests/test_gpu_realtime.py tests/test_gpu_svgf.py -x -q 2>&1 | tail -3

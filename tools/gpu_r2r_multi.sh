# r2r: everything multi-GPU on ONE 8-GPU box: strong scaling of configs[1] (fixed 1080p x 256 spp image), configs 4 and 5 at N = 4, 8,
# the C++ host's own ncclReduce (cornell_box --gpus N).   usage: bash tools/gpu_r2r_multi.sh <tag>
TAG=${1:-r2r}; O=gpurun_out; set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -gt $NG ] && continue
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --scaling strong --steps 6 --warmup 3 --no-cpu-baseline > $O/${TAG}_strong_n1.json 2> $O/${TAG}_strong_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+N)) bench.py --gpus $N --scaling strong --steps 6 --warmup 3 > $O/${TAG}_strong_n$N.json 2> $O/${TAG}_strong_n$N.err
  fi
  tail -c 300 $O/${TAG}_strong_n$N.err
done
python - <<PY
import json
base = None
for n in (1, 2, 4, 8):
    try: d = json.loads([l for l in open("gpurun_out/${TAG}_strong_n%d.json" % n) if l.startswith("{")][0])
    except Exception as e: print(n, "missing", e); continue
    base = base or d["value"]
    print("strong N=%d: %.1f Msamples/s  %.2f ms/image  efficiency %.3f  e2e %.1f" % (n, d["value"], d["ms_per_step"], d["value"] / (n * base), d["e2e"]["value"]))
PY
for N in 4 8; do
  [ $N -gt $NG ] && continue
  CHECK=""; [ $N -eq $NG ] && CHECK="--check"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) tools/converged_4k.py --spp 1024 $CHECK > $O/${TAG}_config4_n$N.json 2> $O/${TAG}_config4_n$N.err
  tail -c 500 $O/${TAG}_config4_n$N.json
done
timeout 300 python tools/converged_4k.py --spp 256 > $O/${TAG}_config4_n1.json 2> $O/${TAG}_config4_n1.err; tail -c 300 $O/${TAG}_config4_n1.json
make -C host all > /dev/null 2>&1
for N in 1 4 8; do
  [ $N -gt $NG ] && continue
  ( time timeout 600 host/inverse_fit --backend pt --preset lantern --quality draft --iters 6 --gpus $N --schedule flat ) > $O/${TAG}_config5_n$N.log 2>&1; grep -E "probes/s|real" $O/${TAG}_config5_n$N.log | tail -3
done
for N in 1 2 4 8; do
  [ $N -gt $NG ] && continue
  timeout 300 host/cornell_box $O/${TAG}_cornell_host_n$N.png 256 --gpus $N > $O/${TAG}_cornell_host_n$N.log 2>&1; grep Done $O/${TAG}_cornell_host_n$N.log
done
python - <<PY
import numpy as np
from PIL import Image
try:
    a = np.asarray(Image.open("gpurun_out/${TAG}_cornell_host_n1.png").convert("RGB"), np.int16)
    for n in (2, 4, 8):
        try: b = np.asarray(Image.open("gpurun_out/${TAG}_cornell_host_n%d.png" % n).convert("RGB"), np.int16)
        except Exception: continue
        d = np.abs(a - b); print("cornell_box --gpus %d vs 1: max |diff| %d LSB, pixels differing %.2e" % (n, d.max(), (d.max(-1) > 0).mean()))
except Exception as e: print("compare failed", e)
PY
rm -f $O/${TAG}_cornell_host_n2.png $O/${TAG}_cornell_host_n4.png

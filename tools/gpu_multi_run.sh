# N-GPU pass on one box: the driver's torchrun launch of bench.py, both arms.  usage: bash tools/gpu_multi_run.sh <tag> <N>
TAG=$1; N=$2; O=gpurun_out; set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
tail -c 1500 $O/${TAG}_bench_n$N.json; tail -5 $O/${TAG}_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 --workload synthetic2m > $O/${TAG}_bench_s2m_n$N.json 2> $O/${TAG}_bench_s2m_n$N.err
tail -c 600 $O/${TAG}_bench_s2m_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/${TAG}_bench_ref_n$N.json 2> $O/${TAG}_bench_ref_n$N.err
tail -c 400 $O/${TAG}_bench_ref_n$N.json
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
tail -c 300 $O/${TAG}_bench_n1.json

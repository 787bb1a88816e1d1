set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1_tests.log 2>&1
tail -5 gpurun_out/r1_tests.log
timeout 600 python bench.py > gpurun_out/r1_bench_helmet.json 2> gpurun_out/r1_bench_helmet.err
tail -c 600 gpurun_out/r1_bench_helmet.json
timeout 600 python bench.py --workload synthetic2m --no-cpu-baseline > gpurun_out/r1_bench_s2m.json 2> gpurun_out/r1_bench_s2m.err
timeout 600 python bench.py --workload synthetic2m --integrator realtime --steps 60 --no-cpu-baseline > gpurun_out/r1_bench_s2m_rt.json 2> gpurun_out/r1_bench_s2m_rt.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_helmet.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_(trace|surface|bounce|film|advance|raygen)' -c 56 -o gpurun_out/r1_full_helmet python bench.py --steps 1 --warmup 3 --spp-step 4 --no-cpu-baseline > gpurun_out/r1_ncu_full_helmet.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_(trace|surface|bounce|film|advance|raygen)' -c 56 -o gpurun_out/r1_full_s2m python bench.py --workload synthetic2m --steps 1 --warmup 3 --spp-step 4 --no-cpu-baseline > gpurun_out/r1_ncu_full_s2m.log 2>&1
ls -la gpurun_out

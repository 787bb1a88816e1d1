# r2a: pipe-throughput microbenchmarks, Vulkan/lavapipe probe of the GPU box, baseline numbers of the r1 build on this box.
set -x
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
( nproc; lscpu | head -20 ) > $O/r2a_host.txt 2>&1
{ echo "== which vulkaninfo glslc glslangValidator"; which vulkaninfo glslc glslangValidator cmake;
  echo "== icd dirs"; ls -la /usr/share/vulkan/icd.d /etc/vulkan/icd.d /usr/local/share/vulkan/icd.d 2>&1;
  echo "== libvulkan / lvp"; ldconfig -p | grep -i -E "vulkan|lvp|llvmpipe|mesa" ; find / -xdev \( -name "libvulkan*" -o -name "*lvp*" -o -name "vulkan*.h" -o -name "libOpenImageDenoise*" -o -name "glm" \) 2>/dev/null | head -20;
  echo "== dpkg"; dpkg -l 2>/dev/null | grep -i -E "vulkan|mesa|glslang|shaderc" ;
  echo "== vulkaninfo"; vulkaninfo --summary 2>&1 | head -20; } > $O/r2a_vulkan_probe.txt 2>&1
( cd tools/ubench && ./pipes ) > $O/r2a_pipes.txt 2>&1
cat $O/r2a_pipes.txt
timeout 600 python bench.py --workload synthetic2m --no-cpu-baseline > $O/r2a_bench_s2m.json 2> $O/r2a_bench_s2m.err
timeout 600 python bench.py --workload synthetic2m --integrator realtime --steps 60 --no-cpu-baseline > $O/r2a_bench_s2m_rt.json 2> $O/r2a_bench_s2m_rt.err
cat $O/r2a_bench_s2m.json

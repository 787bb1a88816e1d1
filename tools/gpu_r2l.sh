python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size)
import ctypes
rt = ctypes.CDLL("libcudart.so") if False else None
PY
cat > /tmp/q.cu <<'CU'
#include <cstdio>
#include <cuda_runtime.h>
int main(){cudaDeviceProp p; cudaGetDeviceProperties(&p,0); printf("persistingL2CacheMaxSize %d accessPolicyMaxWindowSize %d l2 %d\n", p.persistingL2CacheMaxSize, p.accessPolicyMaxWindowSize, p.l2CacheSize); return 0;}
CU
nvcc -o /tmp/q /tmp/q.cu && /tmp/q
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) 2>&1 | tail -2
bash tools/gpu_sweep.sh r2l "OHB_TRACE_OCC=7 OHB_L2_PERSIST=0;OHB_TRACE_OCC=7 OHB_L2_PERSIST=1" "synthetic2m helmet"

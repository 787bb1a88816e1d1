#!/usr/bin/env python3
"""Generate tests/golden/ref_vectors.npz from the REFERENCE's own code (run in the build container).

Uses oracle/_ref/libohao_ref.so = the reference's env_cdf.cpp / sobol_generator.cpp /
owen_scramble.cpp compiled unmodified (oracle/Makefile `ref`).  The vectors travel to the GPU box,
the reference checkout does not.  Also copies the reference's golden image
tests/golden/cornell_box.png (test fixture, not source) next to them.
"""
import os, shutil, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O
import ctypes as C

O.build()
ref = O.ref_lib()
assert ref is not None, "oracle/_ref/libohao_ref.so missing (needs /root/reference)"
rng = np.random.default_rng(20261017)
idx = np.concatenate([np.arange(64), rng.integers(0, 2**32, 192, dtype=np.uint64)]).astype(np.uint32)
sob = np.array([[ref.ref_sobol_sample1d(int(i), d) for d in range(4)] for i in idx], np.float32)
ov = rng.integers(0, 2**32, 256, dtype=np.uint64).astype(np.uint32); os_ = rng.integers(0, 2**32, 256, dtype=np.uint64).astype(np.uint32)
owen = np.array([ref.ref_owen(int(v), int(s)) for v, s in zip(ov, os_)], np.uint32)
dirs = np.ctypeslib.as_array(ref.ref_sobol_dirs(), shape=(128,)).copy()

def cdf(img):
    h, w = img.shape[:2]; m = np.zeros(h, np.float32); c = np.zeros((h, w), np.float32); I = C.c_float()
    img = np.ascontiguousarray(img, np.float32)
    ref.ref_env_cdf(img.ctypes.data_as(C.c_void_p), w, h, m.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), C.byref(I))
    return m, c, np.float32(I.value)

env_rand = np.ones((16, 32, 4), np.float32); env_rand[..., :3] = rng.random((16, 32, 3), dtype=np.float32) * 4
env_hot = np.full((16, 32, 4), 0.01, np.float32); env_hot[5, 20, :3] = 1000.0   # env_cdf_test.cpp hot-spot case
env_black = np.zeros((8, 16, 4), np.float32); env_black[3, :, :3] = 1.0          # black rows -> uniform fallback
out = dict(sobol_index=idx, sobol_values=sob, owen_v=ov, owen_seed=os_, owen_out=owen, sobol_dirs=dirs,
           env_rand=env_rand, env_hot=env_hot, env_black=env_black)
for name, img in (("rand", env_rand), ("hot", env_hot), ("black", env_black)):
    m, c, I = cdf(img); out[f"cdf_{name}_marg"] = m; out[f"cdf_{name}_cond"] = c; out[f"cdf_{name}_integral"] = I
np.savez_compressed(os.path.join(ROOT, "tests/golden/ref_vectors.npz"), **out)
src = "/root/reference/tests/golden/cornell_box.png"
if os.path.exists(src):
    shutil.copyfile(src, os.path.join(ROOT, "tests/golden/reference_cornell_box_16spp_640.png"))
print("wrote tests/golden/ref_vectors.npz")

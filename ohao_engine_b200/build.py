"""In-tree build of libohao_b200.so (hand-written CUDA for sm_100a + the C ABI)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libohao_b200.so")
SOURCES = ["ohb_kernels.cu", "ohb_api.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libohao_b200.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in os.listdir(CSRC)) or \
        os.path.getmtime(os.path.join(_HERE, "..", "include", "ohao_b200.h")) > t


def build_native(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *[os.path.join(CSRC, s) for s in SOURCES]]
        if verbose:
            cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))

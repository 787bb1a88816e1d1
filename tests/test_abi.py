"""The C-ABI library loads on a GPU-less box and exports every symbol include/ohao_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from ohao_engine_b200 import binding as B, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "ohao_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ohb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    build.build_native()
    lib = C.CDLL(build.LIB_PATH)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ohao_b200.h but not exported"
    assert set(names) == set(B.ABI.keys())
    assert B.load_library().ohb_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B.OhbError, match="no CUDA device"):
        B.Renderer(64, 64)


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ohao_engine_b200")):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "libemul" not in txt, f

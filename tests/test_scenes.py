import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from ohao_engine_b200 import scenes


def test_cornell_contract(cornell):
    ps, cam = cornell
    assert ps.ntris == 2058 and len(ps.instances) == 7 and ps.nlights == 12 and ps.nmaterials == 7
    assert ps.positions.strides[0] == 100 and ps.light_ssbo.nbytes == 16 + 12 * 80                    # 976 B
    assert ps.env is None and ps.light_ssbo[4:8].tobytes() == b"\xff" * 4
    assert list(ps.instances["first_tri"]) == [0, 2, 4, 6, 8, 10, 1034]
    v = cam.view().reshape(4, 4)
    assert np.allclose(v[3, :3], [0, 0, -13], atol=1e-5)


def test_unordered_map_order_matches_libstdcxx():
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    src = "#include <unordered_map>\n#include <cstdint>\n#include <cstdio>\nint main(){std::unordered_map<uint64_t,int> m;for(uint64_t k=1;k<=20;k++)m[k]=1;for(auto&kv:m)printf(\"%lu \",kv.first);}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "a.cpp"), "w").write(src)
        subprocess.check_call([gxx, "-std=c++17", "-O1", "-o", os.path.join(d, "a"), os.path.join(d, "a.cpp")])
        out = subprocess.check_output([os.path.join(d, "a")]).decode().split()
    assert [int(x) for x in out] == scenes.LIBSTDCXX_ORDER_1_TO_20


def test_solid_colour_layer_quirk_q1(cornell):
    ps, _ = cornell
    # white 0.73 -> sRGB-encoded 222 in a UNORM layer -> albedo = 0.73 * (222/255)^2.2 ~ 0.538
    back = ps.textures[0, 0, 0]
    assert tuple(back[:3]) == (222, 222, 222)


def test_synthetic_scene_sizes():
    ps = scenes.synthetic_2m(nblobs=10, tris_per_blob=2000, env_size=(64, 32))
    assert ps.ntris == 10 * 2000 + 2 and ps.nmaterials == 17 and ps.nlights == 8
    hc = scenes.helmet_class(ntris=50000, tex_size=64, env_size=(64, 32))
    assert abs(hc.ntris - 50002) <= 200 and hc.textures.shape[0] == 5

"""C++20 host side (host/): builds everywhere; on a GPU box the example binaries render through the C ABI and must agree with the
Python binding on the same scene (which in turn is parity-checked against the oracle)."""
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host")


def _build():
    from ohao_engine_b200 import build
    build.build_native()
    subprocess.check_call(["make", "-C", HOST, "all"], stdout=subprocess.DEVNULL)


def test_host_examples_build_and_refuse_to_run_without_a_gpu():
    _build()
    for b in ("cornell_box", "turntable", "inverse_fit"):
        assert os.access(os.path.join(HOST, b), os.X_OK)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([os.path.join(HOST, "cornell_box"), "/tmp/never.png", "1", "--width", "32", "--height", "18"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr            # no CPU fallback in the host path either


@pytest.mark.gpu
def test_cornell_box_binary_matches_python_binding(tmp_path):
    _build()
    from ohao_engine_b200 import binding as B, scenes
    W, H, spp = 320, 180, 8
    out = str(tmp_path / "cornell.png")
    r = subprocess.run([os.path.join(HOST, "cornell_box"), out, str(spp), "--width", str(W), "--height", str(H), "--denoise=none"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Done:" in r.stdout and "Saved" in r.stdout
    cpp = np.asarray(Image.open(out).convert("RGBA"), np.int16)
    ps, cam = scenes.cornell_box(), scenes.cornell_camera()
    rr = B.Renderer(W, H); rr.set_scene(ps)
    for _ in range(spp):
        rr.render(cam.view(), cam.proj(W, H), 1)
    py = rr.get_pixels().astype(np.int16)
    d = np.abs(cpp - py).max(-1)
    assert (d > 1).mean() < 0.01, ((d > 1).mean(), d.max())             # same packer, same ABI; libm-vs-numpy sin/cos may move the camera by an ulp


@pytest.mark.gpu
def test_turntable_and_inverse_fit_binaries(tmp_path):
    _build()
    out = str(tmp_path / "tt")
    for mode, extra in (("env", []), ("cornell", ["--realtime"]), ("dark", ["--two-level", "--bob"])):
        r = subprocess.run([os.path.join(HOST, "turntable"), "blob:5000", mode, "4", "3", "--width", "160", "--height", "90", "--outdir", out] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        frames = sorted(f for f in os.listdir(out) if f.startswith(mode))
        assert len(frames) == 3
        im = np.asarray(Image.open(os.path.join(out, frames[1])).convert("RGB"))
        assert im.shape == (90, 160, 3) and im.mean() > 2
    r = subprocess.run([os.path.join(HOST, "inverse_fit"), "--backend", "pt", "--preset", "lantern", "--quality", "draft", "--iters", "6", "--tris", "3000", "--schedule", "flat"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr                         # exit 0 <=> the fit reduced the loss
    assert "probes/s" in r.stdout


def test_cpp_host_ingests_the_real_assets_like_the_python_packer(tmp_path):
    """host/asset_io.hpp (loadGLB + side-car textures, loadHDR) and the C++ packer against ohao_engine_b200/assets.py on
    DamagedHelmet.glb: `turntable <glb> env --dump-scene` writes the packed §3.2 arrays without needing a GPU."""
    import shutil, struct
    from ohao_engine_b200 import assets
    _build()
    glb = str(tmp_path / "DamagedHelmet.glb"); shutil.copy(os.path.join(ROOT, "tests", "assets", "DamagedHelmet.glb"), glb)
    assets.bake_textures(glb)
    dump = str(tmp_path / "scene.bin")
    r = subprocess.run([os.path.join(HOST, "turntable"), glb, "env", "4", "1", "--hdr", os.path.join(ROOT, "tests", "assets", "env_outdoor.hdr"), "--dump-scene", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(dump, "rb").read(); assert raw[:4] == b"OHBS"
    off, rec = 4, {}
    while off < len(raw):
        tag, n = struct.unpack_from("<IQ", raw, off); off += 12; rec[tag] = raw[off:off + n]; off += n
    ps = assets.turntable_env_scene(glb, os.path.join(ROOT, "tests", "assets", "env_outdoor.hdr"))
    assert np.array_equal(np.frombuffer(rec[1], np.float32).reshape(-1, 25)[:, :3], ps.positions[:, :3])
    assert np.array_equal(np.frombuffer(rec[2], np.uint32), ps.indices) and np.array_equal(np.frombuffer(rec[3], np.uint32), ps.mat_ids)
    assert np.array_equal(np.frombuffer(rec[4], np.float32).reshape(-1, 4), ps.normals) and np.array_equal(np.frombuffer(rec[5], np.float32).reshape(-1, 2), ps.uvs)
    assert np.array_equal(np.frombuffer(rec[6], np.uint32), np.asarray(ps.mat_colors, np.float32).view(np.uint32).reshape(-1))      # incl. the texture-mean base colour
    assert np.frombuffer(rec[7], ps.instances.dtype).tobytes() == ps.instances.tobytes()                                            # T * R_y(180) * S
    assert rec[8] == ps.light_ssbo.tobytes()                                                                                        # key light + emissive-mesh auto light + env index
    assert tuple(np.frombuffer(rec[9], np.uint32)) == (2048, 2048, 4) and np.array_equal(np.frombuffer(rec[10], np.uint8).reshape(ps.textures.shape), ps.textures)


@pytest.mark.gpu
def test_inverse_fit_staged_schedule_reduces_the_loss():
    """StagedFitter restated (staged_fit.hpp): multi-start probe, the env -> lights -> albedo -> brdf -> ... -> refine stage list with
    per-stage Adam, hybridSpecularRGB + regularisers; the probes of every stage fan out over the GPUs."""
    _build()
    r = subprocess.run([os.path.join(HOST, "inverse_fit"), "--backend", "pt", "--preset", "lantern", "--quality", "draft", "--iters", "20", "--tris", "3000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr                  # exit 0 <=> final loss < initial loss
    out = r.stdout
    assert "candidate 5/5" in out and all(f"-- stage {s} " in out for s in ("env", "lights", "albedo", "brdf", "brdf2", "pedestal", "lights2", "refine"))
    assert "probes/s" in out and "final loss" in out

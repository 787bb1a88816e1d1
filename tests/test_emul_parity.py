"""Product per-thread code (LBVH builder, traversal, wavefront integrator — the same headers nvcc
compiles for sm_100a) driven by the host emulator, checked against the oracle.  CPU-only."""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests import util
from tests.emul import emul_py as E


@pytest.mark.parametrize("which", ["cornell", "helmet_small", "synthetic_small"])
def test_closest_hit_prim_ids_bit_exact(which, request):
    ps, _ = request.getfixturevalue(which)
    osc, esc = O.OracleScene(ps), E.EmulScene(ps)
    lo, hi = (-4.9, 4.9) if which == "cornell" else ((-3, 3) if which == "helmet_small" else (-60, 60))
    rays = util.random_rays(60000, lo, hi, seed=11)
    ref, got = osc.trace(rays), esc.trace(rays)
    mism, ties = util.compare_hits(ref, got)
    assert mism == 0, (mism, ties)
    assert np.array_equal(ref["t"], got["t"]) and np.array_equal(ref["u"], got["u"]) and np.array_equal(ref["v"], got["v"])
    assert np.array_equal(osc.occluded(rays[:20000]), esc.occluded(rays[:20000]))


def test_oracle_bvh_equals_brute_force(cornell):
    ps, _ = cornell
    osc = O.OracleScene(ps)
    rays = util.random_rays(20000, -4.9, 4.9, seed=5)
    a, b = osc.trace(rays), osc.trace(rays, brute=True)
    assert np.array_equal(a["prim"], b["prim"]) and np.array_equal(a["t"], b["t"])


@pytest.mark.parametrize("which,res,spp", [("cornell", (96, 54), 4), ("helmet_small", (80, 45), 3), ("synthetic_small", (80, 45), 3)])
def test_offline_samples_match_oracle(which, res, spp, request):
    ps, cam = request.getfixturevalue(which)
    W, H = res
    osc, esc = O.OracleScene(ps), E.EmulScene(ps)
    ro = osc.render_offline(cam.view(), cam.proj(W, H), W, H, spp, dump=True, want_aov=True)
    re = esc.render_offline(cam.view(), cam.proj(W, H), W, H, spp, dump=True)
    bad, worst = util.sample_parity(ro["samples"], re["samples"])
    assert bad < 2e-3 and worst < 2e-3, (bad, worst)
    for k in ("closest_rays", "shadow_rays"):
        assert abs(ro["counters"][k] - re["counters"][k]) <= max(8, ro["counters"][k] // 2000), k
    assert np.abs(ro["albedo"] - re["albedo"]).max() < 1e-4 and np.abs(ro["normal"] - re["normal"]).max() < 1e-4


def test_accumulation_continues_across_calls(cornell):
    ps, cam = cornell
    W, H = 64, 36
    esc = E.EmulScene(ps)
    full = esc.render_offline(cam.view(), cam.proj(W, H), W, H, 4)
    part = esc.render_offline(cam.view(), cam.proj(W, H), W, H, 2)
    part = esc.render_offline(cam.view(), cam.proj(W, H), W, H, 2, first_sample=2, history=2, accum=part["accum"])
    assert np.array_equal(full["accum"], part["accum"]) and np.array_equal(full["ldr"], part["ldr"])


def test_tile_rendering_matches_full_frame(cornell):
    ps, cam = cornell
    W, H = 64, 36
    esc = E.EmulScene(ps)
    full = esc.render_offline(cam.view(), cam.proj(W, H), W, H, 2)
    acc = np.zeros((H, W, 4), np.float32)
    for (x0, y0, w, h) in [(0, 0, 37, 19), (37, 0, 27, 19), (0, 19, 64, 17)]:      # ragged tiles (not multiples of 8x4)
        acc = esc.render_offline(cam.view(), cam.proj(W, H), W, H, 2, accum=acc, tile=(x0, y0, w, h))["accum"]
    assert np.array_equal(full["accum"], acc)


def test_env_sampling_matches_oracle(helmet_small):
    ps, _ = helmet_small
    osc, esc = O.OracleScene(ps), E.EmulScene(ps)
    u = np.random.default_rng(4).random((5000, 2), dtype=np.float32)
    a, pa = osc.env_sample(u); b, pb = esc.env_sample(u)
    assert np.allclose(a, b, rtol=1e-6, atol=1e-7) and np.allclose(pa, pb, rtol=1e-6, atol=1e-12)

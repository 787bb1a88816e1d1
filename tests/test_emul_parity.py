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


def test_pcg_sampler_samples_match_oracle(cornell):
    """SAMPLER_PCG (sampler_pcg.glsl:9-29) — dead in the reference's pipelines (quirk Q3), still part of the sampler API."""
    ps, cam = cornell
    W, H, spp = 64, 36, 3
    osc, esc = O.OracleScene(ps), E.EmulScene(ps)
    ro = osc.render_offline(cam.view(), cam.proj(W, H), W, H, spp, dump=True, settings=O.offline_settings(sampler=0))
    re = esc.render_offline(cam.view(), cam.proj(W, H), W, H, spp, dump=True, settings=O.offline_settings(sampler=0))
    bad, worst = util.sample_parity(ro["samples"], re["samples"])
    assert bad < 2e-3 and worst < 2e-3, (bad, worst)
    rs = osc.render_offline(cam.view(), cam.proj(W, H), W, H, spp, dump=True)
    assert util.sample_parity(rs["samples"], re["samples"])[0] > 0.5      # and it is not the Sobol sequence


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


# ---- compressed 8-wide BVH: builder edge cases and the paths a single host thread would not take ------------------
def _soup(tris, name="soup"):
    """PackedScene of one mesh made of independent triangles (n,3,3)."""
    from ohao_engine_b200 import scenes as S
    tris = np.asarray(tris, np.float32)
    n = len(tris)
    m = S.Mesh(positions=tris.reshape(-1, 3), normals=np.tile(np.array([0, 1, 0], np.float32), (3 * n, 1)), uvs=np.zeros((3 * n, 2), np.float32),
               indices=np.arange(3 * n, dtype=np.uint32))
    return S.pack_scene([m], [S.Light(position=(0, 5, 0))], name=name)


def _assert_same_hits(ps, rays):
    osc, esc = O.OracleScene(ps), E.EmulScene(ps)
    ref, got = osc.trace(rays), esc.trace(rays)
    for k in ("prim", "t", "u", "v"):
        assert np.array_equal(ref[k], got[k]), k
    assert np.array_equal(osc.occluded(rays), esc.occluded(rays))
    return esc, ref


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9, 25])
def test_wide_bvh_tiny_scenes(n):
    """Fewer triangles than one wide node holds: the root is the only node, or has a single inner child."""
    rng = np.random.default_rng(n)
    tris = rng.uniform(-1, 1, (n, 3, 3))
    esc, ref = _assert_same_hits(_soup(tris), util.random_rays(4000, -2, 2, seed=n))
    assert (ref["prim"] != 0xFFFFFFFF).any()
    assert esc.levels() <= 3


def test_wide_bvh_coincident_and_flat_geometry():
    """Many identical triangles (equal Morton codes, equal-t ties -> lowest id wins) on an axis-aligned plane
    (zero grid extent on one axis)."""
    base = np.array([[-1, 0, -1], [1, 0, -1], [0, 0, 1]], np.float32)
    tris = np.concatenate([np.tile(base, (40, 1, 1)), np.tile(base + np.array([0.5, 0, 0.25], np.float32), (23, 1, 1))])
    rays = util.random_rays(6000, -1.5, 1.5, seed=2)
    esc, ref = _assert_same_hits(_soup(tris), rays)
    hit = ref["prim"] != 0xFFFFFFFF
    assert hit.sum() > 500 and set(np.unique(ref["prim"][hit])) <= {0, 40}      # ties resolve to the lowest id of each stack


def test_wide_bvh_far_from_origin_and_mixed_scales():
    """Grid origins around 1e4 with millimetre triangles next to a 100 m quad: the error bound of the quantised slab
    test scales with |(p - o)/d|, so nothing may be culled wrongly."""
    rng = np.random.default_rng(9)
    c = np.array([1.0e4, -2.0e3, 5.0e3], np.float32)
    small = c + rng.uniform(-0.5, 0.5, (300, 1, 3)) + rng.uniform(-2e-3, 2e-3, (300, 3, 3))
    big = c + np.array([[[-50, -1, -50], [50, -1, -50], [0, -1, 60]]], np.float32)
    tris = np.concatenate([small, big]).astype(np.float32)
    rays = util.random_rays(8000, -1.0, 1.0, seed=4)
    rays["origin"] += c
    k = 3000                                    # aim a share of the rays straight at the small triangles
    tgt = small[rng.integers(0, 300, k)].mean(1)
    d = tgt - rays["origin"][:k]; rays["dir"][:k] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    esc, ref = _assert_same_hits(_soup(tris), rays)
    assert (ref["prim"][:k] < 300).sum() > 1000


def test_wide_bvh_postponing_and_pause_paths(synthetic_small):
    """Pseudo-random active-lane counts drive triangle postponing (stack pushes of triangle groups) in the emulator;
    results must not change."""
    ps, _ = synthetic_small
    esc = E.EmulScene(ps)
    rays = util.random_rays(30000, -60, 60, seed=21)
    a, oa = esc.trace(rays), esc.occluded(rays)
    E.set_warp_noise(True)
    try:
        b, ob = esc.trace(rays), esc.occluded(rays)
    finally:
        E.set_warp_noise(False)
    for k in ("prim", "t", "u", "v"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(oa, ob)


def test_wide_bvh_stats_and_front_to_back_order(helmet_small):
    ps, cam = helmet_small
    esc = E.EmulScene(ps)
    nodes, sah = esc.stats()
    assert 0 < nodes < ps.ntris / 3 and esc.levels() <= 12 and 1.0 < sah < 40.0
    # the octant order must pay off: closest-hit queries visit clearly fewer nodes than any-hit-free full enumeration would
    rays = util.random_rays(20000, -3, 3, seed=8)
    E.trav_stats()
    esc.trace(rays)
    n, t = E.trav_stats()
    assert n / len(rays) < 8.0 and t / len(rays) < 6.0, (n / len(rays), t / len(rays))

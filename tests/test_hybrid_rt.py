"""Hybrid-RT techniques: the soft-shadow and one-bounce-GI raygens of the deferred path
(shaders/rt/rt_shadow.rgen, rt_gi.rgen + rt_gi.rchit) on a G-buffer, against their CPU restatement.  The G-buffer a raster
pass would produce is synthesised here from the oracle's primary hits (world position, octahedron-encoded shading normal,
albedo)."""
import numpy as np
import pytest

from ohao_engine_b200 import scenes
from oracle import oracle_py as O


def _gbuffer(ps, cam, W, H):
    osc = O.OracleScene(ps)
    v = np.asarray(cam.view(), np.float64).reshape(4, 4).T; p = np.asarray(cam.proj(W, H), np.float64).reshape(4, 4).T
    iv, ip = np.linalg.inv(v), np.linalg.inv(p)
    ys, xs = np.mgrid[0:H, 0:W]
    ndc = np.stack([(xs + 0.5) / W * 2 - 1, (ys + 0.5) / H * 2 - 1, np.ones((H, W)), np.ones((H, W))], -1).reshape(-1, 4)
    t = (ip @ ndc.T).T; t = t[:, :3] / t[:, 3:4]
    d = (iv[:3, :3] @ (t / np.linalg.norm(t, axis=1, keepdims=True)).T).T
    rays = np.zeros(W * H, O.RAY_DTYPE); rays["origin"] = iv[:3, 3]; rays["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["tmin"] = 1e-3; rays["tmax"] = 1e4
    hit = osc.trace(rays)
    ok = hit["prim"] != 0xFFFFFFFF
    prim = np.where(ok, hit["prim"], 0)
    pos = (rays["origin"] + rays["dir"] * hit["t"][:, None]).astype(np.float32)
    idx = ps.indices.reshape(-1, 3)[prim]
    w = (1 - hit["u"] - hit["v"])[:, None]
    n = w * ps.normals[idx[:, 0], :3] + hit["u"][:, None] * ps.normals[idx[:, 1], :3] + hit["v"][:, None] * ps.normals[idx[:, 2], :3]
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-12)
    n = np.where((n * rays["dir"]).sum(1, keepdims=True) > 0, -n, n)                         # face the camera, like a raster G-buffer
    l1 = np.abs(n).sum(1, keepdims=True); o = n[:, :2] / l1
    fold = (1 - np.abs(o[:, ::-1])) * np.where(o >= 0, 1.0, -1.0)
    o = np.where(n[:, 2:3] < 0, fold, o)                                                     # encodeNormalOctahedron (encoding.glsl:25-32)
    gpos = np.zeros((H, W, 4), np.float32); gpos.reshape(-1, 4)[ok, :3] = pos[ok]; gpos.reshape(-1, 4)[ok, 3] = 1
    gnrm = (o * 0.5 + 0.5).astype(np.float32).reshape(H, W, 2)
    mc = np.asarray(ps.mat_colors, np.float32).reshape(-1, 3, 4)
    alb = np.zeros((H, W, 4), np.float32); alb.reshape(-1, 4)[ok, :3] = mc[ps.mat_ids[prim[ok]], 0, :3]; alb[..., 3] = 1
    inst_of_tri = np.zeros(ps.ntris, np.int64)
    for i, ins in enumerate(ps.instances): inst_of_tri[ins["first_tri"]:ins["first_tri"] + ins["tri_count"]] = i
    inst_mat = np.ones((len(ps.instances), 4), np.float32)
    for i, ins in enumerate(ps.instances): inst_mat[i, :3] = mc[ps.mat_ids[ins["first_tri"]], 0, :3]
    if len(ps.instances) > 2: inst_mat[-1, 3] = 0.0                                          # an "animated" instance: GI rays treat it as a miss
    return osc, gpos, gnrm, alb, inst_mat, ok.reshape(H, W)


def _params():
    sp = O.HybridShadowParams(); sp.light_dir[:] = [0.3, -1.0, 0.2]; sp.light_radius = 0.05; sp.light_pos[:] = [0.0, 4.0, 0.0]; sp.light_range = 0.0; sp.light_type = 0; sp.sample_count = 4
    pp = O.HybridShadowParams(); pp.light_dir[:] = [0, -1, 0]; pp.light_radius = 0.3; pp.light_pos[:] = [2.0, 3.0, 2.5]; pp.light_range = 50.0; pp.light_type = 1; pp.sample_count = 6
    gp = O.HybridGiParams(); gp.light_pos[:] = [2.0, 3.0, 2.5]; gp.light_intensity = 40.0; gp.sample_count = 4; gp.frame_index = 0
    return sp, pp, gp


def _h2f(bits): return bits.view(np.float16).astype(np.float32)


def _check(osc, other, W, H, gpos, gnrm, alb, inst_mat, covered, call_shadow, call_gi, exact):
    sp, pp, gp = _params()
    for prm in (sp, pp):
        ref, got = osc.hybrid_shadow(W, H, gpos, gnrm, prm), call_shadow(prm)
        assert (~covered).any() and (ref[~covered] == 255).all() and 0.03 < (ref[covered] < 128).mean() < 0.97   # sky, lit and shadowed surfaces
        if exact: assert np.array_equal(ref, got)
        else: assert (ref != got).mean() < 5e-3                                                          # sin/cos/acos rounding moves a ray across a silhouette
    hist = np.zeros((H, W, 4), np.float32)
    ref0, got0 = osc.hybrid_gi(W, H, gpos, gnrm, alb, hist, inst_mat, gp), call_gi(hist, gp, inst_mat)
    a, b = _h2f(ref0), _h2f(got0)
    assert a[covered][:, :3].max() > 0.01 and (a[~covered] == 0).all()
    tol = dict(rtol=0, atol=0) if exact else dict(rtol=2e-3, atol=2e-4)
    assert np.allclose(a, b, **tol) if exact else (~np.isclose(a, b, **tol)).mean() < 5e-3
    gp.frame_index = 3; hist2 = a.copy()                                                                 # temporal blend: mix(history, gi, 0.3)
    ref1, got1 = osc.hybrid_gi(W, H, gpos, gnrm, alb, hist2, inst_mat, gp), call_gi(hist2, gp, inst_mat)
    assert np.allclose(_h2f(ref1), _h2f(got1), **tol) if exact else (~np.isclose(_h2f(ref1), _h2f(got1), **tol)).mean() < 5e-3
    assert not np.array_equal(ref0, ref1)
    im2 = inst_mat.copy(); im2[0, 3] = 0.0; gp.frame_index = 0                                          # instance 0 "animated": its hits count as misses (rt_gi.rchit:20-25)
    ref2, got2 = osc.hybrid_gi(W, H, gpos, gnrm, alb, hist, im2, gp), call_gi(hist, gp, im2)
    assert np.allclose(_h2f(ref2), _h2f(got2), **tol) if exact else (~np.isclose(_h2f(ref2), _h2f(got2), **tol)).mean() < 5e-3
    assert _h2f(ref2)[..., :3].sum() < 0.98 * a[..., :3].sum()


def test_hybrid_techniques_product_code_matches_oracle_on_the_host(helmet_small):
    from tests.emul import emul_py as E
    ps, cam = helmet_small
    W, H = 96, 54
    osc, gpos, gnrm, alb, inst_mat, covered = _gbuffer(ps, cam, W, H)
    esc = E.EmulScene(ps)
    _check(osc, esc, W, H, gpos, gnrm, alb, inst_mat, covered, lambda p: esc.hybrid_shadow(W, H, gpos, gnrm, p),
           lambda hist, p, im: esc.hybrid_gi(W, H, gpos, gnrm, alb, hist, im, p), exact=True)


@pytest.mark.gpu
def test_hybrid_techniques_gpu_matches_oracle(helmet_small):
    from ohao_engine_b200 import binding as B
    ps, cam = helmet_small
    W, H = 480, 270
    osc, gpos, gnrm, alb, inst_mat, covered = _gbuffer(ps, cam, W, H)
    r = B.Renderer(W, H); r.set_scene(ps)
    _check(osc, r, W, H, gpos, gnrm, alb, inst_mat, covered, lambda p: r.hybrid_shadow(gpos, gnrm, p),
           lambda hist, p, im: r.hybrid_gi(gpos, gnrm, alb, hist, im, p), exact=False)

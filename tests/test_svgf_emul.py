"""SVGF denoiser of DenoiseMode::Atrous: the product's per-pixel code (ohb_svgf.h, host build driven by the
emulator) against the oracle's restatement of rt_svgf_temporal.comp / rt_svgf_atrous.comp / atrous_denoise.cpp.  CPU-only."""
import copy

import numpy as np
import pytest

from oracle import oracle_py as O
from tests.emul import emul_py as E


def test_fp16_conversions_match_ieee():
    """Host f2h / h2f of ohb_svgf.h and the oracle's _Float16 path against numpy's float16 (round-to-nearest-even)."""
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.normal(size=4000) * 10.0 ** rng.integers(-9, 6, 4000), [0.0, -0.0, 1.0, 65504.0, 65519.9, 65520.0, 1e30, -1e30, 5.96e-8, 2.98e-8,
                           2.9802322e-8, 2.9802326e-8, 6.1e-5, 6.097e-5, 0.1, 1.0 / 3.0, np.inf, -np.inf]]).astype(np.float32)
    # ties: exactly half-way between two fp16 values
    h = rng.integers(0, 0x7BFF, 2000).astype(np.uint16)
    mid = ((h.view(np.float16).astype(np.float64) + (h + 1).astype(np.uint16).view(np.float16).astype(np.float64)) / 2).astype(np.float32)
    vals = np.concatenate([vals, mid, -mid])
    ref = vals.astype(np.float16).view(np.uint16)
    el, ol = E.lib(), O.lib()
    got_e = np.array([el.emul_f2h(float(v)) for v in vals], np.uint16); got_o = np.array([ol.orc_f2h(float(v)) for v in vals], np.uint16)
    assert np.array_equal(got_e, ref) and np.array_equal(got_o, ref)
    allh = np.arange(0, 65536, 7, dtype=np.uint16); allh = allh[~np.isnan(allh.view(np.float16))]
    back = np.array([el.emul_h2f(int(x)) for x in allh], np.float32)
    assert np.array_equal(back.view(np.uint32), allh.view(np.float16).astype(np.float32).view(np.uint32))


def _synthetic_inputs(W, H, frame, rng):
    """A moving two-plane scene with noise: enough to exercise reprojection, disocclusion, the bootstrap variance and all edge stops."""
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    split = W * 0.55 + 3.0 * frame
    left = xx < split
    depth = np.where(left, 5.0 + 0.01 * yy, 9.0 + 0.02 * xx).astype(np.float32)
    depth[: H // 6] = 1e30                                                    # sky
    n = np.zeros((H, W, 4), np.float32); n[..., :3] = np.where(left[..., None], [0.0, 0.6, 0.8], [0.7071, 0.0, 0.7071]); n[: H // 6, :, :3] = 0.0
    normal = n * 0.5 + 0.5; normal[..., 3] = 1.0; normal[: H // 6] = 0.0
    base = np.where(left[..., None], [0.55, 0.35, 0.2], [0.2, 0.4, 0.6]).astype(np.float32) * (0.6 + 0.4 * np.sin(xx / 7.0)[..., None] ** 2)
    base[: H // 6] = [0.5, 0.7, 0.9]
    noisy = np.clip(base + rng.normal(scale=0.12, size=base.shape), 0, 1)
    beauty = np.zeros((H, W, 4), np.uint8); beauty[..., :3] = np.rint(noisy * 255); beauty[..., 3] = 255
    mv = np.zeros((H, W, 2), np.float16); mv[..., 0] = np.where(left, 3.0, 1.25); mv[..., 1] = 0.5 * (frame % 2)
    if frame == 0: mv[:] = 0
    motion = mv[..., 0].view(np.uint16).astype(np.uint32) | (mv[..., 1].view(np.uint16).astype(np.uint32) << 16)
    return beauty, normal, depth, motion


def _half(x): return x.view(np.float16).astype(np.float32)


def test_svgf_dispatch_matches_oracle_over_frames():
    W, H = 96, 54
    rng = np.random.default_rng(11)
    so, se = O.SvgfState(W, H), O.SvgfState(W, H)
    for f in range(6):
        beauty, normal, depth, motion = _synthetic_inputs(W, H, f, rng)
        reset = f == 0 or f == 4                                             # explicit resetAccumulation mid-sequence
        do = O.svgf_dispatch(so, beauty, motion, depth, normal, reset)
        de = E.svgf_dispatch(se, beauty, motion, depth, normal, reset)
        # same libm on both sides: only fp-association differences (operator/ by reciprocal in ohb_common.h) remain
        assert (np.abs(do.astype(np.int16) - de.astype(np.int16)).max(-1) > 1).mean() < 1e-3, f
        assert (do != de).any(-1).mean() < 0.02, f
        for name in ("color", "moments", "geom"):
            a, b = _half(getattr(so, name)[so.cur]), _half(getattr(se, name)[se.cur])
            rel = np.abs(a - b) / (np.abs(a) + 1e-3)
            assert (rel > 2e-3).mean() < 2e-3, (f, name, float((rel > 2e-3).mean()))
        # the filter does something: it lowers the noise of the flat regions and keeps the depth / normal edge
        if f >= 2 and not reset:
            flat = (slice(H // 3, H - 4), slice(4, W // 3))
            assert de[flat][..., :3].astype(np.float32).std(axis=(0, 1)).mean() < 0.6 * beauty[flat][..., :3].astype(np.float32).std(axis=(0, 1)).mean()
    # history length grows up to the cap and resets where the two planes disocclude
    ln = _half(se.moments[se.cur])[..., 2]
    assert ln.max() >= 2.0 and ln.min() == 1.0 and ln.max() <= 32.0


def test_svgf_reset_and_out_of_bounds_reprojection():
    W, H = 48, 32
    rng = np.random.default_rng(5)
    beauty, normal, depth, motion = _synthetic_inputs(W, H, 1, rng)
    st = O.SvgfState(W, H)
    E.svgf_dispatch(st, beauty, motion, depth, normal, True)
    assert (_half(st.moments[st.cur])[..., 2] == 1.0).all()                   # reset: every history length is 1
    big = np.zeros((H, W, 2), np.float16); big[..., 0] = -200.0               # reprojects outside the frame everywhere
    mo = big[..., 0].view(np.uint16).astype(np.uint32) | (big[..., 1].view(np.uint16).astype(np.uint32) << 16)
    so = copy.deepcopy(st)
    de = E.svgf_dispatch(st, beauty, mo, depth, normal, False); do = O.svgf_dispatch(so, beauty, mo, depth, normal, False)
    assert (_half(st.moments[st.cur])[..., 2] == 1.0).all() and (np.abs(de.astype(np.int16) - do.astype(np.int16)) <= 1).all()


@pytest.mark.parametrize("which", ["cornell", "helmet_small"])
def test_guides_and_fresh_sample_pipeline(which, request):
    """Integrated path: realtime frames in fresh-sample mode -> guide AOVs -> SVGF, emulator vs oracle, with a moving camera."""
    ps, cam = request.getfixturevalue(which)
    osc, esc = O.OracleScene(ps), E.EmulScene(ps)
    W, H = 64, 36
    so, se = O.RealtimeState(W, H), O.RealtimeState(W, H)
    vo, ve = O.SvgfState(W, H), O.SvgfState(W, H)
    st = O.realtime_settings()
    for f in range(4):
        c = copy.deepcopy(cam); c.yaw = cam.yaw + 0.4 * f
        view, proj = c.view(), c.proj(W, H)
        pvo, pve, fi = so.prev_view_proj.copy(), se.prev_view_proj.copy(), so.frame_index
        ro = osc.render_realtime(so, view, proj, settings=st, fresh=True)
        re = esc.render_realtime(se, view, proj, settings=st, fresh=True)
        mo, dpo = O.svgf_guides(ro["surf"], view, proj, pvo, fi)
        me, dpe = E.svgf_guides(re["surf"], view, proj, pve, fi)
        hit = ro["surf"][..., 3] > 0
        same = hit == (re["surf"][..., 3] > 0)
        assert same.mean() > 0.995
        assert np.allclose(dpo[same], dpe[same], rtol=2e-5, atol=1e-4)
        mvo = np.stack([_half((mo & 0xFFFF).astype(np.uint16)), _half((mo >> 16).astype(np.uint16))], -1)
        mve = np.stack([_half((me & 0xFFFF).astype(np.uint16)), _half((me >> 16).astype(np.uint16))], -1)
        assert (np.abs(mvo - mve)[same] > 0.02).mean() < 0.01                  # pixel units; fp16 storage + inverse(inverse()) vs proj*view
        if f == 0: assert (mo == 0).all() and (me == 0).all()                # frame 0: stale prevViewProj, zero motion
        else: assert np.abs(mve[hit & same]).max() > 0.05                     # the camera moved
        do = O.svgf_dispatch(vo, ro["ldr"], mo, dpo, so.normal, f == 0)
        de = E.svgf_dispatch(ve, re["ldr"], me, dpe, se.normal, f == 0)
        assert (np.abs(do.astype(np.int16) - de.astype(np.int16)).max(-1) > 2).mean() < 0.03, f

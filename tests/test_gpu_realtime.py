"""Realtime profile on the GPU (k_raygen_rt / k_bounce_rt / k_rt_pixel / k_rt_denoise through the C ABI) against the oracle's
restatement of pt_raygen_realtime.rgen.  Parity is CUDA-vs-oracle only: the reference's realtime output is nondeterministic
(quirk Q12) and has no golden."""
import numpy as np
import pytest

from ohao_engine_b200 import binding as B
from oracle import oracle_py as O
from tests import util

pytestmark = pytest.mark.gpu


def _settings(r, spf=1, flags=None):
    st = r.get_settings(); st.samples_per_frame = spf
    if flags is not None: st.flags = flags
    r.set_rt_render_settings(st)


@pytest.mark.parametrize("which,spf,res", [("cornell", 2, (160, 90)), ("helmet_small", 1, (160, 90)), ("synthetic_small", 1, (128, 72))])
def test_realtime_frames_match_oracle(which, spf, res, request):
    ps, cam = request.getfixturevalue(which)
    W, H = res
    osc = O.OracleScene(ps)
    so = O.RealtimeState(W, H)
    r = B.Renderer(W, H, profile=B.PROFILE_REALTIME); r.set_scene(ps); _settings(r, spf)
    s0 = r.get_settings()
    assert s0.max_bounces == 2 and s0.firefly_clamp_lum == 10.0 and s0.flags & 7 == 7      # kRealtimeRTSettings
    st = O.realtime_settings(spf=spf)
    for f in range(5):
        r.reset_counters()
        ro = osc.render_realtime(so, cam.view(), cam.proj(W, H), settings=st, dumps=True)
        got = r.render_realtime(cam.view(), cam.proj(W, H), dumps=True)
        acc, _, _ = r.readback_hdr_buffers(want_aov=False)
        got["accum"] = acc
        for key in ("radiance", "gi", "accum", "denoised"):
            bad, worst = util.sample_parity(ro[key][None], got[key][None])
            assert bad < 0.015 and worst < 2e-3, (f, key, bad, worst)
        state = r.realtime_state()
        assert (ro["reservoirs"][0][..., 3] != state["reservoirs"][0][..., 3]).mean() < 0.015
        assert np.allclose(ro["surf"], state["surf"], rtol=1e-5, atol=1e-6 * float(np.abs(ro["surf"]).max()))
        ldr = r.get_pixels()
        assert (np.abs(ldr.astype(np.int16) - ro["ldr"].astype(np.int16)).max(-1) > 1).mean() < 0.015
        c = r.counters()
        assert c["samples"] == W * H * spf
        for k in ("closest_rays", "shadow_rays"):
            assert abs(ro["counters"][k] - c[k]) <= max(16, ro["counters"][k] // 300), (f, k, ro["counters"][k], c[k])
    assert r.frame_index() == 5
    assert state["reservoirs"][0][..., 3].max() > 2 * spf          # temporal reuse active


def test_realtime_temporal_mean_converges_to_oracle(cornell):
    """Statistical check over many frames (SURVEY §7 "compare statistically"): time-averaged accum images agree."""
    ps, cam = cornell
    W, H, F = 96, 54, 24
    osc = O.OracleScene(ps); so = O.RealtimeState(W, H)
    r = B.Renderer(W, H, profile=B.PROFILE_REALTIME); r.set_scene(ps)
    mo = np.zeros((H, W, 3)); mg = np.zeros((H, W, 3))
    for f in range(F):
        ro = osc.render_realtime(so, cam.view(), cam.proj(W, H))
        r.render_realtime(cam.view(), cam.proj(W, H))
        acc, _, _ = r.readback_hdr_buffers(want_aov=False)
        mo += ro["accum"][..., :3]; mg += acc[..., :3]
    assert abs(mo.mean() - mg.mean()) / mo.mean() < 0.01
    assert util.psnr(mg / F, mo / F, float((mo / F).max())) > 35.0


def test_realtime_reset_view_change_and_errors(cornell):
    ps, cam = cornell
    W, H = 64, 36
    r = B.Renderer(W, H, profile=B.PROFILE_REALTIME); r.set_scene(ps)
    v, p = cam.view(), cam.proj(W, H)
    a = r.render_realtime(v, p, dumps=True); r.render_realtime(v, p)
    r.reset_accumulation()
    assert r.frame_index() == 0
    b = r.render_realtime(v, p, dumps=True)
    assert np.array_equal(a["radiance"], b["radiance"])                 # same frame index + no history -> same frame
    r.notify_camera_changed()                                           # realtime keeps accumulating (viewChanged flag only)
    r.render_realtime(v, p); assert r.frame_index() == 2
    st = r.get_settings(); st.flags |= 1 << 7                           # legacy Stage C: refused loudly
    r.set_rt_render_settings(st)
    with pytest.raises(B.OhbError, match="LEGACY"):
        r.render_realtime(v, p)
    off = B.Renderer(W, H); off.set_scene(ps)
    st = off.get_settings(); st.profile = 1
    with pytest.raises(B.OhbError):
        off.set_rt_render_settings(st)

"""Realtime integrator + ReSTIR GI: the product's per-thread code (ohb_realtime.h, driven by the host emulator) against the
oracle's restatement of pt_raygen_realtime.rgen.  CPU-only.  State feeds back frame to frame, so a libm-vs-libm branch flip
in frame k shows up as a differing pixel in later frames; the gates are on the fraction of such pixels."""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests import util
from tests.emul import emul_py as E


@pytest.mark.parametrize("which,spf", [("cornell", 2), ("helmet_small", 1)])
def test_realtime_frames_match_oracle(which, spf, request):
    ps, cam = request.getfixturevalue(which)
    osc, esc = O.OracleScene(ps), E.EmulScene(ps)
    W, H = 80, 45
    so, se = O.RealtimeState(W, H), O.RealtimeState(W, H)
    st = O.realtime_settings(spf=spf)
    for f in range(4):
        ro = osc.render_realtime(so, cam.view(), cam.proj(W, H), settings=st, dumps=True)
        re = esc.render_realtime(se, cam.view(), cam.proj(W, H), settings=st, dumps=True)
        for key in ("radiance", "gi", "accum", "denoised"):
            bad, worst = util.sample_parity(ro[key][None], re[key][None])
            assert bad < 0.01 and worst < 2e-3, (f, key, bad, worst)
        assert (ro["reservoirs"][0][..., 3] != re["reservoirs"][0][..., 3]).mean() < 0.01            # M
        assert (ro["ldr"] != re["ldr"]).mean() < 0.01
        for k in ("closest_rays", "shadow_rays"):
            assert abs(ro["counters"][k] - re["counters"][k]) <= max(16, ro["counters"][k] // 500), (f, k)
    # temporal reuse is active: confidence M grows past the per-frame sample count and the reservoirs hold valid samples
    M = re["reservoirs"][0][..., 3]
    assert M.max() > spf * 2 and (re["reservoirs"][2][..., 3] > 0.5).mean() > 0.3


def test_realtime_flags_and_view_change(cornell):
    ps, cam = cornell
    esc, osc = E.EmulScene(ps), O.OracleScene(ps)
    W, H = 64, 36
    for flags in (1 | 2 | 4 | (1 << 5), 1 | 2 | 4 | (1 << 8), 1 | 4):      # GI reuse off / no spatial / no internal denoise
        so, se = O.RealtimeState(W, H), O.RealtimeState(W, H)
        st = O.realtime_settings(flags=flags)
        for f in range(3):
            ro = osc.render_realtime(so, cam.view(), cam.proj(W, H), settings=st, view_changed=(f == 2), dumps=True)
            re = esc.render_realtime(se, cam.view(), cam.proj(W, H), settings=st, view_changed=(f == 2), dumps=True)
            bad, worst = util.sample_parity(ro["accum"][None], re["accum"][None])
            assert bad < 0.01 and worst < 2e-3, (flags, f, bad, worst)
        if flags & (1 << 5):
            assert re["reservoirs"][0][..., 3].max() <= 1.0          # no temporal merge: M stays at the per-frame count
        if not (flags & 2):
            assert np.allclose(re["denoised"][..., :3], re["accum"][..., :3])

"""End-to-end pin of the oracle's offline integrator against the reference's golden image.

tests/golden/reference_cornell_box_16spp_640.png is the reference's own tests/golden/cornell_box.png
(`cornell_box out.png 16 --denoise=none`, 1920x1080 downscaled to 640 wide, manifest.json:4-10).
The golden was rendered by a revision whose Stage C spec-lobe continuation still used Stage B's
throughput (DESIGN.md "Golden image"); OHB_FLAG_GOLDEN_COMPAT reproduces it.  With it the oracle
matches to <= 1 LSB on > 99.8 % of pixels; without it (HEAD behaviour) the images differ in the
expected one-sided way.  A band of rows is rendered to keep the CPU suite fast."""
import os

import numpy as np
from PIL import Image

from oracle import oracle_py as O
from tests import util

BAND = (297, 513)     # full-res rows [y0, y1), multiples of 3 -> 640-scale rows 99..170


def _band_stats(ldr_band, golden, y0, y1):
    im = Image.fromarray(np.ascontiguousarray(ldr_band[..., :3]))
    small = np.asarray(im.resize((640, (y1 - y0) // 3), Image.BILINEAR), np.int16)
    a = small[2:-2]; g = golden[y0 // 3 + 2: y1 // 3 - 2]      # drop rows whose filter support crosses the band edge
    return util.golden_stats(a, g)


def test_oracle_matches_reference_golden(cornell, golden_dir):
    ps, cam = cornell
    golden = np.asarray(Image.open(os.path.join(golden_dir, "reference_cornell_box_16spp_640.png")).convert("RGB"), np.int16)
    W, H = 1920, 1080
    y0, y1 = BAND
    sc = O.OracleScene(ps)
    st = O.offline_settings(flags=1 | (1 << 16))
    r = sc.render_offline(cam.view(), cam.proj(W, H), W, H, 16, settings=st, tile=(0, y0, W, y1 - y0))
    s = _band_stats(r["ldr"][y0:y1], golden, y0, y1)
    assert s["frac_gt1"] < 2e-3, s      # measured 6e-4 on the full frame
    assert s["frac_gt4"] < 5e-4, s
    assert s["rmse"] < 0.4, s           # measured 0.19
    # HEAD behaviour (no compat flag) is measurably different from the stale golden, one-sidedly darker
    r2 = sc.render_offline(cam.view(), cam.proj(W, H), W, H, 16, tile=(0, y0, W, y1 - y0))
    s2 = _band_stats(r2["ldr"][y0:y1], golden, y0, y1)
    assert s2["frac_gt1"] > 0.1 and s2["rmse"] > 1.0, s2

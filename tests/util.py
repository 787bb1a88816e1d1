"""Shared helpers for the parity tests (oracle = checker, never the thing under test)."""
import numpy as np
from PIL import Image

from oracle import oracle_py as O


def random_rays(n, lo, hi, seed, tmax=1.0e4):
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, O.RAY_DTYPE)
    rays["origin"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["dir"] = d.astype(np.float32); rays["tmin"] = 0.001; rays["tmax"] = tmax
    # a few axis-parallel and grazing rays: the cases a slab test gets wrong first
    k = min(n // 8, 4096)
    ax = rng.integers(0, 3, k); rays["dir"][:k] = 0.0
    rays["dir"][np.arange(k), ax] = rng.choice([-1.0, 1.0], k)
    return rays


def compare_hits(ref, got):
    """Bit-exact primitive ids; returns (mismatches, ties) where a tie = same t, different id."""
    mism = ref["prim"] != got["prim"]
    ties = mism & (ref["t"] == got["t"])
    return int(mism.sum()), int(ties.sum())


def sample_parity(ref_samples, got_samples, rel=2e-3, floor=1e-3):
    """Per-sample radiance parity: fraction of samples off by more than `rel` (branch flips from
    libm-vs-CUDA rounding land here) and the max relative error of the rest."""
    a = ref_samples[..., :3].astype(np.float64); b = got_samples[..., :3].astype(np.float64)
    err = np.abs(a - b).max(-1) / (np.abs(a).max(-1) + floor)
    bad = err > rel
    return float(bad.mean()), float(err[~bad].max() if (~bad).any() else 0.0)


def downscale640(img_rgb8):
    """render_golden.py:38-42 — bilinear to 640 wide."""
    im = Image.fromarray(np.ascontiguousarray(img_rgb8[..., :3]))
    w, h = im.size
    return np.asarray(im.resize((640, max(1, round(h * 640 / w))), Image.BILINEAR), np.int16)


def golden_stats(actual640, golden640):
    d = np.abs(actual640 - golden640)
    m = d.max(2)
    return dict(max_abs=int(m.max()), diff_frac=float((m > 0).mean()), frac_gt1=float((m > 1).mean()), frac_gt4=float((m > 4).mean()),
                rmse=float(np.sqrt((d.astype(np.float64) ** 2).mean())))


def psnr(a, b, peak=1.0):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)

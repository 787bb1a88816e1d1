"""Where do the per-sample differences between the CUDA path and the oracle come from on the 2 M-triangle scene?
Classifies the samples of a 1080p tile that differ by more than 2e-3 by the roughness / metalness of the material the
primary ray hits, and repeats the comparison on the same scene with every roughness clamped to >= 0.3.
  python tools/diag_parity_2m.py            (GPU box)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ohao_engine_b200 import binding as B, scenes
from oracle import oracle_py as O

W, H, spp = 1920, 1080, 2
TILE = (832, 476, 256, 128); x0, y0, w, h = TILE
ps, cam = scenes.synthetic_2m(), scenes.synthetic_camera()
v, p = cam.view(), cam.proj(W, H)

def compare(ps, label):
    osc = O.OracleScene(ps); r = B.Renderer(W, H); r.set_scene(ps)
    ro = osc.render_offline(v, p, W, H, spp, tile=TILE, dump=True)
    r.set_tile(*TILE); got = r.render(v, p, spp, dump=True)
    a = ro["samples"][:, y0:y0 + h, x0:x0 + w, :3].astype(np.float64); b = got[:, y0:y0 + h, x0:x0 + w, :3].astype(np.float64)
    err = np.abs(a - b).max(-1) / (np.abs(a).max(-1) + 1e-3)
    print(f"[{label}] beyond 2e-3: {np.mean(err > 2e-3):.3e}   beyond 2e-2: {np.mean(err > 2e-2):.3e}   beyond 0.5: {np.mean(err > 0.5):.3e}   "
          f"mean image rel err {np.abs(a.mean(0) - b.mean(0)).mean() / a.mean():.2e}")
    return err, osc

err, osc = compare(ps, "as generated")
# primary-hit material of each pixel of the tile (pixel-centre ray)
rays, hits, kinds = osc.record_rays(v, p, W, H, 1, tile=TILE, cap=1 << 21)
# classify by the first closest-hit ray of each pixel: rays are recorded pixel by pixel, primary first; simpler: trace pixel centres
iv = np.linalg.inv(np.asarray(v, np.float64).reshape(4, 4).T); ip = np.linalg.inv(np.asarray(p, np.float64).reshape(4, 4).T)
ys, xs = np.mgrid[y0:y0 + h, x0:x0 + w]
ndc = np.stack([(xs + 0.5) / W * 2 - 1, (ys + 0.5) / H * 2 - 1, np.ones_like(xs, float), np.ones_like(xs, float)], -1).reshape(-1, 4)
t = (ip @ ndc.T).T; t = t[:, :3] / t[:, 3:4]; d = (iv[:3, :3] @ (t / np.linalg.norm(t, axis=1, keepdims=True)).T).T
pr = np.zeros(len(d), O.RAY_DTYPE); pr["origin"] = iv[:3, 3]; pr["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True); pr["tmin"] = 1e-3; pr["tmax"] = 1e4
ph = osc.trace(pr)["prim"].reshape(h, w)
mat = np.where(ph != 0xFFFFFFFF, ps.mat_ids[np.minimum(ph, ps.ntris - 1)].astype(np.int64), -1)
mc = np.asarray(ps.mat_colors, np.float32).reshape(-1, 3, 4)
rough = np.where(mat >= 0, mc[np.maximum(mat, 0), 1, 0], np.nan); metal = np.where(mat >= 0, mc[np.maximum(mat, 0), 1, 1], np.nan)
bad = (err > 2e-3)
for lo, hi in [(0.0, 0.1), (0.1, 0.2), (0.2, 0.4), (0.4, 1.01)]:
    m = (rough >= lo) & (rough < hi)
    if m.sum(): print(f"  primary hit rough [{lo},{hi}): {m.mean():.3f} of pixels, bad fraction {bad[:, m].mean():.3e} (metal share {np.nanmean(metal[m]):.2f})")
m = np.isnan(rough); print(f"  primary miss: {m.mean():.3f} of pixels, bad fraction {bad[:, m].mean() if m.sum() else 0:.3e}")
print("  materials: rough", np.round(mc[:, 1, 0], 3).tolist())
import copy
ps2 = copy.deepcopy(ps); mc2 = np.asarray(ps2.mat_colors, np.float32).reshape(-1, 3, 4).copy(); mc2[:, 1, 0] = np.maximum(mc2[:, 1, 0], 0.3); ps2.mat_colors = mc2.reshape(ps.mat_colors.shape)
compare(ps2, "roughness clamped to >= 0.3")

"""tools/emul_stats.py — node visits / triangle tests per ray of the traversal on the host emulator (no GPU needed).

Builds a workload's acceleration structure with the emulator build of the product headers and traces primary rays and
cosine-distributed bounce rays from their hit points.  EMUL_SO selects an alternative emulator build (A/B of builder or
traversal variants compiled with different -D flags).  Measurement tooling only.
  python tools/emul_stats.py [synthetic2m|helmet|cornell] [nrays]
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.emul import emul_py as E
from ohao_engine_b200 import scenes
import ctypes as C

if os.environ.get("EMUL_SO"):
    E.lib()
    so = os.environ["EMUL_SO"]
    L = C.CDLL(so)
    for name in ("emul_scene_create", "emul_scene_destroy", "emul_trace_batch", "emul_occluded_batch", "emul_accel_stats", "emul_accel_levels", "emul_trav_stats"):
        f = getattr(L, name); g = getattr(E._LIB, name); f.argtypes = g.argtypes; f.restype = g.restype
    E._LIB = L

wl = sys.argv[1] if len(sys.argv) > 1 else "synthetic2m"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
ps, cam = {"synthetic2m": (scenes.synthetic_2m, scenes.synthetic_camera), "helmet": (scenes.helmet_class, scenes.helmet_camera),
           "cornell": (scenes.cornell_box, scenes.cornell_camera)}[wl]
ps, cam = ps(), cam()
t0 = time.time(); es = E.EmulScene(ps); t1 = time.time()
nodes, sah = es.stats()
print(f"{wl}: {ps.ntris} tris, {nodes} wide nodes, SAH {sah:.2f}, levels {es.levels()}, build {t1 - t0:.1f} s (emulated)")
W, H = 1920, 1080
rng = np.random.default_rng(1)
px = rng.integers(0, W, n); py = rng.integers(0, H, n)
view = np.asarray(cam.view(), np.float64).reshape(4, 4); proj = np.asarray(cam.proj(W, H), np.float64).reshape(4, 4)
iv = np.linalg.inv(view.T).T if False else np.linalg.inv(view.T)   # column-major storage -> view.T is the math matrix
ip = np.linalg.inv(proj.T)
ndc = np.stack([(px + 0.5) / W * 2 - 1, (py + 0.5) / H * 2 - 1, np.ones(n), np.ones(n)], 1)
tgt = (ip @ ndc.T).T; tgt = tgt[:, :3] / tgt[:, 3:4]
d = (iv[:3, :3] @ (tgt / np.linalg.norm(tgt, axis=1, keepdims=True)).T).T
o = np.broadcast_to(iv[:3, 3], (n, 3))
rays = np.zeros(n, E.O.RAY_DTYPE); rays["origin"] = o; rays["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True); rays["tmin"] = 1e-3; rays["tmax"] = 1e4
def run(r, label):
    E.trav_stats(True); t = time.time(); h = es.trace(r); dt = time.time() - t
    a, b = E.trav_stats(True)
    print(f"  {label:10s} {len(r)} rays: {a / len(r):6.2f} nodes/ray {b / len(r):6.2f} tris/ray  hit {np.mean(h['prim'] != 0xFFFFFFFF):.3f}  ({dt:.1f} s)")
    return h
h = run(rays, "primary")
for bounce in range(2):
    m = h["prim"] != 0xFFFFFFFF
    p = rays["origin"][m] + rays["dir"][m] * h["t"][m][:, None]
    k = int(m.sum())
    nd = rng.normal(size=(k, 3)); nd /= np.linalg.norm(nd, axis=1, keepdims=True)
    # flip into the hemisphere facing back along the incoming ray (a stand-in for the surface normal)
    s = np.sign(-(nd * rays["dir"][m]).sum(1)); nd *= s[:, None]
    r2 = np.zeros(k, E.O.RAY_DTYPE); r2["origin"] = p + nd * 1e-2; r2["dir"] = nd; r2["tmin"] = 1e-3; r2["tmax"] = 1e4
    rays = r2; h = run(rays, f"bounce{bounce + 1}")

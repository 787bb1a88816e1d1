"""A/B of the two realisations of the BLAS/TLAS interface on the 2 M-triangle scene split into 1001 instances (1000 blobs of
2000 triangles + the ground; SURVEY §8d config 3 "as 1 000 actors (TLAS stress)"): build time, MODE_UPDATE time, ray-batch
and render throughput, hit agreement.      python tools/tlas_ab.py [outfile]          (GPU box)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ohao_engine_b200 import binding as B, scenes
from tests import util

ps, cam = scenes.synthetic_2m(), scenes.synthetic_camera()
nb, per = 1000, (ps.ntris - 2) // 1000
inst = np.zeros(nb + 1, ps.instances.dtype)
inst["first_tri"] = np.arange(nb + 1) * per; inst["tri_count"] = per; inst["tri_count"][-1] = 2; inst["xform"] = scenes.trs(); inst["mask"] = 0xFF
ps.instances = inst
rng = np.random.default_rng(2)
moved = inst.copy()
for k in range(nb):                                         # every blob drifts (translation only keeps the object-space geometry identical)
    x = moved[k]["xform"].copy(); x[3] += rng.uniform(-2, 2); x[7] += rng.uniform(0, 1); x[11] += rng.uniform(-2, 2); moved[k]["xform"] = x
W, H = 1920, 1080
rays = util.random_rays(1 << 21, -60, 60, seed=5)
res = {"scene": f"{ps.ntris} triangles in {len(inst)} instances", "modes": {}}
hits = {}
for name, two in (("flatten", False), ("two_level", True)):
    r = B.Renderer(W, H); r.set_accel_mode(two); r.set_scene(ps)
    r.build_accel(); st = r.accel_stats()                   # second build: without module load / allocation
    e = {"build_ms": float(st.build_ms), "levels": int(st.levels)}
    t0 = time.perf_counter(); r.trace(rays[:4096]); r.synchronize()
    t0 = time.perf_counter(); h = r.trace(rays); e["trace_batch_mrays_s_incl_copies"] = len(rays) / (time.perf_counter() - t0) / 1e6
    v, p = cam.view(), cam.proj(W, H)
    r.render(v, p, 16); r.render(v, p, 16); r.synchronize(); r.reset_accumulation(); r.reset_counters()      # warm-up with the same batch size (path-state allocation)
    r.timer_start(); r.render(v, p, 16); r.render(v, p, 16); ms = r.timer_stop(); c = r.counters()
    e["render_msamples_s"] = W * H * 32 / ms / 1e3; e["render_grays_s"] = (c["closest_rays"] + c["shadow_rays"]) / ms / 1e6
    ups = []
    for k in range(3):
        st = r.update_instances(moved if k % 2 == 0 else inst); ups.append(float(st.update_ms))
    e["update_ms"] = min(ups)
    r.update_instances(moved); hits[name] = r.trace(rays)
    res["modes"][name] = e
    del r
a, b = hits["flatten"], hits["two_level"]
same = a["prim"] == b["prim"]
res["agreement_after_update"] = {"rays": len(rays), "hit_rate": float((a["prim"] != 0xFFFFFFFF).mean()), "same_prim": float(same.mean()),
                                 "t_within_1e-4": float(np.isclose(a["t"][same], b["t"][same], rtol=1e-4, atol=1e-4).mean())}
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2t_tlas_ab.json")
json.dump(res, open(out, "w"), indent=1); print(json.dumps(res, indent=1))

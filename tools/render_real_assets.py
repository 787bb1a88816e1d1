"""Config 2 with its first-named inputs: `turntable DamagedHelmet.glb env` (examples/turntable.cpp, 1280x720, env_outdoor.hdr)
rendered by the CUDA path and by the CPU oracle; writes the GPU image, throughput and the oracle-vs-GPU agreement.
  python tools/render_real_assets.py [spp=256] [oracle_spp=64] [outdir=gpurun_out]            (GPU box)
The reference's own tests/reference_scenes/custom/envlit_turntable/truth_4096spp.png was rendered by a different entry
point (env_demo: studio HDRI, ground plane, 3-point area lights) with an older integrator, so it is not comparable to
this scene; agreement is therefore stated against the oracle (the restatement of HEAD's shaders)."""
import json, os, sys, time
import numpy as np
from PIL import Image
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ohao_engine_b200 import assets, binding as B
from oracle import oracle_py as O
from tests import util

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ospp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out")
W, H = 1280, 720
t0 = time.time()
ps = assets.turntable_env_scene(os.path.join(ROOT, "tests/assets/DamagedHelmet.glb"), os.path.join(ROOT, "tests/assets/env_outdoor.hdr"))
t_load = time.time() - t0
r = B.Renderer(W, H); t0 = time.time(); r.set_scene(ps); t_upload = time.time() - t0
st = r.accel_stats()
res = {"scene": ps.name, "tris": int(st.num_tris), "texture_layers": list(ps.textures.shape), "load_s": t_load, "upload_and_build_s": t_upload, "bvh_build_ms": float(st.build_ms), "frames": []}
for frame in (0, 30, 75):
    cam = assets.turntable_camera(frame)
    v, p = cam.view(), cam.proj(W, H)
    r.reset_accumulation(); r.reset_counters(); r.timer_start(); r.render(v, p, spp); ms = r.timer_stop()
    c = r.counters()
    ldr = r.get_pixels(); acc, _, _ = r.readback_hdr_buffers(want_aov=False)
    Image.fromarray(ldr[..., :3]).resize((640, 360), Image.BILINEAR).save(os.path.join(out, f"r2_turntable_env_helmet_{frame:03d}_{spp}spp_640.png"))
    e = {"frame": frame, "spp": spp, "ms": ms, "msamples_per_s": W * H * spp / ms / 1e3, "rays_per_sample": (c["closest_rays"] + c["shadow_rays"]) / c["samples"]}
    if frame == 30 and ospp:
        osc = O.OracleScene(ps)
        t0 = time.time(); ro = osc.render_offline(v, p, W, H, ospp); e["oracle_s"] = time.time() - t0
        r.reset_accumulation(); r.render(v, p, ospp); acc2, _, _ = r.readback_hdr_buffers(want_aov=False)
        ref = ro["accum"][..., :3]; got = acc2[..., :3]
        peak = float(np.percentile(ref, 99.9))
        e["vs_oracle_same_samples"] = {"spp": ospp, "psnr_db": util.psnr(np.minimum(got, peak), np.minimum(ref, peak), peak), "mean_rel_err": float(np.abs(got - ref).mean() / ref.mean()),
                                       "ldr_pixels_differing": float((np.abs(r.get_pixels().astype(int) - ro["ldr"].astype(int)).max(-1) > 1).mean())}
    res["frames"].append(e)
json.dump(res, open(os.path.join(out, "r2_real_assets.json"), "w"), indent=1)
print(json.dumps(res, indent=1))

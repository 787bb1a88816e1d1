import numpy as np, sys
sys.path.insert(0, "/root/repo")
from ohao_engine_b200 import binding as B, scenes
for name, (ps, cam) in {"cornell": (scenes.cornell_box(), scenes.cornell_camera()), "helmet": (scenes.helmet_class(ntris=5000, tex_size=256, env_size=(256,128)), scenes.helmet_camera())}.items():
    W, H = 384, 216
    imgs = []
    for dev in (0, 1, 0, 1):
        r = B.Renderer(W, H, device=dev); r.set_scene(ps); r.set_render_seed(1234); r.render(cam.view(), cam.proj(W, H), 32)
        acc, _, _ = r.readback_hdr_buffers(want_aov=False); imgs.append(acc.copy()); print(name, dev, acc[..., :3].mean(), r.counters())
    for i in range(1, 4): print(name, "diff vs first", i, np.abs(imgs[i] - imgs[0]).max(), (imgs[i] != imgs[0]).mean())

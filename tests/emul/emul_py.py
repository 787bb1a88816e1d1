"""ctypes binding of tests/emul/libemul.so — host emulator of the CUDA kernels (TEST TOOLING ONLY)."""
import ctypes as C
import os
import subprocess
import numpy as np

from oracle import oracle_py as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SceneDesc(C.Structure):
    _fields_ = O.SceneDesc._fields_ + [("marg", C.c_void_p), ("cond", C.c_void_p), ("env_integral", C.c_float)]


class RenderArgs(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("width", C.c_uint32), ("height", C.c_uint32),
                ("first_sample_index", C.c_uint32), ("history_count", C.c_uint32), ("nsamples", C.c_uint32),
                ("tile_x", C.c_uint32), ("tile_y", C.c_uint32), ("tile_w", C.c_uint32), ("tile_h", C.c_uint32),
                ("settings", O.Settings),
                ("accum", C.c_void_p), ("ldr", C.c_void_p), ("albedo", C.c_void_p), ("normal", C.c_void_p), ("sample_dump", C.c_void_p),
                ("counters", C.c_uint64 * 4)]


class RTArgs(C.Structure):
    _fields_ = O.RTArgs._fields_[:-2] + [("counters", C.c_uint64 * 4)]


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", _HERE, "libemul.so"], stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(os.path.join(_HERE, "libemul.so"))
        _LIB.emul_scene_create.restype = C.c_void_p
        _LIB.emul_scene_create.argtypes = [C.POINTER(SceneDesc)]
        _LIB.emul_scene_destroy.argtypes = [C.c_void_p]
        _LIB.emul_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _LIB.emul_occluded_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _LIB.emul_env_sample_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        _LIB.emul_render_offline.argtypes = [C.c_void_p, C.POINTER(RenderArgs)]
        _LIB.emul_render_realtime.argtypes = [C.c_void_p, C.POINTER(RTArgs)]
        _LIB.emul_accel_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        _LIB.emul_accel_levels.argtypes = [C.c_void_p]; _LIB.emul_accel_levels.restype = C.c_uint32
        _LIB.emul_set_warp_noise.argtypes = [C.c_uint]
        _LIB.emul_svgf_dispatch.argtypes = [C.POINTER(O.SvgfArgs)]
        _LIB.emul_svgf_guides.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32, C.c_void_p, C.c_void_p]
        _LIB.emul_f2h.restype = C.c_uint16; _LIB.emul_f2h.argtypes = [C.c_float]
        _LIB.emul_h2f.restype = C.c_float; _LIB.emul_h2f.argtypes = [C.c_uint16]
        _LIB.emul_trav_stats.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.c_int]
        _LIB.emul_scene_set_accel_mode.argtypes = [C.c_void_p, C.c_int]
        _LIB.emul_scene_update_instances.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        _LIB.emul_hybrid_shadow.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.emul_hybrid_gi.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.emul_nrd_pack_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


_p = O._p


class EmulScene:
    def __init__(self, ps):
        self.ps = ps
        d = SceneDesc()
        k = self._keep = [np.ascontiguousarray(x) for x in (ps.positions, ps.indices, ps.normals, ps.uvs, ps.mat_ids, ps.instances,
                                                            ps.mat_colors, ps.textures, ps.light_ssbo)]
        d.positions = _p(k[0]); d.stride_bytes = k[0].strides[0]; d.nverts = ps.nverts
        d.indices = _p(k[1]); d.ntris = ps.ntris; d.normals = _p(k[2]); d.uvs = _p(k[3]); d.mat_ids = _p(k[4])
        d.instances = _p(k[5]); d.ninstances = len(k[5]); d.mat_colors = _p(k[6]); d.nmaterials = ps.nmaterials
        d.textures = _p(k[7]); d.tex_layers, d.tex_h, d.tex_w = k[7].shape[0], k[7].shape[1], k[7].shape[2]
        d.light_ssbo = _p(k[8]); d.light_bytes = k[8].nbytes
        if ps.env is not None:
            self._env = np.ascontiguousarray(ps.env, np.float32)
            self._marg, self._cond, integral = O.env_cdf(self._env)   # the device CDF kernel is not emulated
            d.env = _p(self._env); d.env_h, d.env_w = self._env.shape[:2]
            d.marg = _p(self._marg); d.cond = _p(self._cond); d.env_integral = integral
        self.h = lib().emul_scene_create(C.byref(d))

    def __del__(self):
        if getattr(self, "h", None):
            lib().emul_scene_destroy(self.h); self.h = None

    def stats(self):
        n = C.c_uint32(); s = C.c_float(); lib().emul_accel_stats(self.h, C.byref(n), C.byref(s)); return n.value, s.value

    def levels(self):
        return lib().emul_accel_levels(self.h)

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, O.RAY_DTYPE); hits = np.zeros(len(rays), O.HIT_DTYPE)
        lib().emul_trace_batch(self.h, _p(rays), len(rays), _p(hits)); return hits

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays, O.RAY_DTYPE); occ = np.zeros(len(rays), np.uint8)
        lib().emul_occluded_batch(self.h, _p(rays), len(rays), _p(occ)); return occ

    def set_accel_mode(self, two_level: bool):
        lib().emul_scene_set_accel_mode(self.h, int(two_level))

    def update_instances(self, instances):
        inst = np.ascontiguousarray(instances); lib().emul_scene_update_instances(self.h, _p(inst), len(inst))

    def hybrid_shadow(self, W, H, gpos, gnrm, params):
        gpos = np.ascontiguousarray(gpos, np.float32); gnrm = np.ascontiguousarray(gnrm, np.float32); mask = np.zeros((H, W), np.uint8)
        lib().emul_hybrid_shadow(self.h, W, H, _p(gpos), _p(gnrm), C.byref(params), _p(mask)); return mask

    def hybrid_gi(self, W, H, gpos, gnrm, galbedo, history, inst_mat, params):
        a = [np.ascontiguousarray(x, np.float32) for x in (gpos, gnrm, galbedo, history, inst_mat)]; out = np.zeros((H, W, 4), np.uint16)
        lib().emul_hybrid_gi(self.h, W, H, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), C.byref(params), _p(out)); return out

    def env_sample(self, u12):
        u = np.ascontiguousarray(u12, np.float32); n = len(u); dp = np.zeros((n, 4), np.float32); pd = np.zeros(n, np.float32)
        lib().emul_env_sample_batch(self.h, _p(u), n, _p(dp), _p(pd)); return dp, pd

    def render_realtime(self, st, view, proj, settings=None, view_changed=False, dumps=False, fresh=False):
        """Same contract as OracleScene.render_realtime (st = oracle_py.RealtimeState)."""
        a = RTArgs()
        a.view[:] = [float(x) for x in view]; a.proj[:] = [float(x) for x in proj]; a.prev_view_proj[:] = [float(x) for x in st.prev_view_proj]
        a.width, a.height, a.frame_index, a.history_count, a.view_changed = st.W, st.H, st.frame_index, (0 if fresh else st.history), int(view_changed)
        a.settings = settings or O.realtime_settings()
        p, c = st.cur, 1 - st.cur
        a.accum_prev, a.accum_curr = _p(st.accum[p]), _p(st.accum[c]); a.surf_prev, a.surf_curr = _p(st.surf[p]), _p(st.surf[c])
        a.shad_prev, a.shad_curr = _p(st.shad[p]), _p(st.shad[c])
        for k in range(3):
            a.res_prev[k] = _p(st.res[p][k]); a.res_curr[k] = _p(st.res[c][k])
        a.albedo, a.normal = _p(st.albedo), _p(st.normal)
        rad = np.zeros((st.H, st.W, 4), np.float32) if dumps else None; gi = np.zeros((st.H, st.W, 4), np.float32) if dumps else None
        den = np.zeros((st.H, st.W, 4), np.float32); ldr = np.zeros((st.H, st.W, 4), np.uint8)
        a.radiance_dump, a.gi_dump, a.denoised, a.ldr = _p(rad), _p(gi), _p(den), _p(ldr)
        lib().emul_render_realtime(self.h, C.byref(a))
        st.cur = c; st.frame_index += 1; st.history += 1
        v = np.asarray(view, np.float32).reshape(4, 4); pr = np.asarray(proj, np.float32).reshape(4, 4)
        st.prev_view_proj = (v @ pr).reshape(16).astype(np.float32)
        cnt = a.counters
        return dict(accum=st.accum[c], ldr=ldr, denoised=den, radiance=rad, gi=gi, reservoirs=st.res[c], surf=st.surf[c], shad=st.shad[c],
                    counters=dict(samples=cnt[0], closest_rays=cnt[1], shadow_rays=cnt[2], closest_hits=cnt[3]))

    def render_offline(self, view, proj, width, height, nsamples, first_sample=0, history=0, accum=None, settings=None, tile=None, dump=False):
        a = RenderArgs()
        a.view[:] = [float(x) for x in view]; a.proj[:] = [float(x) for x in proj]
        a.width, a.height, a.first_sample_index, a.history_count, a.nsamples = width, height, first_sample, history, nsamples
        if tile: a.tile_x, a.tile_y, a.tile_w, a.tile_h = tile
        a.settings = settings or O.offline_settings()
        if accum is None: accum = np.zeros((height, width, 4), np.float32)
        ldr = np.zeros((height, width, 4), np.uint8); alb = np.zeros((height, width, 4), np.float32); nrm = np.zeros((height, width, 4), np.float32)
        sd = np.zeros((nsamples, height, width, 4), np.float32) if dump else None
        a.accum, a.ldr, a.albedo, a.normal, a.sample_dump = _p(accum), _p(ldr), _p(alb), _p(nrm), _p(sd)
        lib().emul_render_offline(self.h, C.byref(a))
        c = a.counters
        return dict(accum=accum, ldr=ldr, albedo=alb, normal=nrm, samples=sd,
                    counters=dict(samples=c[0], closest_rays=c[1], shadow_rays=c[2], closest_hits=c[3]))


def svgf_dispatch(st, beauty, motion, depth, normal, reset, sigmas=O.SVGF_SIGMAS):
    """One AtrousDenoiser::dispatch through the host build of ohb_svgf.h (st = oracle_py.SvgfState)."""
    return O._svgf_dispatch(lib().emul_svgf_dispatch, st, beauty, motion, depth, normal, reset, sigmas, 1)


def svgf_guides(surf, view, proj, prev_view_proj, frame_index):
    surf = np.ascontiguousarray(surf, np.float32); H, W = surf.shape[:2]
    motion = np.zeros((H, W), np.uint32); depth = np.zeros((H, W), np.float32)
    v = (C.c_float * 16)(*[float(x) for x in view]); pr = (C.c_float * 16)(*[float(x) for x in proj]); pv = (C.c_float * 16)(*[float(x) for x in prev_view_proj])
    lib().emul_svgf_guides(_p(surf), W, H, v, pr, pv, int(frame_index), _p(motion), _p(depth))
    return motion, depth


def set_warp_noise(on: bool):
    """Pseudo-random active-lane counts: drives triangle postponing and pause/resume in the host emulator."""
    lib().emul_set_warp_noise(1 if on else 0)


def trav_stats(reset=True):
    """(node visits, triangle tests) since the last reset."""
    a = C.c_ulonglong(); b = C.c_ulonglong(); lib().emul_trav_stats(C.byref(a), C.byref(b), 1 if reset else 0); return a.value, b.value


def nrd_pack(rad_hd_vz_rough, normal_rough):
    a = np.ascontiguousarray(rad_hd_vz_rough, np.float32); b = np.ascontiguousarray(normal_rough, np.float32); n = len(a)
    pr = np.zeros((n, 4), np.float32); pn = np.zeros((n, 4), np.float32); back = np.zeros((n, 3), np.float32)
    lib().emul_nrd_pack_batch(_p(a), _p(b), n, _p(pr), _p(pn), _p(back))
    return pr, pn, back

"""ctypes binding of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from ohao_engine_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Settings(C.Structure):   # ohb_settings (include/ohao_b200.h)
    _fields_ = [("profile", C.c_uint32), ("max_bounces", C.c_uint32), ("flags", C.c_uint32), ("firefly_clamp_lum", C.c_float),
                ("sampler_type", C.c_uint32), ("anisotropy_strength", C.c_float), ("anisotropy_rotation", C.c_float),
                ("subsurface_strength", C.c_float), ("samples_per_frame", C.c_uint32), ("denoise_mode", C.c_uint32), ("_pad", C.c_uint32 * 2)]


class Counters(C.Structure):   # ohb_counters
    _fields_ = [("samples", C.c_uint64), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("closest_hits", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("_reserved", C.c_uint64 * 3)]


RAY_DTYPE = np.dtype([("origin", "<f4", (3,)), ("tmin", "<f4"), ("dir", "<f4", (3,)), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])


class SceneDesc(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("stride_bytes", C.c_uint64), ("nverts", C.c_uint32),
                ("indices", C.c_void_p), ("ntris", C.c_uint32),
                ("normals", C.c_void_p), ("uvs", C.c_void_p), ("mat_ids", C.c_void_p),
                ("instances", C.c_void_p), ("ninstances", C.c_uint32),
                ("mat_colors", C.c_void_p), ("nmaterials", C.c_uint32),
                ("textures", C.c_void_p), ("tex_w", C.c_uint32), ("tex_h", C.c_uint32), ("tex_layers", C.c_uint32),
                ("light_ssbo", C.c_void_p), ("light_bytes", C.c_uint64),
                ("env", C.c_void_p), ("env_w", C.c_uint32), ("env_h", C.c_uint32)]


class RenderArgs(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("width", C.c_uint32), ("height", C.c_uint32),
                ("first_sample_index", C.c_uint32), ("history_count", C.c_uint32), ("nsamples", C.c_uint32),
                ("tile_x", C.c_uint32), ("tile_y", C.c_uint32), ("tile_w", C.c_uint32), ("tile_h", C.c_uint32),
                ("settings", Settings),
                ("accum", C.c_void_p), ("ldr", C.c_void_p), ("albedo", C.c_void_p), ("normal", C.c_void_p), ("sample_dump", C.c_void_p),
                ("nthreads", C.c_int32), ("counters", Counters)]


class RTArgs(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("prev_view_proj", C.c_float * 16),
                ("width", C.c_uint32), ("height", C.c_uint32), ("frame_index", C.c_uint32), ("history_count", C.c_uint32), ("view_changed", C.c_uint32),
                ("settings", Settings),
                ("accum_prev", C.c_void_p), ("accum_curr", C.c_void_p), ("surf_prev", C.c_void_p), ("surf_curr", C.c_void_p),
                ("shad_prev", C.c_void_p), ("shad_curr", C.c_void_p), ("res_prev", C.c_void_p * 3), ("res_curr", C.c_void_p * 3),
                ("albedo", C.c_void_p), ("normal", C.c_void_p), ("radiance_dump", C.c_void_p), ("gi_dump", C.c_void_p), ("denoised", C.c_void_p), ("ldr", C.c_void_p),
                ("nthreads", C.c_int32), ("counters", Counters)]


class HybridShadowParams(C.Structure):   # ohb_hybrid_shadow_params (include/ohao_b200.h)
    _fields_ = [("light_dir", C.c_float * 3), ("light_radius", C.c_float), ("light_pos", C.c_float * 3), ("light_range", C.c_float),
                ("light_type", C.c_uint32), ("sample_count", C.c_uint32), ("_pad", C.c_uint32 * 2)]


class HybridGiParams(C.Structure):       # ohb_hybrid_gi_params
    _fields_ = [("light_pos", C.c_float * 3), ("light_intensity", C.c_float), ("sample_count", C.c_uint32), ("frame_index", C.c_uint32), ("_pad", C.c_uint32 * 2)]


class SvgfArgs(C.Structure):   # orc_svgf_args (oracle.cpp) == emul_svgf_args (tests/emul/emul.cpp)
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("reset", C.c_int32), ("nthreads", C.c_int32),
                ("sigma_l", C.c_float), ("sigma_normal", C.c_float), ("sigma_depth", C.c_float), ("_pad", C.c_float),
                ("beauty", C.c_void_p), ("motion", C.c_void_p), ("depth", C.c_void_p), ("normal", C.c_void_p),
                ("prev_color", C.c_void_p), ("prev_moments", C.c_void_p), ("prev_geom", C.c_void_p),
                ("cur_color", C.c_void_p), ("cur_moments", C.c_void_p), ("cur_geom", C.c_void_p)]


SVGF_SIGMAS = (0.4, 0.30, 2.0)   # kSigmaL, kSigmaNormal, kSigmaDepth (atrous_denoise.cpp:38-40)


class SvgfState:
    """Host mirror of AtrousDenoiser's persistent history (atrous_denoise.cpp:79-121): colour, moments, geometry as raw
    fp16 bits, ping-ponged; `cur` = the set written by the last dispatch."""

    def __init__(self, width, height):
        self.W, self.H = width, height
        z = lambda: np.zeros((height, width, 4), np.uint16)
        self.color = [z(), z()]; self.moments = [z(), z()]; self.geom = [z(), z()]; self.cur = 0


def _svgf_dispatch(fn, st: SvgfState, beauty, motion, depth, normal, reset, sigmas=SVGF_SIGMAS, nthreads=None):
    """Runs one AtrousDenoiser::dispatch through `fn` (the oracle's or the emulator's entry point); returns the denoised RGBA8."""
    out = np.ascontiguousarray(beauty, np.uint8).copy()
    motion = np.ascontiguousarray(motion, np.uint32); depth = np.ascontiguousarray(depth, np.float32); normal = np.ascontiguousarray(normal, np.float32)
    a = SvgfArgs(); a.width, a.height, a.reset, a.nthreads = st.W, st.H, int(bool(reset)), nthreads or (os.cpu_count() or 1)
    a.sigma_l, a.sigma_normal, a.sigma_depth = sigmas
    p, c = st.cur, 1 - st.cur
    a.beauty, a.motion, a.depth, a.normal = _p(out), _p(motion), _p(depth), _p(normal)
    a.prev_color, a.prev_moments, a.prev_geom = _p(st.color[p]), _p(st.moments[p]), _p(st.geom[p])
    a.cur_color, a.cur_moments, a.cur_geom = _p(st.color[c]), _p(st.moments[c]), _p(st.geom[c])
    fn(C.byref(a))
    st.cur = c
    return out


def svgf_dispatch(st: SvgfState, beauty, motion, depth, normal, reset, sigmas=SVGF_SIGMAS, nthreads=None):
    return _svgf_dispatch(lib().orc_svgf_dispatch, st, beauty, motion, depth, normal, reset, sigmas, nthreads)


def svgf_guides(surf, view, proj, prev_view_proj, frame_index):
    """(motion RG16F bits packed x | y << 16, linear view Z) of a frame from its surface-history plane."""
    surf = np.ascontiguousarray(surf, np.float32); H, W = surf.shape[:2]
    motion = np.zeros((H, W), np.uint32); depth = np.zeros((H, W), np.float32)
    v = (C.c_float * 16)(*[float(x) for x in view]); pr = (C.c_float * 16)(*[float(x) for x in proj]); pv = (C.c_float * 16)(*[float(x) for x in prev_view_proj])
    lib().orc_svgf_guides(_p(surf), W, H, v, pr, pv, int(frame_index), _p(motion), _p(depth))
    return motion, depth


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h", ".inc", ".inl"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/ohao"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_scene_create.restype = C.c_void_p
        _LIB.orc_scene_create.argtypes = [C.POINTER(SceneDesc)]
        _LIB.orc_scene_destroy.argtypes = [C.c_void_p]
        _LIB.orc_render_offline.argtypes = [C.c_void_p, C.POINTER(RenderArgs)]
        _LIB.orc_render_realtime.argtypes = [C.c_void_p, C.POINTER(RTArgs)]
        _LIB.orc_svgf_dispatch.argtypes = [C.POINTER(SvgfArgs)]
        _LIB.orc_svgf_guides.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32, C.c_void_p, C.c_void_p]
        _LIB.orc_f2h.restype = C.c_uint16; _LIB.orc_f2h.argtypes = [C.c_float]
        _LIB.orc_h2f.restype = C.c_float; _LIB.orc_h2f.argtypes = [C.c_uint16]
        _LIB.orc_sampler_1d.restype = C.c_float
        _LIB.orc_sampler_1d.argtypes = [C.c_uint32] * 5
        _LIB.orc_sobol_raw.restype = C.c_float
        _LIB.orc_sobol_raw.argtypes = [C.c_uint32, C.c_uint32]
        _LIB.orc_owen.restype = C.c_uint32
        _LIB.orc_owen.argtypes = [C.c_uint32, C.c_uint32]
        _LIB.orc_sobol_dirs.restype = C.POINTER(C.c_uint32)
        _LIB.orc_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int]
        _LIB.orc_occluded_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _LIB.orc_env_cdf.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        _LIB.orc_scene_env_cdf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        _LIB.orc_env_sample_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        _LIB.orc_env_pdf_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _LIB.orc_scene_set_accel_mode.argtypes = [C.c_void_p, C.c_int]
        _LIB.orc_scene_update_instances.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        _LIB.orc_hybrid_shadow.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB.orc_hybrid_gi.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB.orc_nrd_pack_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.orc_set_ray_recorder.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        _LIB.orc_ray_recorder_count.restype = C.c_uint32
        _LIB.orc_scene_set_materials.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        _LIB.orc_scene_set_lights.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        _LIB.orc_tonemap.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    return _LIB


def ref_lib():
    """The reference's own TUs (oracle/_ref/libohao_ref.so) or None when not built."""
    p = os.path.join(_HERE, "_ref", "libohao_ref.so")
    if not os.path.exists(p):
        return None
    l = C.CDLL(p)
    l.ref_sobol_sample1d.restype = C.c_float
    l.ref_sobol_sample1d.argtypes = [C.c_uint32, C.c_uint32]
    l.ref_owen.restype = C.c_uint32
    l.ref_owen.argtypes = [C.c_uint32, C.c_uint32]
    l.ref_sobol_dirs.restype = C.POINTER(C.c_uint32)
    l.ref_env_cdf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    return l


def offline_settings(max_bounces=4, flags=1, firefly=0.0, sampler=1) -> Settings:
    """kOfflineRTSettings (rt_settings.hpp:49-61): 4 bounces, Sobol, AOVs on, no clamp."""
    s = Settings()
    s.profile = 0; s.max_bounces = max_bounces; s.flags = flags; s.firefly_clamp_lum = firefly
    s.sampler_type = sampler; s.samples_per_frame = 1
    return s


def realtime_settings(max_bounces=2, flags=1 | 2 | 4, firefly=10.0, sampler=1, spf=1) -> Settings:
    """kRealtimeRTSettings (rt_settings.hpp:37-48) with the sampler the pipeline actually runs (Sobol, quirk Q3)."""
    s = Settings()
    s.profile = 1; s.max_bounces = max_bounces; s.flags = flags; s.firefly_clamp_lum = firefly
    s.sampler_type = sampler; s.samples_per_frame = spf
    return s


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RealtimeState:
    """Host mirror of PathTracer's realtime frame state: ping-ponged history images + frame counters."""

    def __init__(self, width, height):
        self.W, self.H = width, height
        z = lambda: np.zeros((height, width, 4), np.float32)
        self.accum = [z(), z()]; self.surf = [z(), z()]; self.shad = [z(), z()]
        self.res = [[z(), z(), z()], [z(), z(), z()]]
        self.albedo = z(); self.normal = z()
        self.cur = 0; self.frame_index = 0; self.history = 0; self.prev_view_proj = np.eye(4, dtype=np.float32).reshape(16)


class OracleScene:
    def __init__(self, ps):
        self.ps = ps
        d = SceneDesc()
        self._keep = [np.ascontiguousarray(x) for x in (ps.positions, ps.indices, ps.normals, ps.uvs, ps.mat_ids, ps.instances,
                                                        ps.mat_colors, ps.textures, ps.light_ssbo)]
        k = self._keep
        d.positions = _p(k[0]); d.stride_bytes = k[0].strides[0]; d.nverts = ps.nverts
        d.indices = _p(k[1]); d.ntris = ps.ntris
        d.normals = _p(k[2]); d.uvs = _p(k[3]); d.mat_ids = _p(k[4])
        d.instances = _p(k[5]); d.ninstances = len(k[5])
        d.mat_colors = _p(k[6]); d.nmaterials = ps.nmaterials
        d.textures = _p(k[7]); d.tex_layers, d.tex_h, d.tex_w = k[7].shape[0], k[7].shape[1], k[7].shape[2]
        d.light_ssbo = _p(k[8]); d.light_bytes = k[8].nbytes
        if ps.env is not None:
            self._env = np.ascontiguousarray(ps.env, np.float32)
            d.env = _p(self._env); d.env_h, d.env_w = self._env.shape[0], self._env.shape[1]
        self.h = lib().orc_scene_create(C.byref(d))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_destroy(self.h); self.h = None

    def set_accel_mode(self, two_level: bool):
        lib().orc_scene_set_accel_mode(self.h, int(two_level))

    def update_instances(self, instances):
        inst = np.ascontiguousarray(instances)
        lib().orc_scene_update_instances(self.h, _p(inst), len(inst))

    def set_materials(self, mat_colors):
        mc = np.ascontiguousarray(mat_colors, np.float32)
        lib().orc_scene_set_materials(self.h, _p(mc), mc.shape[0] // 3)

    def set_lights(self, ssbo):
        b = np.ascontiguousarray(ssbo, np.uint8)
        lib().orc_scene_set_lights(self.h, _p(b), b.nbytes)

    def render_offline(self, view, proj, width, height, nsamples, first_sample=0, history=0, accum=None, settings=None,
                       tile=None, nthreads=None, dump=False, want_ldr=True, want_aov=False):
        a = RenderArgs()
        a.view[:] = [float(x) for x in view]; a.proj[:] = [float(x) for x in proj]
        a.width, a.height = width, height
        a.first_sample_index, a.history_count, a.nsamples = first_sample, history, nsamples
        if tile: a.tile_x, a.tile_y, a.tile_w, a.tile_h = tile
        a.settings = settings or offline_settings()
        if accum is None: accum = np.zeros((height, width, 4), np.float32)
        ldr = np.zeros((height, width, 4), np.uint8) if want_ldr else None
        alb = np.zeros((height, width, 4), np.float32) if want_aov else None
        nrm = np.zeros((height, width, 4), np.float32) if want_aov else None
        sd = np.zeros((nsamples, height, width, 4), np.float32) if dump else None
        a.accum, a.ldr, a.albedo, a.normal, a.sample_dump = _p(accum), _p(ldr), _p(alb), _p(nrm), _p(sd)
        a.nthreads = nthreads or (os.cpu_count() or 1)
        lib().orc_render_offline(self.h, C.byref(a))
        c = a.counters
        return dict(accum=accum, ldr=ldr, albedo=alb, normal=nrm, samples=sd,
                    counters=dict(samples=c.samples, closest_rays=c.closest_rays, shadow_rays=c.shadow_rays, closest_hits=c.closest_hits))

    def render_realtime(self, st: "RealtimeState", view, proj, settings=None, view_changed=False, nthreads=None, dumps=False, fresh=False):
        """One realtime frame; flips the ping-pong buffers of `st` and advances its counters like PathTracer::render.
        fresh = denoiseWantsFreshSample (path_tracer_render.cpp:707-712): the raygen sees historyFrameCount = 0."""
        a = RTArgs()
        a.view[:] = [float(x) for x in view]; a.proj[:] = [float(x) for x in proj]; a.prev_view_proj[:] = [float(x) for x in st.prev_view_proj]
        a.width, a.height, a.frame_index, a.history_count, a.view_changed = st.W, st.H, st.frame_index, (0 if fresh else st.history), int(view_changed)
        a.settings = settings or realtime_settings()
        p, c = st.cur, 1 - st.cur
        a.accum_prev, a.accum_curr = _p(st.accum[p]), _p(st.accum[c]); a.surf_prev, a.surf_curr = _p(st.surf[p]), _p(st.surf[c])
        a.shad_prev, a.shad_curr = _p(st.shad[p]), _p(st.shad[c])
        for k in range(3):
            a.res_prev[k] = _p(st.res[p][k]); a.res_curr[k] = _p(st.res[c][k])
        a.albedo, a.normal = _p(st.albedo), _p(st.normal)
        rad = np.zeros((st.H, st.W, 4), np.float32) if dumps else None; gi = np.zeros((st.H, st.W, 4), np.float32) if dumps else None
        den = np.zeros((st.H, st.W, 4), np.float32); ldr = np.zeros((st.H, st.W, 4), np.uint8)
        a.radiance_dump, a.gi_dump, a.denoised, a.ldr = _p(rad), _p(gi), _p(den), _p(ldr)
        a.nthreads = nthreads or (os.cpu_count() or 1)
        lib().orc_render_realtime(self.h, C.byref(a))
        st.cur = c; st.frame_index += 1; st.history += 1
        v = np.asarray(view, np.float32).reshape(4, 4); pr = np.asarray(proj, np.float32).reshape(4, 4)    # [col][row]
        st.prev_view_proj = (v @ pr).reshape(16).astype(np.float32)        # column-major proj*view == (row-vector) view@proj
        cnt = a.counters
        return dict(accum=st.accum[c], ldr=ldr, denoised=den, radiance=rad, gi=gi, reservoirs=st.res[c], surf=st.surf[c], shad=st.shad[c],
                    counters=dict(samples=cnt.samples, closest_rays=cnt.closest_rays, shadow_rays=cnt.shadow_rays, closest_hits=cnt.closest_hits))

    def trace(self, rays, brute=False, nthreads=None):
        rays = np.ascontiguousarray(rays, RAY_DTYPE); hits = np.zeros(len(rays), HIT_DTYPE)
        lib().orc_trace_batch(self.h, _p(rays), len(rays), _p(hits), int(brute), nthreads or (os.cpu_count() or 1))
        return hits

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays, RAY_DTYPE); occ = np.zeros(len(rays), np.uint8)
        lib().orc_occluded_batch(self.h, _p(rays), len(rays), _p(occ))
        return occ

    def env_cdf(self):
        h, w = self.ps.env.shape[:2]
        marg = np.zeros(h, np.float32); cond = np.zeros((h, w), np.float32); I = C.c_float()
        lib().orc_scene_env_cdf(self.h, _p(marg), _p(cond), C.byref(I))
        return marg, cond, I.value

    def env_sample(self, u12):
        u = np.ascontiguousarray(u12, np.float32); n = len(u)
        dp = np.zeros((n, 4), np.float32); pd = np.zeros(n, np.float32)
        lib().orc_env_sample_batch(self.h, _p(u), n, _p(dp), _p(pd))
        return dp, pd

    def env_pdf(self, dirs):
        d = np.ascontiguousarray(dirs, np.float32); pd = np.zeros(len(d), np.float32)
        lib().orc_env_pdf_batch(self.h, _p(d), len(d), _p(pd))
        return pd

    def hybrid_shadow(self, W, H, gpos, gnrm, params: HybridShadowParams, nthreads=None):
        gpos = np.ascontiguousarray(gpos, np.float32); gnrm = np.ascontiguousarray(gnrm, np.float32); mask = np.zeros((H, W), np.uint8)
        lib().orc_hybrid_shadow(self.h, W, H, _p(gpos), _p(gnrm), C.byref(params), _p(mask), nthreads or (os.cpu_count() or 1))
        return mask

    def hybrid_gi(self, W, H, gpos, gnrm, galbedo, history, inst_mat, params: HybridGiParams, nthreads=None):
        a = [np.ascontiguousarray(x, np.float32) for x in (gpos, gnrm, galbedo, history, inst_mat)]; out = np.zeros((H, W, 4), np.uint16)
        lib().orc_hybrid_gi(self.h, W, H, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), C.byref(params), _p(out), nthreads or (os.cpu_count() or 1))
        return out

    def record_rays(self, view, proj, width, height, nsamples, tile, cap=1 << 20, settings=None, first_sample=0):
        rays = np.zeros(cap, RAY_DTYPE); hits = np.zeros(cap, HIT_DTYPE); kinds = np.zeros(cap, np.uint8)
        lib().orc_set_ray_recorder(_p(rays), _p(hits), _p(kinds), cap)
        try:
            self.render_offline(view, proj, width, height, nsamples, first_sample=first_sample, tile=tile, nthreads=1, settings=settings, want_ldr=False)
            n = lib().orc_ray_recorder_count()
        finally:
            lib().orc_set_ray_recorder(None, None, None, 0)
        return rays[:n].copy(), hits[:n].copy(), kinds[:n].copy()


def nrd_pack(rad_hd_vz_rough, normal_rough):
    a = np.ascontiguousarray(rad_hd_vz_rough, np.float32); b = np.ascontiguousarray(normal_rough, np.float32); n = len(a)
    pr = np.zeros((n, 4), np.float32); pn = np.zeros((n, 4), np.float32); back = np.zeros((n, 3), np.float32)
    lib().orc_nrd_pack_batch(_p(a), _p(b), n, _p(pr), _p(pn), _p(back))
    return pr, pn, back


def env_cdf(rgba):
    rgba = np.ascontiguousarray(rgba, np.float32); h, w = rgba.shape[:2]
    marg = np.zeros(h, np.float32); cond = np.zeros((h, w), np.float32); I = C.c_float()
    lib().orc_env_cdf(_p(rgba), w, h, _p(marg), _p(cond), C.byref(I))
    return marg, cond, I.value

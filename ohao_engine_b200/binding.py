"""ctypes binding of libohao_b200.so + a Python mirror of the reference's image-level seam.

``Renderer`` keeps the method names of ``ohao::VulkanRenderer`` that the examples touch
(ohao/gpu/vulkan/renderer.hpp:114-300): set_scene / update_scene_buffers, set_render_seed,
reset_accumulation, render, get_pixels, readback_hdr_buffers, update_rt_material_params,
update_rt_light_params.  Everything goes through the C ABI of include/ohao_b200.h; there is no
fallback: importing works without a GPU, creating a Renderer does not.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import build as _build

PROFILE_OFFLINE, PROFILE_REALTIME = 0, 1
DENOISE_NONE, DENOISE_ATROUS = 0, 4      # DenoiseMode values (denoise_types.hpp:13-19)
FLAG_AOVS, FLAG_INTERNAL_DENOISE, FLAG_FIREFLY = 1, 2, 4
FLAG_GOLDEN_COMPAT = 1 << 16
ACCEL_FLATTEN, ACCEL_TWO_LEVEL = 0, 1
SAMPLER_PCG, SAMPLER_SOBOL = 0, 1             # GLSL SAMPLER_PCG / SAMPLER_SOBOL (sampler_api.glsl:12-13)

RAY_DTYPE = np.dtype([("origin", "<f4", (3,)), ("tmin", "<f4"), ("dir", "<f4", (3,)), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])


class Settings(C.Structure):
    _fields_ = [("profile", C.c_uint32), ("max_bounces", C.c_uint32), ("flags", C.c_uint32), ("firefly_clamp_lum", C.c_float),
                ("sampler_type", C.c_uint32), ("anisotropy_strength", C.c_float), ("anisotropy_rotation", C.c_float),
                ("subsurface_strength", C.c_float), ("samples_per_frame", C.c_uint32), ("denoise_mode", C.c_uint32), ("_pad", C.c_uint32 * 2)]


class Counters(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("closest_hits", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("_reserved", C.c_uint64 * 3)]


class HybridShadowParams(C.Structure):   # ohb_hybrid_shadow_params
    _fields_ = [("light_dir", C.c_float * 3), ("light_radius", C.c_float), ("light_pos", C.c_float * 3), ("light_range", C.c_float),
                ("light_type", C.c_uint32), ("sample_count", C.c_uint32), ("_pad", C.c_uint32 * 2)]


class HybridGiParams(C.Structure):       # ohb_hybrid_gi_params
    _fields_ = [("light_pos", C.c_float * 3), ("light_intensity", C.c_float), ("sample_count", C.c_uint32), ("frame_index", C.c_uint32), ("_pad", C.c_uint32 * 2)]


class AccelStats(C.Structure):
    _fields_ = [("num_tris", C.c_uint32), ("num_nodes", C.c_uint32), ("levels", C.c_uint32), ("max_leaf_tris", C.c_uint32),
                ("sah_cost", C.c_float), ("build_ms", C.c_float), ("treelet_passes", C.c_uint32), ("update_ms", C.c_float)]


# every symbol include/ohao_b200.h declares: name -> (restype, argtypes)
_VP, _U32, _I, _F = C.c_void_p, C.c_uint32, C.c_int, C.c_float
ABI = {
    "ohb_abi_version": (_U32, []),
    "ohb_create": (_VP, [_I, _U32, _U32, _I]),
    "ohb_destroy": (None, [_VP]),
    "ohb_resize": (_I, [_VP, _U32, _U32]),
    "ohb_last_error": (C.c_char_p, [_VP]),
    "ohb_set_geometry": (_I, [_VP, _VP, C.c_size_t, _U32, _VP, _U32, _VP, _VP, _VP]),
    "ohb_set_instances": (_I, [_VP, _VP, _U32]),
    "ohb_set_materials": (_I, [_VP, _VP, _U32]),
    "ohb_set_textures": (_I, [_VP, _VP, _U32, _U32, _U32]),
    "ohb_set_lights": (_I, [_VP, _VP, C.c_size_t]),
    "ohb_set_env": (_I, [_VP, _VP, _U32, _U32]),
    "ohb_get_env_cdf": (_I, [_VP, _VP, _VP, C.POINTER(_F)]),
    "ohb_env_sample_batch": (_I, [_VP, _VP, _U32, _VP, _VP]),
    "ohb_env_pdf_batch": (_I, [_VP, _VP, _U32, _VP]),
    "ohb_build_accel": (_I, [_VP]),
    "ohb_get_accel_stats": (_I, [_VP, C.POINTER(AccelStats)]),
    "ohb_set_settings": (_I, [_VP, C.POINTER(Settings)]),
    "ohb_get_settings": (_I, [_VP, C.POINTER(Settings)]),
    "ohb_set_seed": (None, [_VP, _U32]),
    "ohb_reset_accumulation": (None, [_VP]),
    "ohb_notify_view_changed": (None, [_VP]),
    "ohb_frame_index": (_U32, [_VP]),
    "ohb_render": (_I, [_VP, _VP, _VP, _U32]),
    "ohb_set_tile": (_I, [_VP, _U32, _U32, _U32, _U32]),
    "ohb_read_ldr": (_I, [_VP, _VP]),
    "ohb_read_hdr": (_I, [_VP, _VP, _VP, _VP]),
    "ohb_synchronize": (_I, [_VP]),
    "ohb_accum_dev_ptr": (_VP, [_VP, C.POINTER(C.c_size_t)]),
    "ohb_set_accum_mode": (_I, [_VP, _I]),
    "ohb_resolve": (_I, [_VP]), "ohb_clear_accum": (_I, [_VP]),
    "ohb_set_accel_mode": (_I, [_VP, C.c_int]), "ohb_update_instances": (_I, [_VP, _VP, C.c_uint32]),
    "ohb_hybrid_shadow": (_I, [_VP, _VP, _VP, _VP, _VP]), "ohb_hybrid_gi": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, C.c_uint32, _VP, _VP]),
    "ohb_nrd_pack_batch": (_I, [_VP, _VP, _VP, C.c_uint32, _VP, _VP, _VP]),
    "ohb_trace_batch": (_I, [_VP, _VP, _U32, _VP]),
    "ohb_occluded_batch": (_I, [_VP, _VP, _U32, _VP]),
    "ohb_set_sample_dump": (_I, [_VP, _VP, C.c_size_t]),
    "ohb_get_counters": (_I, [_VP, C.POINTER(Counters)]),
    "ohb_reset_counters": (None, [_VP]),
    "ohb_get_timing": (_I, [_VP, C.POINTER(_F), C.POINTER(_F), C.POINTER(_F)]),
    "ohb_get_timing_detail": (_I, [_VP, C.POINTER(_F * 8), C.POINTER(C.c_uint64 * 8)]),
    "ohb_enable_timing": (_I, [_VP, _I]),
    "ohb_set_realtime_dump": (_I, [_VP, _VP, _VP, _VP]),
    "ohb_read_realtime_state": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "ohb_svgf_dispatch": (_I, [_VP, _VP, _VP, _VP, _VP, _I]),
    "ohb_read_denoise_state": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "ohb_timer_start": (_I, [_VP]),
    "ohb_timer_stop": (_I, [_VP, C.POINTER(_F)]),
}

_LIB: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Load libohao_b200.so and bind every ABI symbol.  Raises if the library is missing."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("OHAO_B200_LIB") or _build.LIB_PATH      # override: A/B runs of two builds of the same ABI
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the CUDA extension is mandatory, there is no CPU fallback)")
        lib = C.CDLL(path)
        for name, (res, args) in ABI.items():
            fn = getattr(lib, name)
            fn.restype = res; fn.argtypes = args
        _LIB = lib
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OhbError(RuntimeError):
    pass


class Renderer:
    """Python mirror of VulkanRenderer's RT seam on top of one ohb_ctx (one per GPU)."""

    def __init__(self, width: int, height: int, profile: int = PROFILE_OFFLINE, device: int = 0):
        self.lib = load_library()
        self.width, self.height, self.profile = width, height, profile
        self.h = self.lib.ohb_create(device, width, height, profile)
        if not self.h:
            raise OhbError("ohb_create failed: " + self.lib.ohb_last_error(None).decode())
        self._dump = None
        self.scene = None

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc, what):
        if rc != 0:
            raise OhbError(f"{what}: {self.lib.ohb_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.ohb_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene: setScene / updateSceneBuffers / buildAccelerationStructures --------------------
    def set_scene(self, ps, build: bool = True):
        self.scene = ps
        pos = np.ascontiguousarray(ps.positions, np.float32)
        self._ck(self.lib.ohb_set_geometry(self.h, _p(pos), pos.strides[0], ps.nverts, _p(np.ascontiguousarray(ps.indices, np.uint32)), ps.ntris,
                                           _p(np.ascontiguousarray(ps.normals, np.float32)), _p(np.ascontiguousarray(ps.uvs, np.float32)),
                                           _p(np.ascontiguousarray(ps.mat_ids, np.uint32))), "ohb_set_geometry")
        inst = np.ascontiguousarray(ps.instances)
        self._ck(self.lib.ohb_set_instances(self.h, _p(inst), len(inst)), "ohb_set_instances")
        self.update_rt_material_params(ps.mat_colors)
        tex = np.ascontiguousarray(ps.textures, np.uint8)
        self._ck(self.lib.ohb_set_textures(self.h, _p(tex), tex.shape[2], tex.shape[1], tex.shape[0]), "ohb_set_textures")
        self.update_rt_light_params(ps.light_ssbo)
        self.set_environment_map(ps.env)
        if build:
            self.build_accel()

    update_scene_buffers = set_scene

    def update_rt_material_params(self, mat_colors):
        mc = np.ascontiguousarray(mat_colors, np.float32)
        self._ck(self.lib.ohb_set_materials(self.h, _p(mc), mc.size // 12), "ohb_set_materials")

    def update_rt_light_params(self, light_ssbo):
        b = np.ascontiguousarray(light_ssbo, np.uint8)
        self._ck(self.lib.ohb_set_lights(self.h, _p(b), b.nbytes), "ohb_set_lights")

    def set_environment_map(self, env):
        if env is None:
            self._ck(self.lib.ohb_set_env(self.h, None, 0, 0), "ohb_set_env")
        else:
            e = np.ascontiguousarray(env, np.float32)
            self._ck(self.lib.ohb_set_env(self.h, _p(e), e.shape[1], e.shape[0]), "ohb_set_env")

    def set_accel_mode(self, two_level: bool):
        """OHB_ACCEL_FLATTEN (default) or OHB_ACCEL_TWO_LEVEL (per-instance BLAS + TLAS); invalidates the structure."""
        self._ck(self.lib.ohb_set_accel_mode(self.h, ACCEL_TWO_LEVEL if two_level else ACCEL_FLATTEN), "ohb_set_accel_mode")

    def update_instances(self, instances) -> AccelStats:
        """MODE_UPDATE: same instances, new transforms (TLAS refit, or re-transform + refit of the flattened tree)."""
        inst = np.ascontiguousarray(instances)
        self._ck(self.lib.ohb_update_instances(self.h, _p(inst), len(inst)), "ohb_update_instances")
        return self.accel_stats()

    def build_accel(self) -> AccelStats:
        self._ck(self.lib.ohb_build_accel(self.h), "ohb_build_accel")
        return self.accel_stats()

    def accel_stats(self) -> AccelStats:
        s = AccelStats(); self._ck(self.lib.ohb_get_accel_stats(self.h, C.byref(s)), "ohb_get_accel_stats"); return s

    # -- settings / accumulation --------------------------------------------------------------
    def get_settings(self) -> Settings:
        s = Settings(); self._ck(self.lib.ohb_get_settings(self.h, C.byref(s)), "ohb_get_settings"); return s

    def set_rt_render_settings(self, s: Settings):
        self._ck(self.lib.ohb_set_settings(self.h, C.byref(s)), "ohb_set_settings")

    def set_render_seed(self, seed: int): self.lib.ohb_set_seed(self.h, seed)
    def reset_accumulation(self): self.lib.ohb_reset_accumulation(self.h)
    def notify_camera_changed(self): self.lib.ohb_notify_view_changed(self.h)
    def frame_index(self) -> int: return self.lib.ohb_frame_index(self.h)
    def set_tile(self, x0, y0, w, h): self._ck(self.lib.ohb_set_tile(self.h, x0, y0, w, h), "ohb_set_tile")
    def set_accum_mode(self, sum_mode: bool): self._ck(self.lib.ohb_set_accum_mode(self.h, int(sum_mode)), "ohb_set_accum_mode")
    def resolve(self): self._ck(self.lib.ohb_resolve(self.h), "ohb_resolve")
    def clear_accum(self): self._ck(self.lib.ohb_clear_accum(self.h), "ohb_clear_accum")
    def resize(self, w, h):
        self._ck(self.lib.ohb_resize(self.h, w, h), "ohb_resize"); self.width, self.height = w, h

    # -- render / readback ----------------------------------------------------------------------
    def render(self, view, proj, nsamples: int = 1, dump: bool = False):
        """`nsamples` consecutive VulkanRenderer::render() calls (offline: one spp each)."""
        v = np.ascontiguousarray(view, np.float32); p = np.ascontiguousarray(proj, np.float32)
        if dump:
            self._dump = np.zeros((nsamples, self.height, self.width, 4), np.float32)
            self._ck(self.lib.ohb_set_sample_dump(self.h, _p(self._dump), self._dump.size), "ohb_set_sample_dump")
        try:
            self._ck(self.lib.ohb_render(self.h, _p(v), _p(p), nsamples), "ohb_render")
        finally:
            if dump:
                self.lib.ohb_set_sample_dump(self.h, None, 0)
        return self._dump if dump else None

    def synchronize(self): self._ck(self.lib.ohb_synchronize(self.h), "ohb_synchronize")

    def get_pixels(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """getPixelSpan(): RGBA8, top row first."""
        if out is None: out = np.empty((self.height, self.width, 4), np.uint8)
        self._ck(self.lib.ohb_read_ldr(self.h, _p(out)), "ohb_read_ldr"); return out

    def readback_hdr_buffers(self, want_aov: bool = True):
        acc = np.empty((self.height, self.width, 4), np.float32)
        alb = np.empty_like(acc) if want_aov else None; nrm = np.empty_like(acc) if want_aov else None
        self._ck(self.lib.ohb_read_hdr(self.h, _p(acc), _p(alb), _p(nrm)), "ohb_read_hdr")
        return acc, alb, nrm

    def accum_dev_ptr(self):
        n = C.c_size_t(); ptr = self.lib.ohb_accum_dev_ptr(self.h, C.byref(n)); return ptr, n.value

    # -- parity / measurement hooks ---------------------------------------------------------------
    def trace(self, rays) -> np.ndarray:
        rays = np.ascontiguousarray(rays, RAY_DTYPE); hits = np.zeros(len(rays), HIT_DTYPE)
        self._ck(self.lib.ohb_trace_batch(self.h, _p(rays), len(rays), _p(hits)), "ohb_trace_batch"); return hits

    def occluded(self, rays) -> np.ndarray:
        rays = np.ascontiguousarray(rays, RAY_DTYPE); occ = np.zeros(len(rays), np.uint8)
        self._ck(self.lib.ohb_occluded_batch(self.h, _p(rays), len(rays), _p(occ)), "ohb_occluded_batch"); return occ

    def env_cdf(self):
        h, w = self.scene.env.shape[:2]
        marg = np.zeros(h, np.float32); cond = np.zeros((h, w), np.float32); I = C.c_float()
        self._ck(self.lib.ohb_get_env_cdf(self.h, _p(marg), _p(cond), C.byref(I)), "ohb_get_env_cdf"); return marg, cond, I.value

    def env_sample(self, u12):
        u = np.ascontiguousarray(u12, np.float32); n = len(u); dp = np.zeros((n, 4), np.float32); pd = np.zeros(n, np.float32)
        self._ck(self.lib.ohb_env_sample_batch(self.h, _p(u), n, _p(dp), _p(pd)), "ohb_env_sample_batch"); return dp, pd

    def env_pdf(self, dirs):
        d = np.ascontiguousarray(dirs, np.float32); pd = np.zeros(len(d), np.float32)
        self._ck(self.lib.ohb_env_pdf_batch(self.h, _p(d), len(d), _p(pd)), "ohb_env_pdf_batch"); return pd

    def hybrid_shadow(self, gpos, gnrm, params) -> np.ndarray:
        """RTShadowTechnique (rt_shadow.rgen): R8 shadow mask from the G-buffer; params = HybridShadowParams (or any ctypes struct of that layout)."""
        gpos = np.ascontiguousarray(gpos, np.float32); gnrm = np.ascontiguousarray(gnrm, np.float32); mask = np.zeros((self.height, self.width), np.uint8)
        self._ck(self.lib.ohb_hybrid_shadow(self.h, _p(gpos), _p(gnrm), C.byref(params), _p(mask)), "ohb_hybrid_shadow"); return mask

    def hybrid_gi(self, gpos, gnrm, galbedo, history, inst_mat, params) -> np.ndarray:
        """RTGITechnique (rt_gi.rgen): RGBA16F one-bounce GI (fp16 bit patterns)."""
        a = [np.ascontiguousarray(x, np.float32) for x in (gpos, gnrm, galbedo, history, inst_mat)]; out = np.zeros((self.height, self.width, 4), np.uint16)
        self._ck(self.lib.ohb_hybrid_gi(self.h, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), len(a[4]), C.byref(params), _p(out)), "ohb_hybrid_gi"); return out

    def nrd_pack(self, rad_hd_vz_rough, normal_rough):
        a = np.ascontiguousarray(rad_hd_vz_rough, np.float32); b = np.ascontiguousarray(normal_rough, np.float32); n = len(a)
        pr = np.zeros((n, 4), np.float32); pn = np.zeros((n, 4), np.float32); back = np.zeros((n, 3), np.float32)
        self._ck(self.lib.ohb_nrd_pack_batch(self.h, _p(a), _p(b), n, _p(pr), _p(pn), _p(back)), "ohb_nrd_pack_batch")
        return pr, pn, back

    def counters(self) -> dict:
        c = Counters(); self._ck(self.lib.ohb_get_counters(self.h, C.byref(c)), "ohb_get_counters")
        return dict(samples=c.samples, closest_rays=c.closest_rays, shadow_rays=c.shadow_rays, closest_hits=c.closest_hits, kernel_launches=c.kernel_launches)

    def reset_counters(self): self.lib.ohb_reset_counters(self.h)
    def enable_timing(self, on: bool = True): self._ck(self.lib.ohb_enable_timing(self.h, int(on)), "ohb_enable_timing")

    def render_realtime(self, view, proj, dumps: bool = False):
        """One realtime frame (profile REALTIME).  With dumps=True also returns the parity dumps."""
        v = np.ascontiguousarray(view, np.float32); p = np.ascontiguousarray(proj, np.float32)
        out = {}
        if dumps:
            out = {k: np.zeros((self.height, self.width, 4), np.float32) for k in ("radiance", "gi", "denoised")}
            self._ck(self.lib.ohb_set_realtime_dump(self.h, _p(out["radiance"]), _p(out["gi"]), _p(out["denoised"])), "ohb_set_realtime_dump")
        try:
            self._ck(self.lib.ohb_render(self.h, _p(v), _p(p), 1), "ohb_render")
        finally:
            if dumps:
                self.lib.ohb_set_realtime_dump(self.h, None, None, None)
        return out

    def realtime_state(self):
        z = lambda: np.zeros((self.height, self.width, 4), np.float32)
        r0, r1, r2, su, sh = z(), z(), z(), z(), z()
        self._ck(self.lib.ohb_read_realtime_state(self.h, _p(r0), _p(r1), _p(r2), _p(su), _p(sh)), "ohb_read_realtime_state")
        return dict(reservoirs=[r0, r1, r2], surf=su, shad=sh)

    def svgf_dispatch(self, beauty, normal, depth, motion, reset: bool) -> np.ndarray:
        """AtrousDenoiser::dispatch on host images (RGBA8, RGBA32F N*0.5+0.5, R32F view Z, RG16F bits): returns the denoised RGBA8."""
        out = np.ascontiguousarray(beauty, np.uint8).copy()
        n = np.ascontiguousarray(normal, np.float32); d = np.ascontiguousarray(depth, np.float32); m = np.ascontiguousarray(motion, np.uint32)
        assert out.shape == (self.height, self.width, 4) and n.shape == (self.height, self.width, 4) and d.shape == (self.height, self.width) and m.shape == (self.height, self.width)
        self._ck(self.lib.ohb_svgf_dispatch(self.h, _p(out), _p(n), _p(d), _p(m), int(bool(reset))), "ohb_svgf_dispatch")
        return out

    def read_denoise_state(self) -> dict:
        """History written by the last SVGF dispatch (raw fp16 bits) and the guide AOVs of the last frame."""
        z = lambda: np.zeros((self.height, self.width, 4), np.uint16)
        col, mom, geo = z(), z(), z(); mot = np.zeros((self.height, self.width), np.uint32); dep = np.zeros((self.height, self.width), np.float32)
        self._ck(self.lib.ohb_read_denoise_state(self.h, _p(col), _p(mom), _p(geo), _p(mot), _p(dep)), "ohb_read_denoise_state")
        return dict(color=col, moments=mom, geom=geo, motion=mot, depth=dep)

    def timer_start(self): self._ck(self.lib.ohb_timer_start(self.h), "ohb_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float(); self._ck(self.lib.ohb_timer_stop(self.h, C.byref(ms)), "ohb_timer_stop"); return float(ms.value)

    def timing(self) -> dict:
        ms = (C.c_float * 8)(); cnt = (C.c_uint64 * 8)()
        self._ck(self.lib.ohb_get_timing_detail(self.h, C.byref(ms), C.byref(cnt)), "ohb_get_timing_detail")
        names = ("trace_closest", "bounce", "trace_shadow", "film", "surface", "rt_pixel", "svgf", "sort_hits")
        return {n: dict(ms=float(ms[i]), launches=int(cnt[i])) for i, n in enumerate(names)}

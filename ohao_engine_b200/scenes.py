"""Host-side scene packing — the data contract of SURVEY.md §3.2 as numpy arrays.

Mirrors what VulkanRenderer::updateSceneBuffers / buildAccelerationStructures hand to the
RT profile renderer (ohao/gpu/vulkan/scene_upload.cpp:21-120, rt_build.cpp:27-934,
light_upload.cpp:153-293): one shared vertex buffer (100-B ``Vertex`` records, position at
byte 0), GLOBAL uint32 indices, vec4 normals, vec2 uvs, per-triangle material ids, 3 vec4 per
material, one R8G8B8A8_UNORM texture array, the 16-B-header light SSBO, and one TLAS instance
per actor with ``customIndex`` = running triangle offset.

These builders synthesise INPUTS (scene description); nothing here is on the timed path.
"""
from __future__ import annotations

import dataclasses
import math
import struct
from typing import List, Optional

import numpy as np

f32 = np.float32
NO_TEX = 0xFFFFFFFF
VERTEX_STRIDE = 100  # sizeof(ohao::Vertex), scene/asset/model.hpp:17-35

INSTANCE_DTYPE = np.dtype(
    [("first_tri", "<u4"), ("tri_count", "<u4"), ("xform", "<f4", (12,)), ("mask", "<u4"), ("_pad", "<u4")]
)
assert INSTANCE_DTYPE.itemsize == 64


# ------------------------------------------------------------------------------------------
# glm-compatible fp32 helpers (column-major 4x4 stored as numpy [col][row] flattened to 16)
# ------------------------------------------------------------------------------------------
def _n(v):
    v = np.asarray(v, dtype=f32)
    return (v / f32(np.sqrt(np.dot(v, v), dtype=f32))).astype(f32)


def look_at(eye, center, up) -> np.ndarray:
    """glm::lookAt (RH).  Returns float32[16] column-major."""
    eye = np.asarray(eye, f32); center = np.asarray(center, f32); up = np.asarray(up, f32)
    f = _n(center - eye)
    s = _n(np.cross(f, up).astype(f32))
    u = np.cross(s, f).astype(f32)
    m = np.zeros((4, 4), f32)  # m[col][row]
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0] = -np.dot(s, eye); m[3][1] = -np.dot(u, eye); m[3][2] = np.dot(f, eye)
    m[3][3] = 1.0
    return m.reshape(16).copy()


def perspective(fovy_rad, aspect, z_near, z_far) -> np.ndarray:
    """glm::perspective (RH, depth -1..1).  float32[16] column-major."""
    t = f32(math.tan(float(f32(fovy_rad)) / 2.0))
    m = np.zeros((4, 4), f32)
    m[0][0] = f32(1.0) / (f32(aspect) * t)
    m[1][1] = f32(1.0) / t
    m[2][2] = -(f32(z_far) + f32(z_near)) / (f32(z_far) - f32(z_near))
    m[2][3] = -1.0
    m[3][2] = -(f32(2.0) * f32(z_far) * f32(z_near)) / (f32(z_far) - f32(z_near))
    return m.reshape(16).copy()


@dataclasses.dataclass
class Camera:
    """ohao::Camera (render/camera/camera.cpp:25-48): yaw/pitch -> front/right/up -> lookAt."""
    position: tuple = (0.0, 0.0, 2.5)
    yaw: float = -90.0
    pitch: float = 0.0
    fov: float = 45.0

    def view(self) -> np.ndarray:
        yaw = f32(np.radians(f32(self.yaw))); pitch = f32(np.radians(f32(self.pitch)))
        front = _n([np.cos(yaw, dtype=f32) * np.cos(pitch, dtype=f32), np.sin(pitch, dtype=f32),
                    np.sin(yaw, dtype=f32) * np.cos(pitch, dtype=f32)])
        right = _n(np.cross(front, np.array([0, 1, 0], f32)).astype(f32))
        up = _n(np.cross(right, front).astype(f32))
        pos = np.asarray(self.position, f32)
        return look_at(pos, pos + front, up)

    def proj(self, width: int, height: int) -> np.ndarray:
        # renderRTPipeline: glm::perspective(radians(fov), W/H, 0.1, 1000)  (render_dispatch.cpp:158-163)
        return perspective(f32(np.radians(f32(self.fov))), f32(width) / f32(height), 0.1, 1000.0)


def trs(position=(0, 0, 0), scale=(1, 1, 1)) -> np.ndarray:
    """TransformComponent local matrix T*R*S with identity rotation -> 3x4 ROW-major float32[12]."""
    m = np.zeros((3, 4), f32)
    m[0, 0], m[1, 1], m[2, 2] = scale
    m[0, 3], m[1, 3], m[2, 3] = position
    return m.reshape(12)


# ------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Mesh:
    positions: np.ndarray            # (n,3) f32
    normals: np.ndarray              # (n,3) f32
    uvs: np.ndarray                  # (n,2) f32
    indices: np.ndarray              # (3t,) u32, mesh-local
    xform: np.ndarray = dataclasses.field(default_factory=trs)   # row-major 3x4
    base_color: tuple = (0.8, 0.8, 0.8)
    roughness: float = 0.5
    metallic: float = 0.0
    # optional per-mesh textures: RGBA8 arrays (h,w,4) or None
    albedo_tex: Optional[np.ndarray] = None
    normal_tex: Optional[np.ndarray] = None
    rough_metal_tex: Optional[np.ndarray] = None
    emissive_tex: Optional[np.ndarray] = None
    name: str = "mesh"


@dataclasses.dataclass
class Light:
    position: tuple
    type: int = 0                    # LightType: 0 sphere, 1 directional, 2 spot, 3 area-rect
    color: tuple = (1.0, 1.0, 1.0)
    intensity: float = 1.0
    radius: float = 0.5
    direction: tuple = (0.0, -1.0, 0.0)
    inner_cone: float = 30.0
    outer_cone: float = 45.0
    edge1: tuple = (1.0, 0.0, 0.0)
    edge2: tuple = (0.0, 0.0, 1.0)

    def pack(self) -> bytes:
        """GPULight, 80 B (render/rt/gpu_light.hpp:10-17; filled at light_upload.cpp:157-181)."""
        dir_param = self.inner_cone if self.type == 2 else self.radius
        if self.type == 3:
            e1 = np.asarray(self.edge1, f32); e2 = np.asarray(self.edge2, f32)
            c = np.cross(e1, e2).astype(f32)
            area = float(np.sqrt(np.dot(c, c), dtype=f32))
            extra = (*self.edge1, 0.0); extra2 = (*self.edge2, area)
        else:
            extra = (0.0, 0.0, 0.0, self.outer_cone); extra2 = (0.0, 0.0, 0.0, 0.0)
        vals = (*self.position, float(self.type), *self.color, self.intensity, *self.direction, dir_param, *extra, *extra2)
        return struct.pack("<20f", *vals)


@dataclasses.dataclass
class PackedScene:
    positions: np.ndarray      # (nverts, 25) f32  == 100-B Vertex records, position at [:, 0:3]
    indices: np.ndarray        # (ntris*3,) u32 global
    normals: np.ndarray        # (nverts, 4) f32
    uvs: np.ndarray            # (nverts, 2) f32
    mat_ids: np.ndarray        # (ntris,) u32
    instances: np.ndarray      # INSTANCE_DTYPE
    mat_colors: np.ndarray     # (nmat*3, 4) f32
    textures: np.ndarray       # (layers, h, w, 4) u8
    light_ssbo: np.ndarray     # u8 bytes
    env: Optional[np.ndarray]  # (h, w, 4) f32 or None
    name: str = "scene"

    @property
    def ntris(self) -> int: return int(self.mat_ids.shape[0])
    @property
    def nverts(self) -> int: return int(self.positions.shape[0])
    @property
    def nmaterials(self) -> int: return int(self.mat_colors.shape[0] // 3)
    @property
    def nlights(self) -> int: return int(np.frombuffer(self.light_ssbo[:4].tobytes(), "<u4")[0])


def _linear_to_srgb8(v: float) -> int:
    # rt_build.cpp:462-465 — the 1x1 "solid colour" layer every untextured material receives (quirk Q1)
    v = f32(v)
    s = v * f32(12.92) if v <= f32(0.0031308) else f32(1.055) * f32(math.pow(float(v), 1.0 / 2.4)) - f32(0.055)
    s = min(max(float(s), 0.0), 1.0)
    return int(f32(s) * f32(255.0) + f32(0.5))


def resize_rgba8_bilinear(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """resizeRGBA8Bilinear (rt_build.cpp:548-590): centre-aligned, clamp, truncating store."""
    sh, sw = src.shape[:2]
    if (sw, sh) == (dw, dh):
        return src.copy()
    sx = f32(sw) / f32(dw); sy = f32(sh) / f32(dh)
    ys = np.clip((np.arange(dh, dtype=f32) + f32(0.5)) * sy - f32(0.5), 0, sh - 1).astype(f32)
    xs = np.clip((np.arange(dw, dtype=f32) + f32(0.5)) * sx - f32(0.5), 0, sw - 1).astype(f32)
    y0 = np.floor(ys).astype(np.int64); x0 = np.floor(xs).astype(np.int64)
    y1 = np.minimum(y0 + 1, sh - 1); x1 = np.minimum(x0 + 1, sw - 1)
    fy = (ys - y0.astype(f32))[:, None, None]; fx = (xs - x0.astype(f32))[None, :, None]
    s = src.astype(f32)
    c00 = s[y0][:, x0]; c10 = s[y0][:, x1]; c01 = s[y1][:, x0]; c11 = s[y1][:, x1]
    top = c00 + (c10 - c00) * fx; bot = c01 + (c11 - c01) * fx
    return np.clip(top + (bot - top) * fy, 0, 255).astype(np.uint8)


def pack_scene(meshes: List[Mesh], lights: List[Light], env: Optional[np.ndarray] = None,
               env_intensity: float = 1.0, name: str = "scene") -> PackedScene:
    """Pack meshes (in actor iteration order) + lights exactly like the reference's upload path."""
    pos_l, nrm_l, uv_l, idx_l, mat_l, inst = [], [], [], [], [], []
    mat_colors: List[List[float]] = []
    tex_sources: List[np.ndarray] = []
    voff = 0; toff = 0
    nobits = np.array([NO_TEX], "<u4").view("<f4")[0]
    for mi, m in enumerate(meshes):
        nv = m.positions.shape[0]; nt = m.indices.shape[0] // 3
        pos_l.append(m.positions.astype(f32)); nrm_l.append(m.normals.astype(f32)); uv_l.append(m.uvs.astype(f32))
        idx_l.append(m.indices.astype(np.uint32) + np.uint32(voff))
        mat_l.append(np.full(nt, mi, np.uint32))
        rec = np.zeros(1, INSTANCE_DTYPE)
        rec["first_tri"] = toff; rec["tri_count"] = nt; rec["xform"] = m.xform; rec["mask"] = 0xFF
        inst.append(rec)
        # texture layers in the reference's order: diffuse (or 1x1 solid), normal, roughMetal, emissive
        def layer(t):
            if t is None: return -1
            tex_sources.append(np.ascontiguousarray(t, np.uint8)); return len(tex_sources) - 1
        if m.albedo_tex is not None:
            d_idx = layer(m.albedo_tex)
        else:
            solid = np.array([[[ _linear_to_srgb8(m.base_color[0]), _linear_to_srgb8(m.base_color[1]),
                                 _linear_to_srgb8(m.base_color[2]), 255]]], np.uint8)
            d_idx = layer(solid)
        n_idx = layer(m.normal_tex); rm_idx = layer(m.rough_metal_tex); e_idx = layer(m.emissive_tex)
        def bits(i): return np.array([NO_TEX if i < 0 else i], "<u4").view("<f4")[0]
        mat_colors.append([m.base_color[0], m.base_color[1], m.base_color[2], bits(d_idx)])
        mat_colors.append([m.roughness, m.metallic, bits(n_idx), bits(e_idx)])
        mat_colors.append([bits(rm_idx), 0.0, 0.0, 0.0])
        voff += nv; toff += nt
    nverts = voff
    positions = np.zeros((nverts, VERTEX_STRIDE // 4), f32)
    positions[:, 0:3] = np.concatenate(pos_l)
    nrm3 = np.concatenate(nrm_l); positions[:, 6:9] = nrm3  # Vertex.normal at offset 24
    normals = np.zeros((nverts, 4), f32); normals[:, :3] = nrm3
    uvs = np.concatenate(uv_l).astype(f32)
    # texture array: every layer resized to min(maxSrc, 2048)  (rt_build.cpp:533-546)
    tw = min(max(t.shape[1] for t in tex_sources), 2048); th = min(max(t.shape[0] for t in tex_sources), 2048)
    textures = np.stack([resize_rgba8_bilinear(t, tw, th) for t in tex_sources])
    # light SSBO: header memset 0xFF, then count, (env idx), env intensity  (light_upload.cpp:270-283)
    hdr = bytearray(b"\xff" * 16)
    hdr[0:4] = struct.pack("<I", len(lights))
    hdr[8:12] = struct.pack("<f", env_intensity)
    if env is not None and len(lights) > 0:   # quirk Q13: env is only wired when >= 1 light exists
        hdr[4:8] = struct.pack("<I", len(tex_sources))   # bindless index appended after the layers
    ssbo = bytes(hdr) + b"".join(l.pack() for l in lights)
    mc = np.array(mat_colors, dtype=f32)
    # keep the exact bit patterns of the packed texture indices (NaN payloads survive np.array on f32 inputs)
    return PackedScene(positions=positions, indices=np.concatenate(idx_l).astype(np.uint32), normals=normals, uvs=uvs,
                       mat_ids=np.concatenate(mat_l), instances=np.concatenate(inst), mat_colors=mc,
                       textures=textures, light_ssbo=np.frombuffer(ssbo, np.uint8).copy(),
                       env=None if env is None else np.ascontiguousarray(env, f32), name=name)


# ------------------------------------------------------------------------------------------
# Primitive meshes
# ------------------------------------------------------------------------------------------
def quad_mesh(a, b, c, d, normal) -> Mesh:
    """addQuad (examples/cornell_box.cpp:32-53): 4 verts, tris (0,1,2),(0,2,3), uv = 0."""
    p = np.array([a, b, c, d], f32)
    return Mesh(positions=p, normals=np.tile(np.asarray(normal, f32), (4, 1)), uvs=np.zeros((4, 2), f32),
                indices=np.array([0, 1, 2, 0, 2, 3], np.uint32))


def uv_sphere_mesh(sectors: int = 32, stacks: int = 16, radius: float = 0.5) -> Mesh:
    """ComponentFactory::generateSphereMesh (scene/component/component_factory.cpp:347-399)."""
    pi = f32(np.pi)
    i = np.arange(stacks + 1, dtype=f32); j = np.arange(sectors + 1, dtype=f32)
    phi = (pi * i / f32(stacks)).astype(f32); theta = (f32(2.0) * pi * j / f32(sectors)).astype(f32)
    sp, cp = np.sin(phi, dtype=f32), np.cos(phi, dtype=f32)
    st, ct = np.sin(theta, dtype=f32), np.cos(theta, dtype=f32)
    x = (ct[None, :] * sp[:, None]).astype(f32); y = np.broadcast_to(cp[:, None], x.shape).astype(f32)
    z = (st[None, :] * sp[:, None]).astype(f32)
    n = np.stack([x, y, z], -1).reshape(-1, 3)
    uv = np.stack([np.broadcast_to((j / f32(sectors))[None, :], x.shape), np.broadcast_to((i / f32(stacks))[:, None], x.shape)], -1).reshape(-1, 2)
    idx = []
    for a in range(stacks):
        for b in range(sectors):
            first = a * (sectors + 1) + b; second = first + sectors + 1
            idx += [first, second, first + 1, second, second + 1, first + 1]
    return Mesh(positions=(n * f32(radius)).astype(f32), normals=n.astype(f32), uvs=uv.astype(f32), indices=np.array(idx, np.uint32))


# Iteration order of libstdc++'s std::unordered_map<uint64_t, Actor::Ptr> after inserting keys
# 1..20 in order (Scene::actors, scene/scene.hpp:46,151).  BLAS / matID / texture-layer / light
# order all follow it (quirk Q9).  tests/test_scenes.py re-derives this list with a real
# std::unordered_map compiled on the spot, so a libstdc++ change cannot silently break it.
LIBSTDCXX_ORDER_1_TO_20 = [20, 19, 18, 17, 16, 15, 14, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]


# ------------------------------------------------------------------------------------------
# Config 1: Cornell box exactly as examples/cornell_box.cpp:85-151
# ------------------------------------------------------------------------------------------
CORNELL_LIGHT_COLORS = [(1.0, 0.3, 0.2), (0.2, 1.0, 0.3), (0.3, 0.4, 1.0), (1.0, 0.9, 0.3), (1.0, 0.5, 0.0), (0.8, 0.2, 1.0),
                        (0.0, 1.0, 1.0), (1.0, 0.0, 0.5), (1.0, 1.0, 1.0), (0.5, 1.0, 0.5), (1.0, 0.7, 0.5), (0.5, 0.7, 1.0)]
CORNELL_LIGHT_POS = [(-3, 4, -3), (0, 4, -3), (3, 4, -3), (-3, 4, 0), (0, 4, 0), (3, 4, 0), (-3, 4, 3), (0, 4, 3), (3, 4, 3),
                     (-4, 0, 0), (4, 0, 0), (0, 0, -4)]


def cornell_box(actor_order: str = "libstdc++") -> PackedScene:
    """2058 triangles, 7 instances, 12 sphere lights, no env map.

    Actor ids: World=1 (Scene ctor, scene.cpp:27-29), walls 2..6, MetalSphere 7, GlassSphere 8,
    Light0..11 = 9..20.  ``actor_order`` selects how Scene::actors is iterated.
    """
    S = 5.0
    white, red, green = (0.73, 0.73, 0.73), (0.65, 0.05, 0.05), (0.12, 0.45, 0.15)
    LBB, RBB, LTB, RTB = (-S, -S, -S), (S, -S, -S), (-S, S, -S), (S, S, -S)
    LBF, RBF, LTF, RTF = (-S, -S, S), (S, -S, S), (-S, S, S), (S, S, S)
    actors = {}
    def wall(aid, name, a, b, c, d, n, col):
        m = quad_mesh(a, b, c, d, n); m.base_color = col; m.roughness = 0.95; m.metallic = 0.0; m.name = name
        actors[aid] = m
    wall(2, "Back", LBB, RBB, RTB, LTB, (0, 0, 1), white)
    wall(3, "Left", LBB, LTB, LTF, LBF, (1, 0, 0), red)
    wall(4, "Right", RBB, RBF, RTF, RTB, (-1, 0, 0), green)
    wall(5, "Floor", LBB, LBF, RBF, RBB, (0, 1, 0), white)
    wall(6, "Ceiling", LTB, RTB, RTF, LTF, (0, -1, 0), white)
    ms = uv_sphere_mesh(); ms.xform = trs((-2.0, -S + 2.0, 0.0), (2.0, 2.0, 2.0))
    ms.base_color = (0.95, 0.93, 0.88); ms.roughness = 0.05; ms.metallic = 1.0; ms.name = "MetalSphere"
    gs = uv_sphere_mesh(); gs.xform = trs((2.5, -S + 1.8, 1.5), (1.8, 1.8, 1.8))
    gs.base_color = (0.9, 0.95, 1.0); gs.roughness = 0.02; gs.metallic = 0.0; gs.name = "GlassSphere"
    actors[7] = ms; actors[8] = gs
    lights = {9 + i: Light(position=tuple(float(v) for v in CORNELL_LIGHT_POS[i]), type=0, color=CORNELL_LIGHT_COLORS[i],
                           intensity=5.0, radius=0.3) for i in range(12)}
    ids = list(range(1, 21))
    if actor_order == "libstdc++": order = LIBSTDCXX_ORDER_1_TO_20
    elif actor_order == "insertion": order = ids
    elif actor_order == "reverse": order = ids[::-1]
    else: raise ValueError(actor_order)
    meshes = [actors[i] for i in order if i in actors]
    ls = [lights[i] for i in order if i in lights]
    return pack_scene(meshes, ls, env=None, name="cornell_box")


def cornell_camera() -> Camera:
    return Camera(position=(0.0, 0.0, 13.0), yaw=-90.0, pitch=0.0, fov=38.0)

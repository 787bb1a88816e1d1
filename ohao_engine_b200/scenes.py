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
class SubMaterial:
    """One material of a multi-material model (a glTF material): Model::materialColors / materialMetallic and the
    per-material texture tables of model_gltf.cpp:373-541."""
    base_color: tuple = (0.8, 0.8, 0.8)
    roughness: float = 0.5
    metallic: float = 0.0
    albedo_tex: Optional[np.ndarray] = None
    normal_tex: Optional[np.ndarray] = None
    rough_metal_tex: Optional[np.ndarray] = None
    emissive_tex: Optional[np.ndarray] = None
    name: str = "material"


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
    # multi-material models (glTF): one SubMaterial per material and one index per triangle (Model::materialPerTriangle);
    # when absent the mesh is one material described by the fields above (the MaterialComponent path)
    materials: Optional[List[SubMaterial]] = None
    material_per_triangle: Optional[np.ndarray] = None

    def sub_materials(self) -> List[SubMaterial]:
        if self.materials:
            return self.materials
        return [SubMaterial(self.base_color, self.roughness, self.metallic, self.albedo_tex, self.normal_tex, self.rough_metal_tex, self.emissive_tex, self.name)]


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
        subs = m.sub_materials()
        color_offset = len(mat_colors) // 3                 # rt_build.cpp:208: material ids are offset by the materials packed so far
        if m.material_per_triangle is not None:
            mat_l.append(np.asarray(m.material_per_triangle, np.uint32) + np.uint32(color_offset))
        else:
            mat_l.append(np.full(nt, color_offset, np.uint32))
        rec = np.zeros(1, INSTANCE_DTYPE)
        rec["first_tri"] = toff; rec["tri_count"] = nt; rec["xform"] = m.xform; rec["mask"] = 0xFF
        inst.append(rec)
        # texture layers in the reference's order: diffuse (or 1x1 solid), normal, roughMetal, emissive — per material
        def layer(t):
            if t is None: return -1
            tex_sources.append(np.ascontiguousarray(t, np.uint8)); return len(tex_sources) - 1
        def bits(i): return np.array([NO_TEX if i < 0 else i], "<u4").view("<f4")[0]
        for sm in subs:
            if sm.albedo_tex is not None:
                d_idx = layer(sm.albedo_tex)
            else:
                solid = np.array([[[ _linear_to_srgb8(sm.base_color[0]), _linear_to_srgb8(sm.base_color[1]),
                                     _linear_to_srgb8(sm.base_color[2]), 255]]], np.uint8)
                d_idx = layer(solid)
            n_idx = layer(sm.normal_tex); rm_idx = layer(sm.rough_metal_tex); e_idx = layer(sm.emissive_tex)
            mat_colors.append([sm.base_color[0], sm.base_color[1], sm.base_color[2], bits(d_idx)])
            mat_colors.append([sm.roughness, sm.metallic, bits(n_idx), bits(e_idx)])
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


def emissive_mesh_lights(meshes: List[Mesh]) -> List[Light]:
    """One auto-generated sphere light per mesh whose first emissive-textured material is bright enough
    (light_upload.cpp:183-247): centre = world position of the bounding-box centre, radius = 0.3 x |bbox diagonal| in
    object space, colour = mean of the texels with luminance > 0.05, intensity = min(0.1 x their summed luminance, 20)."""
    out: List[Light] = []
    for m in meshes:
        for sm in m.sub_materials():
            if sm.emissive_tex is None:
                continue
            px = np.asarray(sm.emissive_tex, np.uint8)[..., :3].reshape(-1, 3).astype(f32) / f32(255.0)
            lum = px[:, 0] * f32(0.2126) + px[:, 1] * f32(0.7152) + px[:, 2] * f32(0.0722)
            mask = lum > f32(0.05)
            power = float(np.cumsum(lum[mask], dtype=f32)[-1]) if mask.any() else 0.0        # sequential fp32 sum like the reference
            if power > 0.1:
                col = px[mask].astype(np.float64).sum(0) / int(mask.sum())
                bmin, bmax = m.positions.min(0).astype(f32), m.positions.max(0).astype(f32)
                mid = (bmin + bmax) * f32(0.5)
                x = np.asarray(m.xform, f32).reshape(3, 4)
                center = x[:, :3] @ mid + x[:, 3]
                out.append(Light(position=tuple(float(v) for v in center), color=tuple(float(f32(v)) for v in col), intensity=min(power * 0.1, 20.0),
                                 radius=float(np.linalg.norm(bmax - bmin) * f32(0.3)), direction=(0.0, -1.0, 0.0), outer_cone=0.0))   # gl.extra = vec4(0)
            break                                        # one light per actor
    return out


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


# ------------------------------------------------------------------------------------------
# Synthetic stand-ins for the assets that cannot travel to the GPU box (configs 2-4).
# /root/reference (DamagedHelmet.glb, env_outdoor.hdr) does not exist there, so the "Helmet-class"
# and "2 M-triangle" workloads are seeded procedural scenes of the named triangle counts, texture
# sets and resolutions.
# ------------------------------------------------------------------------------------------
def procedural_env(width: int = 1024, height: int = 512, sun_dir=(0.35, 0.75, 0.55), sun_radiance: float = 4000.0,
                   seed: int = 7) -> np.ndarray:
    """Outdoor-HDRI stand-in: sky gradient + small very bright sun + dark ground, RGBA32F [h][w][4].

    Row 0 is theta = 0 (+Y) for the CDF/importance sampling convention (env_sampling.glsl:40-49);
    the miss shader looks it up with v flipped (quirk Q4) — both use this one image."""
    rng = np.random.default_rng(seed)
    v = (np.arange(height, dtype=np.float64) + 0.5) / height
    u = (np.arange(width, dtype=np.float64) + 0.5) / width
    theta = v[:, None] * np.pi; phi = (u[None, :] - 0.5) * 2 * np.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta) * np.ones_like(phi), np.sin(theta) * np.sin(phi)], -1)
    up = np.clip(d[..., 1], -1, 1)
    sky = np.where(up[..., None] > 0, (1 - up[..., None]) * np.array([0.9, 0.95, 1.0]) + up[..., None] * np.array([0.25, 0.45, 0.95]),
                   np.array([0.18, 0.16, 0.14]) * (1 + 0.5 * up[..., None]))
    clouds = 0.15 * (np.sin(7 * phi + 3 * theta) * np.sin(5 * theta + 1.3) + 1) * (up > 0)
    sky = sky * (1.0 + clouds[..., None]) * 1.2
    sd = np.asarray(sun_dir, np.float64); sd /= np.linalg.norm(sd)
    cosang = d @ sd
    sun = np.exp(np.minimum((cosang - 1.0) / 2.0e-4, 0.0)) * sun_radiance
    img = sky + sun[..., None] * np.array([1.0, 0.93, 0.82])
    img += 0.01 * rng.random(img.shape)
    out = np.ones((height, width, 4), np.float32); out[..., :3] = img.astype(np.float32)
    return out


def _displaced_sphere(sectors: int, stacks: int, radius: float, amp: float, seed: int):
    rng = np.random.default_rng(seed)
    i = np.arange(stacks + 1, dtype=np.float64); j = np.arange(sectors + 1, dtype=np.float64)
    phi = np.pi * i / stacks; theta = 2 * np.pi * j / sectors
    sp, cp = np.sin(phi)[:, None], np.cos(phi)[:, None]; st, ct = np.sin(theta)[None, :], np.cos(theta)[None, :]
    n = np.stack([ct * sp, cp * np.ones_like(ct), st * sp], -1)
    k = rng.integers(2, 9, size=(6, 3)); ph = rng.random(6) * 6.28
    disp = sum(np.sin(n @ k[q].astype(np.float64) + ph[q]) for q in range(6)) / 6.0
    pos = n * (radius * (1.0 + amp * disp))[..., None]
    pos[:, -1] = pos[:, 0]   # close the seam
    P = pos.reshape(-1, 3)
    idx = np.empty((stacks, sectors, 6), np.int64)
    a = (np.arange(stacks)[:, None] * (sectors + 1) + np.arange(sectors)[None, :]); b = a + sectors + 1
    idx[..., 0] = a; idx[..., 1] = b; idx[..., 2] = a + 1; idx[..., 3] = b; idx[..., 4] = b + 1; idx[..., 5] = a + 1
    I = idx.reshape(-1, 3)
    fn = np.cross(P[I[:, 1]] - P[I[:, 0]], P[I[:, 2]] - P[I[:, 0]])
    vn = np.zeros_like(P)
    for c in range(3): np.add.at(vn, I[:, c], fn)
    # orient outward and normalise (degenerate pole fans fall back to the radial direction)
    radial = P / np.maximum(np.linalg.norm(P, axis=1, keepdims=True), 1e-12)
    flip = np.sum(vn * radial, 1) < 0; vn[flip] = -vn[flip]
    ln = np.linalg.norm(vn, axis=1, keepdims=True)
    vn = np.where(ln > 1e-12, vn / np.maximum(ln, 1e-12), radial)
    uv = np.stack([np.broadcast_to(j[None, :] / sectors, (stacks + 1, sectors + 1)), np.broadcast_to(i[:, None] / stacks, (stacks + 1, sectors + 1))], -1).reshape(-1, 2)
    return P.astype(f32), vn.astype(f32), uv.astype(f32), I.reshape(-1).astype(np.uint32)


def _procedural_textures(size: int, seed: int):
    """albedo (sRGB-ish), tangent-space normal, ORM (AO, rough, metal), emissive — RGBA8 [size][size][4]."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size].astype(np.float64) / size
    def noise(fx, fy, p): return 0.5 + 0.5 * np.sin(2 * np.pi * (fx * x + fy * y) + p)
    n1 = noise(3, 5, 0.3) * noise(7, 2, 1.1); n2 = noise(13, 11, 2.0); n3 = noise(29, 31, 0.7)
    alb = np.stack([0.55 + 0.4 * n1, 0.45 + 0.35 * n2, 0.35 + 0.3 * n3, np.ones_like(x)], -1)
    gx = 0.25 * np.cos(2 * np.pi * (13 * x + 11 * y) + 2.0); gy = 0.25 * np.cos(2 * np.pi * (29 * x + 31 * y) + 0.7)
    nz = np.sqrt(np.maximum(1 - gx * gx - gy * gy, 0.0))
    nrm = np.stack([0.5 + 0.5 * gx, 0.5 + 0.5 * gy, 0.5 + 0.5 * nz, np.ones_like(x)], -1)
    panels = ((np.floor(x * 8) + np.floor(y * 8)) % 2)
    orm = np.stack([0.8 + 0.2 * n1, 0.25 + 0.6 * n2, panels * 0.9 + 0.05, np.ones_like(x)], -1)
    em = np.zeros((size, size, 4)); em[..., 3] = 1
    spots = (noise(5, 4, 0.0) > 0.97) & (noise(4, 6, 1.0) > 0.9)
    em[spots, 0] = 1.0; em[spots, 1] = 0.6; em[spots, 2] = 0.2
    q = lambda a: np.clip(a * 255.0 + 0.5, 0, 255).astype(np.uint8)
    _ = rng
    return q(alb), q(nrm), q(orm), q(em)


def helmet_class(ntris: int = 50000, tex_size: int = 2048, env_size=(1024, 512), seed: int = 1) -> PackedScene:
    """Config 2 stand-in: one ~`ntris` displaced-sphere hero with the helmet's 5-texture PBR material
    (albedo, normal, ORM, emissive -> also spawns the reference's auto emissive sphere light,
    light_upload.cpp:183-247), a 2-triangle ground quad, 3 studio sphere lights, outdoor env."""
    stacks = max(4, int(round(math.sqrt(ntris / 5.0)))); sectors = max(8, ntris // (2 * stacks))
    P, N, UV, I = _displaced_sphere(sectors, stacks, 1.0, 0.12, seed)
    alb, nrm, orm, em = _procedural_textures(tex_size, seed)
    hero = Mesh(positions=P, normals=N, uvs=UV, indices=I, xform=trs((0.0, 1.15, 0.0), (1.0, 1.0, 1.0)), base_color=(1.0, 1.0, 1.0),
                roughness=1.0, metallic=1.0, albedo_tex=alb, normal_tex=nrm, rough_metal_tex=orm, emissive_tex=em, name="Hero")
    g = 12.0
    ground = quad_mesh((-g, 0, -g), (-g, 0, g), (g, 0, g), (g, 0, -g), (0, 1, 0)); ground.base_color = (0.5, 0.5, 0.5); ground.roughness = 0.8; ground.name = "Ground"
    lights = [Light(position=(4.0, 5.0, 3.0), color=(1.0, 0.95, 0.9), intensity=60.0, radius=0.5),
              Light(position=(-5.0, 3.0, 2.0), color=(0.6, 0.7, 1.0), intensity=25.0, radius=0.8),
              Light(position=(0.0, 4.0, -5.0), color=(1.0, 1.0, 1.0), intensity=30.0, radius=0.4)]
    # auto-generated emissive mesh light (light_upload.cpp:183-247)
    bright = (em[..., :3].astype(np.float32) / 255.0)
    lum = bright @ np.array([0.2126, 0.7152, 0.0722], np.float32)
    mask = lum > 0.05
    if mask.any() and float(lum[mask].sum()) > 0.1:
        col = bright[mask].mean(0); inten = min(float(lum[mask].sum()) * 0.1, 20.0)
        bmin, bmax = P.min(0), P.max(0)
        center = (bmin + bmax) * 0.5 + np.array([0.0, 1.15, 0.0], np.float32)
        lights.append(Light(position=tuple(float(v) for v in center), color=tuple(float(v) for v in col), intensity=inten,
                            radius=float(np.linalg.norm(bmax - bmin) * 0.3), outer_cone=0.0))
    env = procedural_env(env_size[0], env_size[1])
    return pack_scene([hero, ground], lights, env=env, name=f"helmet_class_{I.size // 3 + 2}")


def helmet_camera() -> Camera:
    return Camera(position=(0.0, 1.4, 4.2), yaw=-90.0, pitch=-4.0, fov=40.0)


def synthetic_2m(nblobs: int = 1000, tris_per_blob: int = 2000, seed: int = 12345, extent: float = 100.0, env_size=(1024, 512),
                 nmaterials: int = 16) -> PackedScene:
    """Config 3/4 stand-in: `nblobs` displaced spheres of `tris_per_blob` triangles scattered in an
    `extent`-metre cube over a 2-triangle ground, 16 materials (rough U[0.05,1], metal in {0,1} p=0.3,
    albedo U[0.1,0.9]^3), 8 sphere lights, outdoor env.  Flattened to ONE actor with per-triangle material ids."""
    rng = np.random.default_rng(seed)
    stacks = max(4, int(round(math.sqrt(tris_per_blob / 3.2)))); sectors = max(8, tris_per_blob // (2 * stacks))
    P0, N0, UV0, I0 = _displaced_sphere(sectors, stacks, 1.0, 0.15, seed)
    nv, nt = P0.shape[0], I0.size // 3
    centers = np.stack([rng.uniform(-extent / 2, extent / 2, nblobs), rng.uniform(1.5, extent * 0.35, nblobs), rng.uniform(-extent / 2, extent / 2, nblobs)], -1)
    radii = rng.uniform(0.8, 3.0, nblobs)
    rot = rng.uniform(0, 2 * np.pi, nblobs)
    c, s = np.cos(rot), np.sin(rot)
    Px = P0[None, :, 0] * c[:, None] + P0[None, :, 2] * s[:, None]; Pz = -P0[None, :, 0] * s[:, None] + P0[None, :, 2] * c[:, None]
    Nx = N0[None, :, 0] * c[:, None] + N0[None, :, 2] * s[:, None]; Nz = -N0[None, :, 0] * s[:, None] + N0[None, :, 2] * c[:, None]
    P = np.stack([Px, np.broadcast_to(P0[None, :, 1], Px.shape), Pz], -1) * radii[:, None, None] + centers[:, None, :]
    N = np.stack([Nx, np.broadcast_to(N0[None, :, 1], Nx.shape), Nz], -1)
    I = (I0[None, :].astype(np.int64) + (np.arange(nblobs, dtype=np.int64) * nv)[:, None]).reshape(-1)
    UV = np.broadcast_to(UV0[None], (nblobs, nv, 2)).reshape(-1, 2)
    blob_mat = rng.integers(0, nmaterials, nblobs)
    g = extent * 0.75
    GP = np.array([(-g, 0, -g), (-g, 0, g), (g, 0, g), (g, 0, -g)], np.float64); GN = np.tile(np.array([0, 1, 0.0]), (4, 1))
    base = nblobs * nv
    positions = np.concatenate([P.reshape(-1, 3), GP]).astype(f32); normals3 = np.concatenate([N.reshape(-1, 3), GN]).astype(f32)
    uvs = np.concatenate([UV, np.zeros((4, 2))]).astype(f32)
    indices = np.concatenate([I, np.array([0, 1, 2, 0, 2, 3], np.int64) + base]).astype(np.uint32)
    mat_ids = np.concatenate([np.repeat(blob_mat, nt), np.array([nmaterials, nmaterials])]).astype(np.uint32)
    nverts = positions.shape[0]
    pos_rec = np.zeros((nverts, VERTEX_STRIDE // 4), f32); pos_rec[:, 0:3] = positions; pos_rec[:, 6:9] = normals3
    nrm4 = np.zeros((nverts, 4), f32); nrm4[:, :3] = normals3
    # materials: 16 random + ground; every material owns a 1x1 solid diffuse layer (quirk Q1)
    albedo = rng.uniform(0.1, 0.9, (nmaterials, 3)); rough = rng.uniform(0.05, 1.0, nmaterials); metal = (rng.random(nmaterials) < 0.3).astype(np.float64)
    albedo = np.concatenate([albedo, [[0.45, 0.45, 0.42]]]); rough = np.append(rough, 0.9); metal = np.append(metal, 0.0)
    nobits = np.array([NO_TEX], "<u4").view("<f4")[0]
    mc = np.zeros(((nmaterials + 1) * 3, 4), f32); tex = np.zeros((nmaterials + 1, 1, 1, 4), np.uint8)
    for m in range(nmaterials + 1):
        mc[m * 3 + 0] = [albedo[m, 0], albedo[m, 1], albedo[m, 2], np.array([m], "<u4").view("<f4")[0]]
        mc[m * 3 + 1] = [rough[m], metal[m], nobits, nobits]
        mc[m * 3 + 2] = [nobits, 0, 0, 0]
        tex[m, 0, 0] = [_linear_to_srgb8(albedo[m, 0]), _linear_to_srgb8(albedo[m, 1]), _linear_to_srgb8(albedo[m, 2]), 255]
    inst = np.zeros(1, INSTANCE_DTYPE); inst["first_tri"] = 0; inst["tri_count"] = mat_ids.shape[0]; inst["xform"] = trs(); inst["mask"] = 0xFF
    lights = [Light(position=(float(rng.uniform(-extent / 3, extent / 3)), float(rng.uniform(extent * 0.2, extent * 0.45)), float(rng.uniform(-extent / 3, extent / 3))),
                    color=tuple(float(v) for v in rng.uniform(0.5, 1.0, 3)), intensity=float(rng.uniform(2000, 6000)), radius=float(rng.uniform(1.0, 3.0))) for _ in range(8)]
    hdr = bytearray(b"\xff" * 16); hdr[0:4] = struct.pack("<I", len(lights)); hdr[4:8] = struct.pack("<I", nmaterials + 1); hdr[8:12] = struct.pack("<f", 1.0)
    ssbo = bytes(hdr) + b"".join(l.pack() for l in lights)
    env = procedural_env(env_size[0], env_size[1])
    return PackedScene(positions=pos_rec, indices=indices, normals=nrm4, uvs=uvs, mat_ids=mat_ids, instances=inst, mat_colors=mc, textures=tex,
                       light_ssbo=np.frombuffer(ssbo, np.uint8).copy(), env=env, name=f"synthetic_{mat_ids.shape[0]}")


def synthetic_camera(extent: float = 100.0, angle_deg: float = 0.0) -> Camera:
    """Slow orbit camera of config 3 (0.5 deg / frame)."""
    r = extent * 0.62; a = math.radians(angle_deg)
    pos = (r * math.sin(a), extent * 0.22, r * math.cos(a))
    yaw = math.degrees(math.atan2(-pos[2], -pos[0]))
    return Camera(position=pos, yaw=yaw, pitch=-12.0, fov=50.0)

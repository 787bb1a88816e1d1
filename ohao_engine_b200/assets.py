"""Scene ingestion for the path-tracing hot path: binary glTF (.glb) models, Radiance
RGBE (.hdr) environment maps and the reference's texture-array rules, producing the §3.2 arrays `scenes.pack_scene`
uploads through the C ABI.

What is mirrored, file by file:
  * Model::loadFromGLTF (ohao/scene/asset/model_gltf.cpp:14-571, via tinygltf 2.9.3): every primitive of every MESH is
    appended in file order — the node hierarchy and its transforms are NOT applied (the loader walks `meshes`, not
    `scenes`/`nodes`); POSITION/NORMAL/TEXCOORD_0 as float accessors with their buffer-view stride, indices u8/u16/u32,
    one material id per triangle; per material baseColorFactor / roughnessFactor / metallicFactor, the base colour
    REPLACED by the average of the albedo texture when there is one (:405-428), albedo / normal / emissive images as
    RGBA8, the metallic-roughness image repacked to (R, G = roughness, B = metallic, 255) (:472-497).
  * stbi_loadf(path, 4) of light_upload.cpp:310 for .hdr: RGBE -> float with f = ldexp(1, e - 136), alpha = 1, rows
    top to bottom as stored, no flip (quirk Q4 lives on that).
  * the texture-array builder of rt_build.cpp:391-823 is `scenes.pack_scene` (layer order diffuse-or-solid, normal,
    rough-metal, emissive per material; every layer resized to min(max source, 2048) with resizeRGBA8Bilinear).
Image decoding (PNG / JPEG payloads embedded in the .glb) is Pillow's; tinygltf uses stb_image there — the two
decoders agree on PNG bit-for-bit and on baseline JPEG to within the IDCT's last bit.
"""
from __future__ import annotations

import io
import json
import struct
from typing import List, Optional

import numpy as np

from . import scenes

f32 = np.float32


# ---------------------------------------------------------------------------------------------------------------------
# Radiance .hdr (RGBE)
# ---------------------------------------------------------------------------------------------------------------------
def load_hdr(path: str) -> np.ndarray:
    """(H, W, 4) float32, alpha = 1 — what stbi_loadf(path, &w, &h, &c, 4) returns for a Radiance file."""
    data = open(path, "rb").read()
    pos = 0

    def line():
        nonlocal pos
        e = data.index(b"\n", pos); s = data[pos:e]; pos = e + 1
        return s.decode("latin-1").rstrip("\r")

    magic = line()
    if magic not in ("#?RADIANCE", "#?RGBE"):
        raise ValueError(f"{path}: not a Radiance HDR file")
    fmt_ok = False
    while True:
        l = line()
        if l == "":
            break
        if l == "FORMAT=32-bit_rle_rgbe":
            fmt_ok = True
    if not fmt_ok:
        raise ValueError(f"{path}: unsupported HDR format")
    tok = line().split()
    if len(tok) != 4 or tok[0] != "-Y" or tok[2] != "+X":
        raise ValueError(f"{path}: unsupported HDR orientation")
    H, W = int(tok[1]), int(tok[3])
    rgbe = np.zeros((H, W, 4), np.uint8)
    buf = np.frombuffer(data, np.uint8)
    if W < 8 or W >= 32768:
        rgbe[:] = buf[pos:pos + H * W * 4].reshape(H, W, 4)
    else:
        for y in range(H):
            if not (buf[pos] == 2 and buf[pos + 1] == 2 and not (buf[pos + 2] & 0x80)):
                # not run-length encoded: the rest of the file is flat RGBE (stbi__hdr_load's fallback)
                rest = (H - y) * W * 4
                rgbe[y:] = buf[pos:pos + rest].reshape(H - y, W, 4); pos += rest
                break
            if (int(buf[pos + 2]) << 8 | int(buf[pos + 3])) != W:
                raise ValueError(f"{path}: corrupt scanline width")
            pos += 4
            for c in range(4):
                x = 0
                row = rgbe[y, :, c]
                while x < W:
                    n = int(buf[pos]); pos += 1
                    if n > 128:
                        n -= 128
                        row[x:x + n] = buf[pos]; pos += 1
                    else:
                        row[x:x + n] = buf[pos:pos + n]; pos += n
                    x += n
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e != 0, np.ldexp(f32(1.0), e - 136), f32(0.0)).astype(f32)
    out = np.empty((H, W, 4), f32)
    out[..., :3] = rgbe[..., :3].astype(f32) * scale[..., None]
    out[..., 3] = 1.0
    return out


# ---------------------------------------------------------------------------------------------------------------------
# binary glTF
# ---------------------------------------------------------------------------------------------------------------------
_COMP = {5120: ("i1", 1), 5121: ("u1", 1), 5122: ("<i2", 2), 5123: ("<u2", 2), 5125: ("<u4", 4), 5126: ("<f4", 4)}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


class Glb:
    def __init__(self, path: str):
        raw = open(path, "rb").read()
        magic, version, total = struct.unpack_from("<4sII", raw, 0)
        if magic != b"glTF" or version != 2:
            raise ValueError(f"{path}: not a glTF 2.0 binary file")
        off = 12; self.json = None; self.bin = b""
        while off < total:
            n, kind = struct.unpack_from("<II", raw, off); off += 8
            chunk = raw[off:off + n]; off += n
            if kind == 0x4E4F534A: self.json = json.loads(chunk.decode("utf-8"))
            elif kind == 0x004E4942: self.bin = chunk
        if self.json is None:
            raise ValueError(f"{path}: no JSON chunk")

    def accessor(self, idx: int) -> np.ndarray:
        """(count, ncomp) array honouring bufferView.byteStride (getBufferData / getAccessorStride, model_gltf.cpp:136-160)."""
        a = self.json["accessors"][idx]; bv = self.json["bufferViews"][a["bufferView"]]
        dt, sz = _COMP[a["componentType"]]; nc = _NCOMP[a["type"]]
        start = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or sz * nc
        cnt = a["count"]
        buf = np.frombuffer(self.bin, np.uint8, count=(cnt - 1) * stride + sz * nc, offset=start) if cnt else np.zeros(0, np.uint8)
        rows = np.lib.stride_tricks.as_strided(buf, shape=(cnt, sz * nc), strides=(stride, 1))
        return np.ascontiguousarray(rows).view(dt).reshape(cnt, nc)

    def image_rgba8(self, image_index: int) -> Optional[np.ndarray]:
        from PIL import Image
        im = self.json["images"][image_index]
        if "bufferView" not in im:
            return None                                       # external uri: not part of a self-contained .glb
        bv = self.json["bufferViews"][im["bufferView"]]
        blob = self.bin[bv.get("byteOffset", 0):bv.get("byteOffset", 0) + bv["byteLength"]]
        return np.asarray(Image.open(io.BytesIO(blob)).convert("RGBA"), np.uint8)

    def texture_image(self, tex: Optional[dict]) -> Optional[np.ndarray]:
        if not tex or tex.get("index", -1) < 0 or tex["index"] >= len(self.json.get("textures", [])):
            return None
        src = self.json["textures"][tex["index"]].get("source", -1)
        if src < 0 or src >= len(self.json.get("images", [])):
            return None
        return self.image_rgba8(src)


def load_glb(path: str, name: Optional[str] = None) -> scenes.Mesh:
    """One `scenes.Mesh` (= one actor / BLAS) holding every primitive of the file with per-triangle material ids and
    one `scenes.SubMaterial` per glTF material — Model::loadFromGLTF."""
    g = Glb(path); j = g.json
    pos_l, nrm_l, uv_l, idx_l, mat_l = [], [], [], [], []
    voff = 0
    for mesh in j.get("meshes", []):
        for prim in mesh["primitives"]:
            at = prim["attributes"]
            if "POSITION" not in at:
                continue
            P = g.accessor(at["POSITION"]).astype(f32); nv = len(P)
            N = g.accessor(at["NORMAL"]).astype(f32) if "NORMAL" in at else np.tile(np.array([0, 1, 0], f32), (nv, 1))
            T = g.accessor(at["TEXCOORD_0"]).astype(f32)[:, :2] if "TEXCOORD_0" in at else np.zeros((nv, 2), f32)
            I = g.accessor(prim["indices"]).reshape(-1).astype(np.uint32) if "indices" in prim else np.arange(nv, dtype=np.uint32)
            nt = len(I) // 3
            pos_l.append(P[:, :3]); nrm_l.append(N[:, :3]); uv_l.append(T); idx_l.append(I[:nt * 3] + np.uint32(voff))
            mat_l.append(np.full(nt, max(prim.get("material", 0), 0), np.uint32))
            voff += nv
    if not pos_l:
        raise ValueError(f"{path}: no triangle geometry")
    subs: List[scenes.SubMaterial] = []
    for m in j.get("materials", []):
        pbr = m.get("pbrMetallicRoughness", {})
        bc = pbr.get("baseColorFactor", [1, 1, 1, 1])
        sm = scenes.SubMaterial(base_color=(float(bc[0]), float(bc[1]), float(bc[2])), roughness=float(pbr.get("roughnessFactor", 1.0)),
                                metallic=float(pbr.get("metallicFactor", 1.0)), name=m.get("name", "gltf_material"))
        alb = g.texture_image(pbr.get("baseColorTexture"))
        if alb is not None:
            # the material's base colour becomes the texture's mean (double sums / (count * 255), model_gltf.cpp:405-428)
            s = alb[..., :3].reshape(-1, 3).astype(np.float64).sum(0) / (alb.shape[0] * alb.shape[1] * 255.0)
            sm.base_color = (float(f32(s[0])), float(f32(s[1])), float(f32(s[2]))); sm.albedo_tex = alb
        sm.normal_tex = g.texture_image(m.get("normalTexture"))
        rm = g.texture_image(pbr.get("metallicRoughnessTexture"))
        if rm is not None:
            rm = rm.copy(); rm[..., 3] = 255; sm.rough_metal_tex = rm      # (AO | R, roughness, metallic, 255)
        sm.emissive_tex = g.texture_image(m.get("emissiveTexture"))
        subs.append(sm)
    if not subs:
        subs.append(scenes.SubMaterial(base_color=(0.8, 0.8, 0.8), roughness=0.5, metallic=0.0))
    return scenes.Mesh(positions=np.concatenate(pos_l), normals=np.concatenate(nrm_l), uvs=np.concatenate(uv_l), indices=np.concatenate(idx_l),
                       materials=subs, material_per_triangle=np.minimum(np.concatenate(mat_l), len(subs) - 1).astype(np.uint32),
                       name=name or path.rsplit("/", 1)[-1])


# ---------------------------------------------------------------------------------------------------------------------
# examples/turntable.cpp, env mode: config 2 ("Helmet-class glTF under outdoor HDRI") with the real assets
# ---------------------------------------------------------------------------------------------------------------------
def turntable_model_xform(mesh: scenes.Mesh, mode: str = "env") -> np.ndarray:
    """turntable.cpp:118-150 for a Y-up model: scale to a height of 4 (5.5 in the mirror room), rotate 180 degrees about
    Y, centre in x/z, feet on the floor (y = 0 in env mode, -5 in the rooms).  Row-major 3x4 object->world (T * R * S)."""
    bmin, bmax = mesh.positions.min(0).astype(f32), mesh.positions.max(0).astype(f32)
    height = max(float(bmax[1] - bmin[1]), float(bmax[2] - bmin[2]))
    if float(bmax[1] - bmin[1]) < float(bmax[2] - bmin[2]):
        raise NotImplementedError("Z-up models (turntable.cpp:139-150) are not needed by the in-tree assets")
    s = f32((5.5 if mode == "mirror" else 4.0) / height)
    c = (bmin + bmax) * f32(0.5)
    floor_y = f32(0.0 if mode == "env" else -5.0)
    m = np.zeros((3, 4), f32)
    m[0, 0], m[1, 1], m[2, 2] = -s, s, -s                       # R_y(180) * S
    m[0, 3], m[1, 3], m[2, 3] = -c[0] * s, floor_y - bmin[1] * s, -c[2] * s
    return m.reshape(12)


def turntable_camera(frame: int = 0, total_frames: int = 120, mode: str = "env") -> scenes.Camera:
    """turntable.cpp:199-217: orbit of radius 8 (env) at height 1, pitch 15 degrees, fov 70."""
    import math
    radius = {"env": 8.0, "mirror": 4.5}.get(mode, 4.2); height = 1.0 if mode == "env" else -3.5
    a = frame / total_frames * 2.0 * 3.14159
    cx, cz = radius * math.cos(a), radius * math.sin(a)
    return scenes.Camera(position=(cx, height, cz), yaw=math.degrees(math.atan2(-cz, -cx)), pitch=15.0, fov=70.0)


def turntable_env_scene(glb_path: str, hdr_path: str, env_intensity: float = 1.0) -> scenes.PackedScene:
    """The scene `turntable <model.glb> env` renders: the model, one warm key sphere light (I = 8, r = 1 at (3, 4, 3)),
    the emissive-mesh auto light of light_upload.cpp:183-247, the HDRI with importance sampling + MIS."""
    hero = load_glb(glb_path, name="Model")
    hero.xform = turntable_model_xform(hero, "env")
    lights = [scenes.Light(position=(3.0, 4.0, 3.0), color=(1.0, 0.95, 0.9), intensity=8.0, radius=1.0)]
    lights += scenes.emissive_mesh_lights([hero])
    return scenes.pack_scene([hero], lights, env=load_hdr(hdr_path), env_intensity=env_intensity, name="turntable_env_" + hero.name)


def bake_textures(glb_path: str, out_path: Optional[str] = None) -> str:
    """Decode the first material's embedded images into the side-car the C++ host reads (host/asset_io.hpp loadGLB)."""
    m = load_glb(glb_path); sm = m.materials[0]
    out_path = out_path or glb_path + ".ohbtex"
    layers = [(k, t) for k, t in enumerate((sm.albedo_tex, sm.normal_tex, sm.rough_metal_tex, sm.emissive_tex)) if t is not None]
    with open(out_path, "wb") as f:
        f.write(b"OHBT" + struct.pack("<I", len(layers)))
        for kind, t in layers:
            f.write(struct.pack("<III", kind, t.shape[1], t.shape[0])); f.write(np.ascontiguousarray(t, np.uint8).tobytes())
    return out_path


if __name__ == "__main__":
    import sys
    if len(sys.argv) >= 3 and sys.argv[1] == "bake":
        print(bake_textures(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None))
    else:
        print("usage: python -m ohao_engine_b200.assets bake <model.glb> [out.ohbtex]")

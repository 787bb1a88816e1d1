// host/asset_io.hpp — scene ingestion on the C++20 host side:
//   loadHDR  : Radiance RGBE -> RGBA32F, the stbi_loadf(path, &w, &h, &c, 4) of light_upload.cpp:310 (f = ldexp(1, e - 136),
//              alpha 1, rows as stored — quirk Q4 depends on the missing flip);
//   loadGLB  : binary glTF -> Model + Material the way Model::loadFromGLTF does (ohao/scene/asset/model_gltf.cpp:14-571, via
//              tinygltf): every primitive of every MESH appended in file order, node transforms NOT applied, POSITION / NORMAL /
//              TEXCOORD_0 float accessors with their buffer-view stride, u8/u16/u32 indices, material factors of the first
//              material (the host Actor carries one Material like the reference's MaterialComponent path; multi-material
//              models go through the Python packer, ohao_engine_b200/assets.py).
// Embedded images are JPEG/PNG payloads; the reference decodes them with stb_image inside tinygltf.  This header has no
// image decoder: decoded RGBA8 layers are read from the side-car `<model>.glb.ohbtex` written by
// `python -m ohao_engine_b200.assets bake <model>.glb` (magic "OHBT", u32 count, then {u32 kind, u32 w, u32 h, w*h*4 bytes};
// kind 0 albedo, 1 normal, 2 rough-metal (R, roughness, metallic, 255), 3 emissive).  Without the side-car the model loads
// untextured with the file's factors.
#pragma once
#include "ohao_b200_host.hpp"
#include <cstdlib>
#include <fstream>
#include <map>
#include <variant>

namespace ohao {

inline bool readFile(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end); out.resize(size_t(f.tellg())); f.seekg(0);
    f.read(reinterpret_cast<char*>(out.data()), std::streamsize(out.size()));
    return bool(f);
}

// ---- Radiance .hdr ---------------------------------------------------------------------------------------------------
inline bool loadHDR(const std::string& path, std::vector<float>& rgba, uint32_t& W, uint32_t& H) {
    std::vector<uint8_t> d;
    if (!readFile(path, d)) { std::cerr << "loadHDR: cannot read " << path << "\n"; return false; }
    size_t pos = 0;
    auto line = [&]() { std::string s; while (pos < d.size() && d[pos] != '\n') s.push_back(char(d[pos++])); pos++; if (!s.empty() && s.back() == '\r') s.pop_back(); return s; };
    const std::string magic = line();
    if (magic != "#?RADIANCE" && magic != "#?RGBE") { std::cerr << "loadHDR: not a Radiance file: " << path << "\n"; return false; }
    bool fmt = false;
    for (;;) { std::string l = line(); if (l.empty()) break; if (l == "FORMAT=32-bit_rle_rgbe") fmt = true; if (pos >= d.size()) return false; }
    if (!fmt) { std::cerr << "loadHDR: unsupported format\n"; return false; }
    int h = 0, w = 0;
    if (std::sscanf(line().c_str(), "-Y %d +X %d", &h, &w) != 2 || h <= 0 || w <= 0) { std::cerr << "loadHDR: unsupported orientation\n"; return false; }
    W = uint32_t(w); H = uint32_t(h);
    std::vector<uint8_t> rgbe(size_t(w) * h * 4);
    auto need = [&](size_t n) { return pos + n <= d.size(); };
    bool flat = w < 8 || w >= 32768;
    for (int y = 0; y < h && !flat; y++) {
        if (!need(4)) return false;
        if (!(d[pos] == 2 && d[pos + 1] == 2 && !(d[pos + 2] & 0x80))) {      // stbi__hdr_load's fallback: the rest is flat RGBE
            size_t rest = size_t(h - y) * w * 4; if (!need(rest)) return false;
            std::memcpy(&rgbe[size_t(y) * w * 4], &d[pos], rest); pos += rest; flat = false; y = h; break;
        }
        if (((int(d[pos + 2]) << 8) | int(d[pos + 3])) != w) return false;
        pos += 4;
        for (int c = 0; c < 4; c++) {
            int x = 0;
            while (x < w) {
                if (!need(2)) return false;
                int n = d[pos++];
                if (n > 128) { n -= 128; if (x + n > w) return false; uint8_t v = d[pos++]; for (int k = 0; k < n; k++) rgbe[(size_t(y) * w + x + k) * 4 + c] = v; }
                else { if (x + n > w || !need(size_t(n))) return false; for (int k = 0; k < n; k++) rgbe[(size_t(y) * w + x + k) * 4 + c] = d[pos++]; }
                x += n;
            }
        }
    }
    if (flat) { size_t all = size_t(w) * h * 4; if (!need(all)) return false; std::memcpy(rgbe.data(), &d[pos], all); }
    rgba.resize(size_t(w) * h * 4);
    for (size_t i = 0; i < size_t(w) * h; i++) {
        const uint8_t* p = &rgbe[i * 4];
        const float f = p[3] ? float(std::ldexp(1.0f, int(p[3]) - 136)) : 0.0f;
        rgba[i * 4 + 0] = p[0] * f; rgba[i * 4 + 1] = p[1] * f; rgba[i * 4 + 2] = p[2] * f; rgba[i * 4 + 3] = 1.0f;
    }
    return true;
}

// ---- a JSON value just large enough for glTF ---------------------------------------------------------------------------
struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    double num = 0; bool b = false; std::string str; std::vector<Json> arr; std::map<std::string, Json> obj;
    const Json& operator[](const std::string& k) const { static const Json nul; auto it = obj.find(k); return it == obj.end() ? nul : it->second; }
    const Json& operator[](size_t i) const { static const Json nul; return i < arr.size() ? arr[i] : nul; }
    bool has(const std::string& k) const { return obj.count(k) != 0; }
    double number(double dflt) const { return kind == Num ? num : dflt; }
    int integer(int dflt) const { return kind == Num ? int(num) : dflt; }
};
struct JsonParser {
    const char* p; const char* end; bool ok = true;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; }
    Json value() {
        ws(); Json v;
        if (p >= end) { ok = false; return v; }
        if (*p == '{') {
            v.kind = Json::Obj; p++; ws();
            if (p < end && *p == '}') { p++; return v; }
            while (ok) { ws(); Json k = value(); ws(); if (p >= end || *p != ':' || k.kind != Json::Str) { ok = false; break; } p++; v.obj[k.str] = value(); ws(); if (p < end && *p == ',') { p++; continue; } if (p < end && *p == '}') { p++; break; } ok = false; }
        } else if (*p == '[') {
            v.kind = Json::Arr; p++; ws();
            if (p < end && *p == ']') { p++; return v; }
            while (ok) { v.arr.push_back(value()); ws(); if (p < end && *p == ',') { p++; continue; } if (p < end && *p == ']') { p++; break; } ok = false; }
        } else if (*p == '"') {
            v.kind = Json::Str; p++;
            while (p < end && *p != '"') { if (*p == '\\' && p + 1 < end) { p++; char c = *p; v.str.push_back(c == 'n' ? '\n' : c == 't' ? '\t' : c); if (c == 'u') p += 4; } else v.str.push_back(*p); p++; }
            if (p >= end) ok = false; else p++;
        } else if (!std::strncmp(p, "true", 4)) { v.kind = Json::Bool; v.b = true; p += 4; }
        else if (!std::strncmp(p, "false", 5)) { v.kind = Json::Bool; p += 5; }
        else if (!std::strncmp(p, "null", 4)) { p += 4; }
        else { char* e = nullptr; v.kind = Json::Num; v.num = std::strtod(p, &e); if (e == p) ok = false; p = e; }
        return v;
    }
};

// ---- binary glTF ----------------------------------------------------------------------------------------------------------
struct GlbAccessor { const uint8_t* data = nullptr; size_t count = 0, stride = 0; int componentType = 0, ncomp = 0; };
inline bool loadGLB(const std::string& path, Model& model, Material& material) {
    std::vector<uint8_t> raw;
    if (!readFile(path, raw) || raw.size() < 20) { std::cerr << "Failed to load GLTF: " << path << "\n"; return false; }
    auto u32at = [&](size_t o) { uint32_t v; std::memcpy(&v, &raw[o], 4); return v; };
    if (u32at(0) != 0x46546C67u || u32at(4) != 2u) { std::cerr << "loadGLB: not a glTF 2.0 binary: " << path << "\n"; return false; }
    const uint8_t* bin = nullptr; size_t binLen = 0; Json j; bool haveJson = false;
    for (size_t off = 12; off + 8 <= raw.size();) {
        uint32_t n = u32at(off), kind = u32at(off + 4); off += 8;
        if (off + n > raw.size()) break;
        if (kind == 0x4E4F534Au) { JsonParser ps{reinterpret_cast<const char*>(&raw[off]), reinterpret_cast<const char*>(&raw[off]) + n}; j = ps.value(); haveJson = ps.ok; }
        else if (kind == 0x004E4942u) { bin = &raw[off]; binLen = n; }
        off += n;
    }
    if (!haveJson || !bin) { std::cerr << "loadGLB: missing JSON or BIN chunk\n"; return false; }
    auto accessor = [&](int idx, GlbAccessor& a) {
        if (idx < 0) return false;
        const Json& ac = j["accessors"][size_t(idx)]; const Json& bv = j["bufferViews"][size_t(ac["bufferView"].integer(0))];
        a.componentType = ac["componentType"].integer(0);
        const std::string& t = ac["type"].str; a.ncomp = t == "SCALAR" ? 1 : t == "VEC2" ? 2 : t == "VEC3" ? 3 : t == "VEC4" ? 4 : 0;
        const size_t csz = (a.componentType == 5120 || a.componentType == 5121) ? 1 : (a.componentType == 5122 || a.componentType == 5123) ? 2 : 4;
        const size_t start = size_t(bv["byteOffset"].number(0)) + size_t(ac["byteOffset"].number(0));
        a.stride = size_t(bv["byteStride"].number(0)); if (!a.stride) a.stride = csz * size_t(a.ncomp);
        a.count = size_t(ac["count"].number(0)); a.data = bin + start;
        return a.ncomp > 0 && (a.count == 0 || start + (a.count - 1) * a.stride + csz * size_t(a.ncomp) <= binLen);
    };
    model.vertices.clear(); model.indices.clear();
    uint32_t vertexOffset = 0;
    for (const Json& mesh : j["meshes"].arr) for (const Json& prim : mesh["primitives"].arr) {
        const Json& at = prim["attributes"];
        GlbAccessor P, N, T, I;
        if (!at.has("POSITION") || !accessor(at["POSITION"].integer(-1), P) || P.componentType != 5126) continue;
        const bool hasN = at.has("NORMAL") && accessor(at["NORMAL"].integer(-1), N) && N.componentType == 5126;
        const bool hasT = at.has("TEXCOORD_0") && accessor(at["TEXCOORD_0"].integer(-1), T) && T.componentType == 5126;
        for (size_t i = 0; i < P.count; i++) {
            Vertex v{};
            std::memcpy(v.position, P.data + i * P.stride, 12);
            if (hasN) std::memcpy(v.normal, N.data + i * N.stride, 12); else { v.normal[1] = 1.0f; }
            if (hasT) std::memcpy(v.texCoord, T.data + i * T.stride, 8);
            model.vertices.push_back(v);
        }
        if (prim.has("indices") && accessor(prim["indices"].integer(-1), I)) {
            for (size_t i = 0; i < I.count / 3 * 3; i++) {
                uint32_t idx = 0; const uint8_t* q = I.data + i * I.stride;
                if (I.componentType == 5123) { uint16_t s; std::memcpy(&s, q, 2); idx = s; } else if (I.componentType == 5125) std::memcpy(&idx, q, 4); else idx = *q;
                model.indices.push_back(vertexOffset + idx);
            }
        } else for (size_t i = 0; i < P.count / 3 * 3; i++) model.indices.push_back(vertexOffset + uint32_t(i));
        vertexOffset += uint32_t(P.count);
    }
    if (model.vertices.empty()) { std::cerr << "loadGLB: no triangle geometry in " << path << "\n"; return false; }
    // first material: factors (materialColors / materialMetallic, model_gltf.cpp:373-386)
    material = Material{};
    if (!j["materials"].arr.empty()) {
        const Json& pbr = j["materials"][0]["pbrMetallicRoughness"];
        const Json& bc = pbr["baseColorFactor"];
        material.baseColor = {float(bc[0].number(1.0)), float(bc[1].number(1.0)), float(bc[2].number(1.0))};
        material.roughness = float(pbr["roughnessFactor"].number(1.0)); material.metallic = float(pbr["metallicFactor"].number(1.0));
    }
    // decoded texture layers from the side-car
    std::vector<uint8_t> tex;
    if (readFile(path + ".ohbtex", tex) && tex.size() >= 8 && !std::memcmp(tex.data(), "OHBT", 4)) {
        uint32_t n; std::memcpy(&n, &tex[4], 4); size_t off = 8;
        for (uint32_t k = 0; k < n && off + 12 <= tex.size(); k++) {
            uint32_t kind, w, h; std::memcpy(&kind, &tex[off], 4); std::memcpy(&w, &tex[off + 4], 4); std::memcpy(&h, &tex[off + 8], 4); off += 12;
            const size_t bytes = size_t(w) * h * 4; if (off + bytes > tex.size()) break;
            Image8 im; im.w = w; im.h = h; im.rgba.assign(tex.begin() + long(off), tex.begin() + long(off + bytes)); off += bytes;
            if (kind == 0) {
                // the base colour becomes the texture's mean (double sums / (count * 255), model_gltf.cpp:405-428)
                double s[3] = {0, 0, 0}; for (size_t p = 0; p < size_t(w) * h; p++) for (int c = 0; c < 3; c++) s[c] += im.rgba[p * 4 + c];
                const double d = double(w) * h * 255.0; material.baseColor = {float(s[0] / d), float(s[1] / d), float(s[2] / d)};
                material.albedoTex = std::move(im);
            } else if (kind == 1) material.normalTex = std::move(im); else if (kind == 2) material.roughMetalTex = std::move(im); else if (kind == 3) material.emissiveTex = std::move(im);
        }
    } else std::cerr << "loadGLB: no decoded textures (" << path << ".ohbtex) — rendering with the material factors only\n";
    std::cout << "GLTF loaded: " << path << " (" << model.vertices.size() << " vertices, " << model.indices.size() << " indices)\n";
    return true;
}

// Packed scene arrays as one flat file (test tooling): "OHBS", then {u32 tag, u64 bytes, payload} records.
inline bool dumpPackedScene(const SceneArrays& a, const std::string& path) {
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    auto rec = [&](uint32_t tag, const void* p, size_t bytes) { uint64_t n = bytes; f.write(reinterpret_cast<const char*>(&tag), 4); f.write(reinterpret_cast<const char*>(&n), 8); f.write(static_cast<const char*>(p), std::streamsize(bytes)); };
    f.write("OHBS", 4);
    rec(1, a.vertices.data(), a.vertices.size() * sizeof(Vertex)); rec(2, a.indices.data(), a.indices.size() * 4); rec(3, a.matIds.data(), a.matIds.size() * 4);
    rec(4, a.normals4.data(), a.normals4.size() * 4); rec(5, a.uvs2.data(), a.uvs2.size() * 4); rec(6, a.matColors.data(), a.matColors.size() * 4);
    rec(7, a.instances.data(), a.instances.size() * sizeof(ohb_instance)); rec(8, a.lightSSBO.data(), a.lightSSBO.size());
    uint32_t dims[3] = {a.texW, a.texH, a.layers}; rec(9, dims, 12); rec(10, a.texels.data(), a.texels.size());
    return bool(f);
}

}  // namespace ohao

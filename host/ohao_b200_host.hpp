// host/ohao_b200_host.hpp — C++20 host side above the C ABI (include/ohao_b200.h).
//
// Mirrors, with the reference's own names and call order, the part of the engine the three render-to-image
// examples touch, so they read like examples/cornell_box.cpp / turntable.cpp / inverse_fit.cpp:
//   ohao::Camera                (ohao/render/camera/camera.cpp:25-48)
//   ohao::Scene / Actor / Material / LightComponent data  (scene/scene.hpp:46,151; the unordered_map actor order
//                                is part of the data contract, quirk Q9)
//   packScene()                 = VulkanRenderer::updateSceneBuffers + buildAccelerationStructures hand-over
//                                (gpu/vulkan/rt_build.cpp:27-934, light_upload.cpp:153-293)
//   ohao::CudaRTRenderer        = the IRTRendererProfile method set (render/rt/rt_profile_renderer.hpp:7-86)
//   ohao::Renderer              = VulkanRenderer's RT seam (gpu/vulkan/renderer.hpp:114-300)
// Error behaviour is the reference's: bool returns + std::cerr.  No Vulkan types (this image has no Vulkan headers);
// INTEGRATION.md shows the in-tree variant deriving from IRTRendererProfile.
#pragma once
#include "../include/ohao_b200.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <span>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace ohao {

// ---- glm-compatible fp32 math (column-major mat4) -------------------------------------------------------------
struct vec3 { float x = 0, y = 0, z = 0; };
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline vec3 normalize(vec3 a) { float l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }
struct mat4 { float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; float& at(int col, int row) { return m[col * 4 + row]; } };
inline mat4 lookAt(vec3 eye, vec3 center, vec3 up) {      // glm::lookAt (RH)
    vec3 f = normalize(center - eye), s = normalize(cross(f, up)), u = cross(s, f);
    mat4 r;
    r.at(0, 0) = s.x; r.at(1, 0) = s.y; r.at(2, 0) = s.z;
    r.at(0, 1) = u.x; r.at(1, 1) = u.y; r.at(2, 1) = u.z;
    r.at(0, 2) = -f.x; r.at(1, 2) = -f.y; r.at(2, 2) = -f.z;
    r.at(3, 0) = -dot(s, eye); r.at(3, 1) = -dot(u, eye); r.at(3, 2) = dot(f, eye);
    return r;
}
inline mat4 perspective(float fovy, float aspect, float zn, float zf) {   // glm::perspective (RH, -1..1)
    float t = float(std::tan(double(fovy) / 2.0));
    mat4 r; std::memset(r.m, 0, sizeof(r.m));
    r.at(0, 0) = 1.0f / (aspect * t); r.at(1, 1) = 1.0f / t;
    r.at(2, 2) = -(zf + zn) / (zf - zn); r.at(2, 3) = -1.0f; r.at(3, 2) = -(2.0f * zf * zn) / (zf - zn);
    return r;
}
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }

// ---- Camera (camera.cpp:25-48) ----------------------------------------------------------------------------------
class Camera {
public:
    void setPosition(vec3 p) { m_pos = p; }
    void setRotation(float pitch, float yaw) { m_pitch = pitch; m_yaw = yaw; }
    void setFov(float f) { m_fov = f; }
    vec3 getPosition() const { return m_pos; }
    float getFov() const { return m_fov; }
    mat4 getViewMatrix() const {
        float yaw = radians(m_yaw), pitch = radians(m_pitch);
        vec3 front = normalize({std::cos(yaw) * std::cos(pitch), std::sin(pitch), std::sin(yaw) * std::cos(pitch)});
        vec3 right = normalize(cross(front, {0, 1, 0}));
        vec3 up = normalize(cross(right, front));
        return lookAt(m_pos, m_pos + front, up);
    }
    mat4 getProjection(uint32_t w, uint32_t h) const { return perspective(radians(m_fov), float(w) / float(h), 0.1f, 1000.0f); }   // render_dispatch.cpp:158-163
private:
    vec3 m_pos{0, 0, 2.5f}; float m_yaw = -90.0f, m_pitch = 0.0f, m_fov = 45.0f;
};

// ---- Scene data (what the packers read) -----------------------------------------------------------------------
struct Vertex {                      // scene/asset/model.hpp:17-35: 100 bytes, position at 0, normal at 24
    float position[3]; float color[3]; float normal[3]; float texCoord[2]; float rest[14];
};
static_assert(sizeof(Vertex) == 100, "Vertex must be 100 bytes like the reference's");
struct Image8 { uint32_t w = 0, h = 0; std::vector<uint8_t> rgba; bool empty() const { return rgba.empty(); } };
struct Material { vec3 baseColor{0.8f, 0.8f, 0.8f}; float roughness = 0.5f, metallic = 0.0f; Image8 albedoTex, normalTex, roughMetalTex, emissiveTex; };
struct Model { std::vector<Vertex> vertices; std::vector<uint32_t> indices; };
enum class LightType { Sphere = 0, Directional = 1, Spot = 2, AreaRect = 3 };
struct Light { LightType type = LightType::Sphere; vec3 color{1, 1, 1}; float intensity = 1.0f, radius = 0.5f; vec3 direction{0, -1, 0}; float innerCone = 30, outerCone = 45; vec3 edge1{1, 0, 0}, edge2{0, 0, 1}; };
struct Actor {
    std::string name; vec3 position{0, 0, 0}, scale{1, 1, 1}; bool visible = true;
    bool yaw180 = false;             // TransformComponent rotation quat(radians(0, 180, 0)) — the only rotation the examples set (turntable.cpp:133)
    std::shared_ptr<Model> model; Material material; bool hasLight = false; Light light;
};
class Scene {
public:
    explicit Scene(std::string name = "Scene") : m_name(std::move(name)) { createActor("World"); }   // Scene ctor creates the root actor (scene.cpp:27-29)
    Actor* createActor(std::string_view name) { uint64_t id = m_next++; auto& a = actors[id]; a = std::make_unique<Actor>(); a->name = std::string(name); return a.get(); }
    std::unordered_map<uint64_t, std::unique_ptr<Actor>> actors;      // iteration order == BLAS / matID / layer / light order (Q9)
private:
    std::string m_name; uint64_t m_next = 1;
};

// addQuad (examples/cornell_box.cpp:32-53)
inline void addQuad(Model& m, vec3 a, vec3 b, vec3 c, vec3 d, vec3 n, vec3 color) {
    uint32_t base = uint32_t(m.vertices.size());
    for (vec3 p : {a, b, c, d}) { Vertex v{}; v.position[0] = p.x; v.position[1] = p.y; v.position[2] = p.z; v.normal[0] = n.x; v.normal[1] = n.y; v.normal[2] = n.z; v.color[0] = color.x; v.color[1] = color.y; v.color[2] = color.z; m.vertices.push_back(v); }
    for (uint32_t i : {0u, 1u, 2u, 0u, 2u, 3u}) m.indices.push_back(base + i);
}
// ComponentFactory::generateSphereMesh (scene/component/component_factory.cpp:347-399)
inline std::shared_ptr<Model> generateSphereMesh(int sectors = 32, int stacks = 16, float radius = 0.5f) {
    auto m = std::make_shared<Model>();
    const float pi = 3.14159265358979323846f;
    for (int i = 0; i <= stacks; i++) {
        float phi = pi * float(i) / float(stacks);
        for (int j = 0; j <= sectors; j++) {
            float theta = 2.0f * pi * float(j) / float(sectors);
            float x = std::cos(theta) * std::sin(phi), y = std::cos(phi), z = std::sin(theta) * std::sin(phi);
            Vertex v{}; v.position[0] = x * radius; v.position[1] = y * radius; v.position[2] = z * radius;
            v.normal[0] = x; v.normal[1] = y; v.normal[2] = z; v.texCoord[0] = float(j) / float(sectors); v.texCoord[1] = float(i) / float(stacks);
            m->vertices.push_back(v);
        }
    }
    for (int a = 0; a < stacks; a++) for (int b = 0; b < sectors; b++) {
        uint32_t first = uint32_t(a * (sectors + 1) + b), second = first + uint32_t(sectors) + 1;
        for (uint32_t i : {first, second, first + 1, second, second + 1, first + 1}) m->indices.push_back(i);
    }
    return m;
}

// ---- packed arrays == the §3.2 data contract ---------------------------------------------------------------------
struct SceneArrays {
    std::vector<Vertex> vertices; std::vector<uint32_t> indices, matIds; std::vector<float> normals4, uvs2, matColors;
    std::vector<ohb_instance> instances; std::vector<uint8_t> texels; uint32_t texW = 1, texH = 1, layers = 0;
    std::vector<uint8_t> lightSSBO; std::vector<Actor*> meshActors, lightActors;
};
inline uint8_t linearToSrgb8(float v) {    // rt_build.cpp:462-465
    float s = v <= 0.0031308f ? v * 12.92f : 1.055f * float(std::pow(double(v), 1.0 / 2.4)) - 0.055f;
    s = std::min(std::max(s, 0.0f), 1.0f);
    return uint8_t(int(s * 255.0f + 0.5f));
}
inline Image8 resizeRGBA8Bilinear(const Image8& src, uint32_t dw, uint32_t dh) {   // rt_build.cpp:548-590
    if (src.w == dw && src.h == dh) return src;
    Image8 o; o.w = dw; o.h = dh; o.rgba.resize(size_t(dw) * dh * 4);
    float sx = float(src.w) / float(dw), sy = float(src.h) / float(dh);
    for (uint32_t y = 0; y < dh; y++) {
        float fy = std::min(std::max((float(y) + 0.5f) * sy - 0.5f, 0.0f), float(src.h - 1));
        uint32_t y0 = uint32_t(std::floor(fy)), y1 = std::min(y0 + 1, src.h - 1); float ty = fy - float(y0);
        for (uint32_t x = 0; x < dw; x++) {
            float fx = std::min(std::max((float(x) + 0.5f) * sx - 0.5f, 0.0f), float(src.w - 1));
            uint32_t x0 = uint32_t(std::floor(fx)), x1 = std::min(x0 + 1, src.w - 1); float tx = fx - float(x0);
            for (int c = 0; c < 4; c++) {
                auto px = [&](uint32_t xx, uint32_t yy) { return float(src.rgba[(size_t(yy) * src.w + xx) * 4 + c]); };
                float top = px(x0, y0) + (px(x1, y0) - px(x0, y0)) * tx, bot = px(x0, y1) + (px(x1, y1) - px(x0, y1)) * tx;
                o.rgba[(size_t(y) * dw + x) * 4 + c] = uint8_t(std::min(std::max(top + (bot - top) * ty, 0.0f), 255.0f));
            }
        }
    }
    return o;
}
inline float bitsToFloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline void packLight(const Actor& a, uint8_t* out) {     // GPULight, 80 B (render/rt/gpu_light.hpp:10-17, light_upload.cpp:157-181)
    const Light& l = a.light; float v[20] = {0};
    v[0] = a.position.x; v[1] = a.position.y; v[2] = a.position.z; v[3] = float(int(l.type));
    v[4] = l.color.x; v[5] = l.color.y; v[6] = l.color.z; v[7] = l.intensity;
    v[8] = l.direction.x; v[9] = l.direction.y; v[10] = l.direction.z; v[11] = l.type == LightType::Spot ? l.innerCone : l.radius;
    if (l.type == LightType::AreaRect) { vec3 c = cross(l.edge1, l.edge2); v[12] = l.edge1.x; v[13] = l.edge1.y; v[14] = l.edge1.z; v[16] = l.edge2.x; v[17] = l.edge2.y; v[18] = l.edge2.z; v[19] = std::sqrt(dot(c, c)); }
    else v[15] = l.outerCone;
    std::memcpy(out, v, 80);
}
// One auto-generated sphere light per mesh actor whose material has an emissive texture (light_upload.cpp:183-247):
// centre = world position of the model's bounding-box centre, radius = 0.3 x |bbox diagonal| (object space, like the
// reference), colour = mean of the "bright" texels (luminance > 0.05), intensity = min(0.1 x summed luminance, 20).
inline bool emissiveMeshLight(const Actor& act, uint8_t* out) {
    const Image8& e = act.material.emissiveTex;
    if (!act.model || e.empty()) return false;
    vec3 bmin{FLT_MAX, FLT_MAX, FLT_MAX}, bmax{-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const Vertex& v : act.model->vertices) {
        bmin = vec3{std::min(bmin.x, v.position[0]), std::min(bmin.y, v.position[1]), std::min(bmin.z, v.position[2])};
        bmax = vec3{std::max(bmax.x, v.position[0]), std::max(bmax.y, v.position[1]), std::max(bmax.z, v.position[2])};
    }
    const vec3 mid{(bmin.x + bmax.x) * 0.5f, (bmin.y + bmax.y) * 0.5f, (bmin.z + bmax.z) * 0.5f}, d{bmax.x - bmin.x, bmax.y - bmin.y, bmax.z - bmin.z};
    const float ry = act.yaw180 ? -1.0f : 1.0f;
    const vec3 center{ry * act.scale.x * mid.x + act.position.x, act.scale.y * mid.y + act.position.y, ry * act.scale.z * mid.z + act.position.z};
    const float radius = std::sqrt(dot(d, d)) * 0.3f;
    double r = 0, g = 0, b = 0; int bright = 0; float power = 0.0f;
    for (size_t p = 0; p < size_t(e.w) * e.h; p++) {
        const float pr = e.rgba[p * 4 + 0] / 255.0f, pg = e.rgba[p * 4 + 1] / 255.0f, pb = e.rgba[p * 4 + 2] / 255.0f;
        const float lum = pr * 0.2126f + pg * 0.7152f + pb * 0.0722f;
        if (lum > 0.05f) { r += pr; g += pg; b += pb; bright++; power += lum; }
    }
    if (!(power > 0.1f)) return false;
    float v[20] = {0};
    v[0] = center.x; v[1] = center.y; v[2] = center.z; v[3] = 0.0f;                                   // sphere
    v[4] = bright ? float(r / bright) : 1.0f; v[5] = bright ? float(g / bright) : 1.0f; v[6] = bright ? float(b / bright) : 1.0f;
    v[7] = std::min(power * 0.1f, 20.0f);
    v[8] = 0.0f; v[9] = -1.0f; v[10] = 0.0f; v[11] = radius;
    std::memcpy(out, v, 80);
    return true;
}
inline void packLights(SceneArrays& a, bool haveEnv, float envIntensity) {    // light_upload.cpp:153-283
    std::vector<uint8_t> recs(a.lightActors.size() * 80);
    for (size_t i = 0; i < a.lightActors.size(); i++) packLight(*a.lightActors[i], recs.data() + i * 80);
    for (Actor* m : a.meshActors) {                                             // after the explicit lights, in actor order
        uint8_t rec[80];
        if (emissiveMeshLight(*m, rec)) recs.insert(recs.end(), rec, rec + 80);
    }
    const size_t n = recs.size() / 80;
    a.lightSSBO.assign(16 + n * 80, 0xFF);
    uint32_t cnt = uint32_t(n); std::memcpy(a.lightSSBO.data(), &cnt, 4); std::memcpy(a.lightSSBO.data() + 8, &envIntensity, 4);
    if (haveEnv && n > 0) { uint32_t idx = a.layers; std::memcpy(a.lightSSBO.data() + 4, &idx, 4); }    // env only wired when >= 1 light exists (Q13)
    if (n) std::memcpy(a.lightSSBO.data() + 16, recs.data(), recs.size());
}
inline void packMaterials(SceneArrays& a, const std::vector<std::array<int, 4>>& layerIdx) {   // rt_build.cpp:188-389
    a.matColors.clear();
    for (size_t i = 0; i < a.meshActors.size(); i++) {
        const Material& m = a.meshActors[i]->material;
        auto bits = [](int idx) { return bitsToFloat(idx < 0 ? OHB_NO_TEXTURE : uint32_t(idx)); };
        float rec[12] = {m.baseColor.x, m.baseColor.y, m.baseColor.z, bits(layerIdx[i][0]), m.roughness, m.metallic, bits(layerIdx[i][1]), bits(layerIdx[i][3]), bits(layerIdx[i][2]), 0, 0, 0};
        a.matColors.insert(a.matColors.end(), rec, rec + 12);
    }
}
struct PackedScene { SceneArrays arrays; std::vector<std::array<int, 4>> layerIdx; };
inline PackedScene packScene(Scene& scene, bool haveEnv, float envIntensity) {
    PackedScene ps; SceneArrays& a = ps.arrays;
    std::vector<Image8> layers;
    uint32_t voff = 0, toff = 0;
    for (auto& [id, up] : scene.actors) {                    // unordered_map order, like the reference (Q9)
        Actor* act = up.get();
        if (act->hasLight) a.lightActors.push_back(act);
        if (!act->model || !act->visible) continue;
        const Model& m = *act->model;
        uint32_t nt = uint32_t(m.indices.size() / 3), mi = uint32_t(a.meshActors.size());
        a.meshActors.push_back(act);
        for (const Vertex& v : m.vertices) {
            a.vertices.push_back(v);
            a.normals4.insert(a.normals4.end(), {v.normal[0], v.normal[1], v.normal[2], 0.0f});
            a.uvs2.insert(a.uvs2.end(), {v.texCoord[0], v.texCoord[1]});
        }
        for (uint32_t i : m.indices) a.indices.push_back(i + voff);
        a.matIds.insert(a.matIds.end(), nt, mi);
        ohb_instance in{}; in.first_tri = toff; in.tri_count = nt; in.mask = 0xFF;
        const float ry = act->yaw180 ? -1.0f : 1.0f;                       // T * R_y(180) * S
        float x[12] = {ry * act->scale.x, 0, 0, act->position.x, 0, act->scale.y, 0, act->position.y, 0, 0, ry * act->scale.z, act->position.z};
        std::memcpy(in.xform, x, 48);
        a.instances.push_back(in);
        auto layer = [&](const Image8& t) -> int { if (t.empty()) return -1; layers.push_back(t); return int(layers.size()) - 1; };
        std::array<int, 4> li{};
        const Material& mat = act->material;
        if (!mat.albedoTex.empty()) li[0] = layer(mat.albedoTex);
        else { Image8 solid; solid.w = solid.h = 1; solid.rgba = {linearToSrgb8(mat.baseColor.x), linearToSrgb8(mat.baseColor.y), linearToSrgb8(mat.baseColor.z), 255}; li[0] = layer(solid); }   // quirk Q1
        li[1] = layer(mat.normalTex); li[2] = layer(mat.roughMetalTex); li[3] = layer(mat.emissiveTex);
        ps.layerIdx.push_back(li);
        voff += uint32_t(m.vertices.size()); toff += nt;
    }
    uint32_t tw = 1, th = 1;
    for (auto& l : layers) { tw = std::max(tw, l.w); th = std::max(th, l.h); }
    tw = std::min(tw, 2048u); th = std::min(th, 2048u);                       // rt_build.cpp:533-546
    a.texW = tw; a.texH = th; a.layers = uint32_t(layers.size());
    for (auto& l : layers) { Image8 r = resizeRGBA8Bilinear(l, tw, th); a.texels.insert(a.texels.end(), r.rgba.begin(), r.rgba.end()); }
    packMaterials(a, ps.layerIdx);
    packLights(a, haveEnv, envIntensity);
    return ps;
}

// ---- settings ------------------------------------------------------------------------------------------------------
enum class RTRenderProfile { Offline = 0, Realtime = 1 };
enum class RenderMode { RTOffline, RTRealtime };
// denoise/denoise_types.hpp:13-19, same values.  Built on this path: None and Atrous (the SVGF denoiser, realtime profile);
// OIDN / NRD / DLSS-RR are third-party libraries (OIDN is the open §8f "next" row) and are refused by ohb_set_settings.
enum class DenoiseMode : uint32_t { None = 0, OIDN = 1, NRD = 3, Atrous = 4, DLSSRR = 5 };
struct RTRenderSettings {                            // rt_settings.hpp:17-35
    RTRenderProfile profile = RTRenderProfile::Offline; uint32_t maxBounces = 4; bool enableAuxiliaryAOVs = true, enableInternalDenoise = false, enableFireflyClamp = false;
    float fireflyClampLuminance = 0.0f, anisotropyStrength = 0.0f, anisotropyRotation = 0.0f, subsurfaceStrength = 0.0f; uint32_t samplesPerFrame = 1;
    DenoiseMode denoiseMode = DenoiseMode::None;
};
inline constexpr RTRenderSettings kOfflineRTSettings{RTRenderProfile::Offline, 4, true, false, false, 0.0f, 0, 0, 0, 1};
inline constexpr RTRenderSettings kRealtimeRTSettings{RTRenderProfile::Realtime, 2, true, true, true, 10.0f, 0, 0, 0, 1};
inline ohb_settings toOhb(const RTRenderSettings& s) {
    ohb_settings o{}; o.profile = uint32_t(s.profile); o.max_bounces = s.maxBounces;
    o.flags = (s.enableAuxiliaryAOVs ? OHB_FLAG_ENABLE_AOVS : 0u) | (s.enableInternalDenoise ? OHB_FLAG_ENABLE_INTERNAL_DENOISE : 0u) | (s.enableFireflyClamp ? OHB_FLAG_ENABLE_FIREFLY_CLAMP : 0u);
    o.firefly_clamp_lum = s.fireflyClampLuminance; o.sampler_type = OHB_SAMPLER_SOBOL;     // both profiles run Sobol (Q3)
    o.anisotropy_strength = s.anisotropyStrength; o.anisotropy_rotation = s.anisotropyRotation; o.subsurface_strength = s.subsurfaceStrength;
    o.samples_per_frame = std::min(std::max(s.samplesPerFrame, 1u), 64u);                 // clampSamplesPerFrame
    o.denoise_mode = uint32_t(s.denoiseMode);
    return o;
}

// ---- IRTRendererProfile-shaped adapter -------------------------------------------------------------------------------
class CudaRTRenderer {
public:
    explicit CudaRTRenderer(RTRenderProfile p, int device = 0) : m_profile(p), m_device(device) {}
    ~CudaRTRenderer() { destroy(); }
    const char* getName() const { return m_profile == RTRenderProfile::Offline ? "CudaRTOffline(B200)" : "CudaRTRealtime(B200)"; }
    RTRenderProfile getProfile() const { return m_profile; }
    RTRenderSettings getDefaultSettings() const { return m_profile == RTRenderProfile::Offline ? kOfflineRTSettings : kRealtimeRTSettings; }
    [[nodiscard]] bool init(uint32_t width, uint32_t height) {
        m_ctx = ohb_create(m_device, width, height, int(m_profile));
        if (!m_ctx) { std::cerr << "[CudaRT] init failed: " << ohb_last_error(nullptr) << "\n"; return false; }
        setRenderSettings(getDefaultSettings());
        return true;
    }
    void destroy() { if (m_ctx) ohb_destroy(m_ctx); m_ctx = nullptr; }
    void resize(uint32_t w, uint32_t h) { check(ohb_resize(m_ctx, w, h), "resize"); }
    void render(const mat4& view, const mat4& proj, uint32_t nsamples = 1) { if (m_ctx) check(ohb_render(m_ctx, view.m, proj.m, nsamples), "render"); }
    void setRenderSettings(const RTRenderSettings& s) { ohb_settings o = toOhb(s); check(ohb_set_settings(m_ctx, &o), "setRenderSettings"); }
    void notifyViewChanged() { ohb_notify_view_changed(m_ctx); }
    void resetAccumulation() { ohb_reset_accumulation(m_ctx); }
    void setRenderSeed(uint32_t seed) { ohb_set_seed(m_ctx, seed); }
    uint32_t getFrameIndex() const { return ohb_frame_index(m_ctx); }
    bool resetsAccumulationOnViewChange() const { return m_profile == RTRenderProfile::Offline; }
    [[nodiscard]] bool uploadScene(const SceneArrays& a, const float* envRGBA32F, uint32_t envW, uint32_t envH) {
        return check(ohb_set_geometry(m_ctx, a.vertices.data(), sizeof(Vertex), uint32_t(a.vertices.size()), a.indices.data(), uint32_t(a.matIds.size()), a.normals4.data(), a.uvs2.data(), a.matIds.data()), "setGeometry")
            && check(ohb_set_instances(m_ctx, a.instances.data(), uint32_t(a.instances.size())), "setInstances")
            && check(ohb_set_materials(m_ctx, a.matColors.data(), uint32_t(a.matColors.size() / 12)), "setMaterials")
            && check(ohb_set_textures(m_ctx, a.texels.data(), a.texW, a.texH, a.layers), "setTextures")
            && check(ohb_set_lights(m_ctx, a.lightSSBO.data(), a.lightSSBO.size()), "setLightBuffer")
            && check(ohb_set_env(m_ctx, envRGBA32F, envW, envH), "setEnvironmentMap")
            && check(ohb_build_accel(m_ctx), "buildTLAS");
    }
    [[nodiscard]] bool setMaterialData(const std::vector<float>& mc) { return check(ohb_set_materials(m_ctx, mc.data(), uint32_t(mc.size() / 12)), "setMaterialData"); }
    [[nodiscard]] bool setLightBuffer(const std::vector<uint8_t>& ssbo) { return check(ohb_set_lights(m_ctx, ssbo.data(), ssbo.size()), "setLightBuffer"); }
    // RTAccelerationStructure: how BLAS / TLAS are realised, and buildTLAS in MODE_UPDATE (rt_acceleration_structure.cpp:467,503-512)
    [[nodiscard]] bool setAccelMode(bool twoLevel) { return check(ohb_set_accel_mode(m_ctx, twoLevel ? OHB_ACCEL_TWO_LEVEL : OHB_ACCEL_FLATTEN), "setAccelMode"); }
    [[nodiscard]] bool updateTLAS(const std::vector<ohb_instance>& inst) { return check(ohb_update_instances(m_ctx, inst.data(), uint32_t(inst.size())), "buildTLAS(update)"); }
    [[nodiscard]] bool readLDR(uint8_t* rgba8) { return check(ohb_read_ldr(m_ctx, rgba8), "getPixels"); }
    [[nodiscard]] bool readHDR(float* beauty, float* albedo, float* normal) { return check(ohb_read_hdr(m_ctx, beauty, albedo, normal), "readbackHDRBuffers"); }
    ohb_ctx* ctx() const { return m_ctx; }
private:
    bool check(int rc, const char* what) const { if (rc) std::cerr << "[CudaRT] " << what << ": " << ohb_last_error(m_ctx) << "\n"; return rc == 0; }
    RTRenderProfile m_profile; int m_device; ohb_ctx* m_ctx = nullptr;
};

// ---- VulkanRenderer's RT seam ------------------------------------------------------------------------------------------
class Renderer {
public:
    Renderer(uint32_t w, uint32_t h, int device = 0) : m_w(w), m_h(h), m_device(device) {}
    [[nodiscard]] bool initialize() { return true; }                       // the CUDA context is created lazily by ensureRTRenderer, like the reference
    void setScene(Scene* s) { m_scene = s; m_sceneDirty = true; }
    void setRenderMode(RenderMode m) { if (m != m_mode) { m_mode = m; m_rt.reset(); m_sceneDirty = true; } }
    void setDenoiseMode(DenoiseMode d) { m_denoise = d; m_settingsDirty = true; }
    DenoiseMode getDenoiseMode() const { return m_denoise; }
    void setEnvironmentMap(std::vector<float> rgba32f, uint32_t w, uint32_t h) { m_env = std::move(rgba32f); m_envW = w; m_envH = h; m_sceneDirty = true; }
    void setEnvIntensityScale(float s) { m_envIntensity = s; m_lightsDirty = true; }
    void setRenderSeed(uint32_t seed) { m_seed = seed; if (m_rt) m_rt->setRenderSeed(seed); }
    void resetAccumulation() { if (m_rt) m_rt->resetAccumulation(); }
    void notifyCameraChanged() { if (m_rt) m_rt->notifyViewChanged(); }
    void setRealtimeSamplesPerFrame(uint32_t n) { m_spf = n; m_settingsDirty = true; }
    void resize(uint32_t w, uint32_t h) { m_w = w; m_h = h; if (m_rt) m_rt->resize(w, h); m_pixels.clear(); }
    Camera& getCamera() { return m_camera; }
    uint32_t width() const { return m_w; } uint32_t height() const { return m_h; }
    [[nodiscard]] bool updateSceneBuffers() {
        if (!m_scene || !ensureRTRenderer()) return false;
        m_packed = packScene(*m_scene, !m_env.empty(), m_envIntensity);
        m_sceneDirty = false; m_lightsDirty = false;
        if (!m_rt->setAccelMode(m_twoLevel)) return false;
        return m_rt->uploadScene(m_packed.arrays, m_env.empty() ? nullptr : m_env.data(), m_envW, m_envH);
    }
    [[nodiscard]] bool updateRTMaterialParams() {                             // no BVH rebuild (render_session.hpp:49-56)
        if (!m_rt || m_sceneDirty) return updateSceneBuffers();
        packMaterials(m_packed.arrays, m_packed.layerIdx);
        return m_rt->setMaterialData(m_packed.arrays.matColors);
    }
    // Actors moved (position / scale / yaw180 only): refit instead of rebuilding — the reference's per-frame buildTLAS(MODE_UPDATE).
    // The mesh actors, their order and their triangle ranges must be those of the last updateSceneBuffers().
    [[nodiscard]] bool updateInstanceTransforms() {
        if (!m_rt || m_sceneDirty) return updateSceneBuffers();
        SceneArrays& a = m_packed.arrays;
        for (size_t i = 0; i < a.meshActors.size() && i < a.instances.size(); i++) {
            const Actor* act = a.meshActors[i]; const float ry = act->yaw180 ? -1.0f : 1.0f;
            float x[12] = {ry * act->scale.x, 0, 0, act->position.x, 0, act->scale.y, 0, act->position.y, 0, 0, ry * act->scale.z, act->position.z};
            std::memcpy(a.instances[i].xform, x, 48);
        }
        m_rt->resetAccumulation();
        packLights(a, !m_env.empty(), m_envIntensity);                       // emissive-mesh auto lights follow their actors
        return m_rt->updateTLAS(a.instances) && m_rt->setLightBuffer(a.lightSSBO);
    }
    void setTwoLevelAccel(bool on) { m_twoLevel = on; m_sceneDirty = true; }
    [[nodiscard]] bool updateRTLightParams() {
        if (!m_rt || m_sceneDirty) return updateSceneBuffers();
        packLights(m_packed.arrays, !m_env.empty(), m_envIntensity); m_lightsDirty = false;
        return m_rt->setLightBuffer(m_packed.arrays.lightSSBO);
    }
    // One call == one accumulation step (offline: 1 spp).  `nsamples` batches that many consecutive steps in one submission.
    void render(uint32_t nsamples = 1) {
        if (m_sceneDirty && !updateSceneBuffers()) return;
        if (m_lightsDirty && !updateRTLightParams()) return;
        if (m_settingsDirty) {
            RTRenderSettings s = m_rt->getDefaultSettings(); s.samplesPerFrame = m_spf;
            if (m_denoise == DenoiseMode::None) s.enableInternalDenoise = false;
            // applyRTRenderSettings (renderer.cpp:482-493): the Atrous override is injected into the tracer's settings (realtime profile only here)
            if (m_denoise == DenoiseMode::Atrous && m_mode == RenderMode::RTRealtime) s.denoiseMode = DenoiseMode::Atrous;
            m_rt->setRenderSettings(s); m_settingsDirty = false;
        }
        m_rt->render(m_camera.getViewMatrix(), m_camera.getProjection(m_w, m_h), nsamples);
    }
    std::span<const uint8_t> getPixelSpan() {                                // RGBA8, top row first; valid immediately (no 3-frame ring lag)
        m_pixels.resize(size_t(m_w) * m_h * 4);
        if (!m_rt || !m_rt->readLDR(m_pixels.data())) return {};
        return m_pixels;
    }
    [[nodiscard]] bool readbackHDRBuffers(std::vector<float>& beauty, std::vector<float>& albedo, std::vector<float>& normal) {
        size_t n = size_t(m_w) * m_h * 4; beauty.resize(n); albedo.resize(n); normal.resize(n);
        return m_rt && m_rt->readHDR(beauty.data(), albedo.data(), normal.data());
    }
    CudaRTRenderer* rtRenderer() { return m_rt.get(); }
private:
    bool ensureRTRenderer() {                                                // renderer.cpp:402-444
        if (m_rt) return true;
        m_rt = std::make_unique<CudaRTRenderer>(m_mode == RenderMode::RTOffline ? RTRenderProfile::Offline : RTRenderProfile::Realtime, m_device);
        if (!m_rt->init(m_w, m_h)) { m_rt.reset(); return false; }
        m_rt->setRenderSeed(m_seed); m_settingsDirty = true;
        return true;
    }
    uint32_t m_w, m_h; int m_device; Scene* m_scene = nullptr; Camera m_camera; RenderMode m_mode = RenderMode::RTOffline; DenoiseMode m_denoise = DenoiseMode::None;
    std::unique_ptr<CudaRTRenderer> m_rt; PackedScene m_packed; std::vector<float> m_env; uint32_t m_envW = 0, m_envH = 0; float m_envIntensity = 1.0f;
    uint32_t m_seed = 0, m_spf = 1; bool m_sceneDirty = true, m_lightsDirty = false, m_settingsDirty = true, m_twoLevel = false; std::vector<uint8_t> m_pixels;
};

// ---- image output: PNG with stored (uncompressed) deflate blocks, no dependencies -----------------------------------------
inline bool writePNG(const std::string& path, const uint8_t* rgba, uint32_t w, uint32_t h) {
    auto crcTable = [] { std::array<uint32_t, 256> t{}; for (uint32_t n = 0; n < 256; n++) { uint32_t c = n; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; t[n] = c; } return t; }();
    auto crc = [&](const uint8_t* p, size_t n, uint32_t c = 0xFFFFFFFFu) { for (size_t i = 0; i < n; i++) c = crcTable[(c ^ p[i]) & 255] ^ (c >> 8); return c; };
    std::vector<uint8_t> raw; raw.reserve((size_t(w) * 4 + 1) * h);
    for (uint32_t y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), rgba + size_t(y) * w * 4, rgba + size_t(y + 1) * w * 4); }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (size_t off = 0; off < raw.size();) {
        size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n == raw.size() ? 1 : 0); z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8)); z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
        z.insert(z.end(), raw.begin() + long(off), raw.begin() + long(off + n));
        for (size_t i = 0; i < n; i++) { a = (a + raw[off + i]) % 65521u; b = (b + a) % 65521u; }
        off += n;
    }
    uint32_t adler = (b << 16) | a; for (int s = 24; s >= 0; s -= 8) z.push_back(uint8_t(adler >> s));
    FILE* f = std::fopen(path.c_str(), "wb"); if (!f) return false;
    auto be = [](uint32_t v, uint8_t* o) { o[0] = uint8_t(v >> 24); o[1] = uint8_t(v >> 16); o[2] = uint8_t(v >> 8); o[3] = uint8_t(v); };
    auto chunk = [&](const char* type, const std::vector<uint8_t>& data) {
        uint8_t len[4]; be(uint32_t(data.size()), len); std::fwrite(len, 1, 4, f);
        std::vector<uint8_t> td(type, type + 4); td.insert(td.end(), data.begin(), data.end()); std::fwrite(td.data(), 1, td.size(), f);
        uint8_t c[4]; be(crc(td.data(), td.size()) ^ 0xFFFFFFFFu, c); std::fwrite(c, 1, 4, f);
    };
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}; std::fwrite(sig, 1, 8, f);
    std::vector<uint8_t> ihdr(13); be(w, &ihdr[0]); be(h, &ihdr[4]); ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = ihdr[11] = ihdr[12] = 0;
    chunk("IHDR", ihdr); chunk("IDAT", z); chunk("IEND", {});
    std::fclose(f);
    return true;
}

// ---- tiny CLI helpers shared by the examples (examples/example_cli.hpp) ---------------------------------------------------
inline std::string argOr(int argc, char** argv, int i, const char* def) { return (i < argc && argv[i][0] != '-') ? argv[i] : def; }
inline long flagValue(int argc, char** argv, const char* name, long def) {
    std::string key = std::string("--") + name;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == key && i + 1 < argc) return std::atol(argv[i + 1]);
        if (a.rfind(key + "=", 0) == 0) return std::atol(a.c_str() + key.size() + 1);
    }
    return def;
}
inline std::string flagString(int argc, char** argv, const char* name, const char* def) {
    std::string key = std::string("--") + name;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == key && i + 1 < argc) return argv[i + 1];
        if (a.rfind(key + "=", 0) == 0) return a.substr(key.size() + 1);
    }
    return def;
}
// --denoise=<mode> of the reference examples (parseDenoiseMode, denoise/denoise_types.cpp): none | atrous (alias svgf);
// oidn / nrd / dlssrr parse but are refused by the library (ohb_set_settings) when the renderer applies them.
inline DenoiseMode denoiseFlag(int argc, char** argv, DenoiseMode def) {
    const std::string key = "--denoise=";
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a.rfind(key, 0) != 0) continue;
        std::string v = a.substr(key.size());
        if (v == "none") return DenoiseMode::None;
        if (v == "atrous" || v == "svgf") return DenoiseMode::Atrous;
        if (v == "oidn") return DenoiseMode::OIDN;
        if (v == "nrd") return DenoiseMode::NRD;
        if (v == "dlssrr") return DenoiseMode::DLSSRR;
        std::cerr << "unknown --denoise mode '" << v << "', using the default\n";
    }
    return def;
}
inline bool hasFlag(int argc, char** argv, const char* name) { std::string key = std::string("--") + name; for (int i = 1; i < argc; i++) if (key == argv[i]) return true; return false; }

}  // namespace ohao

// host/procedural.hpp — seeded stand-ins for the assets that cannot travel to the GPU box (DamagedHelmet.glb,
// env_outdoor.hdr / env_studio.hdr): a displaced UV sphere of a requested triangle count with UVs + smooth normals,
// procedural PBR textures, and an outdoor-HDRI-like equirect image.  Asset loading itself (glTF, RGBE) is out of scope
// (SURVEY §8f row 3); these builders produce INPUTS in the same data contract.
#pragma once
#include "ohao_b200_host.hpp"
#include <random>

namespace ohao {

inline std::shared_ptr<Model> generateBlobMesh(uint32_t targetTris, uint32_t seed, float radius = 1.0f, float amp = 0.18f) {
    int stacks = std::max(4, int(std::sqrt(double(targetTris) / 4.0)));
    int sectors = stacks * 2;
    auto m = generateSphereMesh(sectors, stacks, 1.0f);
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.0f, 6.2831853f);
    float ph[6]; for (float& p : ph) p = U(rng);
    auto disp = [&](float x, float y, float z) {
        return 1.0f + amp * (0.5f * std::sin(3.0f * x + ph[0]) * std::sin(2.0f * y + ph[1]) + 0.3f * std::sin(5.0f * z + ph[2]) * std::sin(4.0f * x + ph[3]) + 0.2f * std::sin(9.0f * y + ph[4]) * std::sin(7.0f * z + ph[5]));
    };
    for (Vertex& v : m->vertices) {
        float x = v.normal[0], y = v.normal[1], z = v.normal[2];
        float r = radius * disp(x, y, z);
        v.position[0] = x * r; v.position[1] = y * r; v.position[2] = z * r;
    }
    // smooth normals: area-weighted face normals accumulated per vertex
    std::vector<vec3> acc(m->vertices.size());
    for (size_t t = 0; t + 2 < m->indices.size(); t += 3) {
        auto P = [&](uint32_t i) { const Vertex& v = m->vertices[i]; return vec3{v.position[0], v.position[1], v.position[2]}; };
        uint32_t a = m->indices[t], b = m->indices[t + 1], c = m->indices[t + 2];
        vec3 n = cross(P(b) - P(a), P(c) - P(a));
        acc[a] = acc[a] + n; acc[b] = acc[b] + n; acc[c] = acc[c] + n;
    }
    for (size_t i = 0; i < acc.size(); i++) {
        vec3 n = acc[i]; float l = std::sqrt(dot(n, n));
        vec3 radial{m->vertices[i].normal[0], m->vertices[i].normal[1], m->vertices[i].normal[2]};
        if (l > 1e-12f) { n = n * (1.0f / l); if (dot(n, radial) < 0.0f) n = n * -1.0f; } else n = radial;
        m->vertices[i].normal[0] = n.x; m->vertices[i].normal[1] = n.y; m->vertices[i].normal[2] = n.z;
    }
    return m;
}

inline Image8 proceduralTexture(uint32_t size, uint32_t seed, int kind) {   // kind 0 albedo, 1 normal, 2 rough/metal, 3 emissive
    Image8 im; im.w = im.h = size; im.rgba.resize(size_t(size) * size * 4);
    std::mt19937 rng(seed); std::uniform_real_distribution<float> U(0.0f, 6.2831853f);
    float p0 = U(rng), p1 = U(rng), p2 = U(rng);
    for (uint32_t y = 0; y < size; y++) for (uint32_t x = 0; x < size; x++) {
        float u = float(x) / float(size), v = float(y) / float(size);
        auto n = [&](float fx, float fy, float p) { return 0.5f + 0.5f * std::sin(6.2831853f * (fx * u + fy * v) + p); };
        float r, g, b;
        if (kind == 0) { r = 0.35f + 0.5f * n(3, 2, p0); g = 0.30f + 0.45f * n(2, 5, p1); b = 0.25f + 0.4f * n(7, 1, p2); }
        else if (kind == 1) { r = 0.5f + 0.12f * (n(11, 3, p0) - 0.5f); g = 0.5f + 0.12f * (n(4, 13, p1) - 0.5f); b = 1.0f; }
        else if (kind == 2) { r = 1.0f; g = 0.25f + 0.6f * n(5, 4, p0); b = n(2, 2, p1) > 0.7f ? 1.0f : 0.0f; }
        else { float e = n(9, 9, p0) > 0.97f ? 1.0f : 0.0f; r = e; g = 0.6f * e; b = 0.2f * e; }
        uint8_t* px = &im.rgba[(size_t(y) * size + x) * 4];
        px[0] = uint8_t(std::min(std::max(r, 0.0f), 1.0f) * 255.0f + 0.5f); px[1] = uint8_t(std::min(std::max(g, 0.0f), 1.0f) * 255.0f + 0.5f);
        px[2] = uint8_t(std::min(std::max(b, 0.0f), 1.0f) * 255.0f + 0.5f); px[3] = 255;
    }
    return im;
}

// Outdoor-HDRI stand-in: sky gradient + small very bright sun + dark ground (RGBA32F, row 0 = +Y for the CDF convention).
inline std::vector<float> proceduralEnv(uint32_t w, uint32_t h, vec3 sunDir = {0.35f, 0.75f, 0.55f}, float sunRadiance = 4000.0f) {
    std::vector<float> e(size_t(w) * h * 4);
    sunDir = normalize(sunDir);
    const float pi = 3.14159265358979f;
    for (uint32_t y = 0; y < h; y++) for (uint32_t x = 0; x < w; x++) {
        float u = (float(x) + 0.5f) / float(w), v = (float(y) + 0.5f) / float(h);
        float phi = (u - 0.5f) * 2.0f * pi, theta = v * pi;
        vec3 d{std::sin(theta) * std::cos(phi), std::cos(theta), std::sin(theta) * std::sin(phi)};
        float t = std::max(d.y, 0.0f);
        vec3 c = d.y >= 0.0f ? vec3{0.35f + 0.25f * (1 - t), 0.55f + 0.2f * (1 - t), 0.95f} * (0.6f + 0.8f * (1 - t)) : vec3{0.12f, 0.10f, 0.08f};
        float cs = dot(d, sunDir);
        if (cs > 0.9995f) c = c + vec3{1.0f, 0.93f, 0.82f} * sunRadiance;
        else if (cs > 0.99f) c = c + vec3{1.0f, 0.9f, 0.75f} * (8.0f * (cs - 0.99f) / 0.0095f);
        float* px = &e[(size_t(y) * w + x) * 4]; px[0] = c.x; px[1] = c.y; px[2] = c.z; px[3] = 1.0f;
    }
    return e;
}

inline void addBox(Model& m, vec3 lo, vec3 hi, vec3 color) {
    vec3 a{lo.x, lo.y, lo.z}, b{hi.x, lo.y, lo.z}, c{hi.x, hi.y, lo.z}, d{lo.x, hi.y, lo.z}, e{lo.x, lo.y, hi.z}, f{hi.x, lo.y, hi.z}, g{hi.x, hi.y, hi.z}, h{lo.x, hi.y, hi.z};
    addQuad(m, b, a, d, c, {0, 0, -1}, color); addQuad(m, e, f, g, h, {0, 0, 1}, color);
    addQuad(m, a, e, h, d, {-1, 0, 0}, color); addQuad(m, f, b, c, g, {1, 0, 0}, color);
    addQuad(m, d, h, g, c, {0, 1, 0}, color); addQuad(m, a, b, f, e, {0, -1, 0}, color);
}

}  // namespace ohao

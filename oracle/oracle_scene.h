// oracle/oracle_scene.h — TEST INFRASTRUCTURE ONLY (see oracle.cpp header).
// Scene container, samplers, env-map CDF/sampling, texture fetch, BVH + watertight
// ray/triangle test of the CPU oracle.
#pragma once
#include "../include/ohao_b200.h"
#include "oracle_math.h"
#include <vector>
#include <memory>
#include <cstdio>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Samplers.  Follows shaders/includes/rt/sampler_sobol.glsl:13-82 (Owen-scrambled padded-4D
// Sobol), sampler_sobol_tables.glsl:10-47 (direction numbers, data in sobol_dirs.inc),
// sampler_pcg.glsl:9-29 and the
// dispatch in sampler_api.glsl:21-41.
// ---------------------------------------------------------------------------------------------
struct SobolTable {
    // 4 dims x 32 bits.  The reference's constants are not reproducible from the Joe-Kuo
    // recurrences (tools/gen_sobol_table.py explains), so they are carried as generated data.
    uint32_t dirs[4][32];
};
const SobolTable& sobolTable();

static inline uint32_t reverseBits(uint32_t v) {
    v = (v << 16) | (v >> 16);
    v = ((v & 0x00FF00FFu) << 8) | ((v & 0xFF00FF00u) >> 8);
    v = ((v & 0x0F0F0F0Fu) << 4) | ((v & 0xF0F0F0F0u) >> 4);
    v = ((v & 0x33333333u) << 2) | ((v & 0xCCCCCCCCu) >> 2);
    v = ((v & 0x55555555u) << 1) | ((v & 0xAAAAAAAAu) >> 1);
    return v;
}
// Burley 2020 hash-based Owen scramble (sampler_sobol.glsl:15-38 / owen_scramble.cpp:12-31):
// Laine-Karras style hash applied in bit-reversed space.
static inline uint32_t owenScramble(uint32_t v, uint32_t seed) {
    v = reverseBits(v);
    v ^= v * 0x3d20adeau;
    v += seed;
    v *= (seed >> 16) | 1u;
    v ^= v * 0x05526c56u;
    v ^= v * 0x53a22864u;
    return reverseBits(v);
}
static inline uint32_t sobolInt(uint32_t index, uint32_t dim) {
    const uint32_t* d = sobolTable().dirs[dim];
    uint32_t r = 0;
    for (uint32_t bit = 0; index != 0u; bit++, index >>= 1)
        if (index & 1u) r ^= d[bit];
    return r;
}
static inline uint32_t hashPixel(uint32_t px, uint32_t py) {  // sampler_sobol.glsl:53-61 (murmur3 finaliser)
    uint32_t h = (px * 0x1b873593u) ^ (py * 0xcc9e2d51u);
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

struct Sampler {
    uint32_t type;       // OHB_SAMPLER_*
    uint32_t index, pixelSeed;  // Sobol state
    uint32_t pcg;               // PCG state
    void init(uint32_t samplerType, uint32_t px, uint32_t py, uint32_t sampleIdx) {
        type = samplerType;
        index = sampleIdx;
        pixelSeed = hashPixel(px, py);
        pcg = px * 1973u + py * 9277u + sampleIdx * 26699u + 1u;
    }
    uint32_t pcgNext() {
        pcg = pcg * 747796405u + 2891336453u;
        uint32_t w = ((pcg >> ((pcg >> 28u) + 4u)) ^ pcg) * 277803737u;
        return (w >> 22u) ^ w;
    }
    float get1D(uint32_t dim) {
        if (type == OHB_SAMPLER_PCG) return float(pcgNext()) / 4294967296.0f;
        uint32_t pad = dim >> 2, local = dim & 3u;
        uint32_t seed = pixelSeed ^ (pad * 0x9e3779b9u);
        uint32_t s = owenScramble(sobolInt(index, local), seed);
        return float(s >> 8) * (1.0f / 16777216.0f);
    }
    V2 get2D(uint32_t dim) { float x = get1D(dim); float y = get1D(dim + 1u); return {x, y}; }
};

// ---------------------------------------------------------------------------------------------
// Scene
// ---------------------------------------------------------------------------------------------
struct GPULight {  // gpu_light.hpp:10-17, 80 B
    V4 positionAndType, colorAndIntensity, dirAndParam, extra, extra2;
};
static_assert(sizeof(GPULight) == 80, "GPULight layout");

struct Instance {
    uint32_t firstTri, triCount, mask;
    float x[12];      // row-major 3x4 object->world
    M3 normalMat;     // transpose(inverse(mat3(objectToWorld)))
    float inv[12];    // row-major 3x4 world->object
};

struct BvhNode {      // oracle-private binned-SAH BVH2
    V3 lo, hi;
    int32_t left, right;   // children, or leaf when count>0
    uint32_t first, count;
};

struct Scene {
    // geometry (object space, as uploaded)
    std::vector<V3> pos;
    std::vector<uint32_t> idx;
    std::vector<V4> nrm;
    std::vector<V2> uv;
    std::vector<uint32_t> matId;
    std::vector<Instance> inst;
    std::vector<uint32_t> triInst;   // per triangle: owning instance, 0xFFFFFFFF if none
    // world-space triangles (3 verts each), only for triangles covered by an instance
    std::vector<V3> wtri;
    std::vector<uint32_t> activeTris;
    // materials / textures
    std::vector<V4> matColors;
    std::vector<uint8_t> tex; uint32_t texW = 0, texH = 0, texLayers = 0;
    // lights
    uint32_t lightCount = 0, envMapTexIdx = 0xFFFFFFFFu; float envIntensity = 1.0f;
    std::vector<GPULight> lights;
    // env
    std::vector<float> env; uint32_t envW = 0, envH = 0;
    std::vector<float> marg, cond; float envIntegral = 0.0f;
    // bvh
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> bvhTris;   // leaf order -> global triangle id
    bool hasEnv() const { return envMapTexIdx != 0xFFFFFFFFu && !env.empty(); }
    // two-level mode (mirrors ohb_set_accel_mode(OHB_ACCEL_TWO_LEVEL)): one object-space sub-scene per visible instance;
    // sub-scene triangle ids are local (0 .. count-1), global id = firstTri + local
    bool twoLevel = false;
    struct Blas { uint32_t inst, firstTri; std::unique_ptr<Scene> sub; };
    std::vector<Blas> blas;
};
void buildTwoLevel(Scene& s);

void buildEnvCDF(const float* rgba, int W, int H, std::vector<float>& marg, std::vector<float>& cond, float& integral);
void buildBvh(Scene& s);

// ---------------------------------------------------------------------------------------------
// Ray / triangle.  The arithmetic below is the SPEC both the oracle and the CUDA kernels follow
// (DESIGN.md "Intersection arithmetic"): Woop-Benthin-Wald watertight test, fp32, a fixed operation
// order (explicit IEEE fma in the shear and in T, two products + one subtraction in the edge
// functions, nothing contracted by the compiler), fp64 fallback for zero edge functions, two-sided,
// accept tmin < t < tmax, ties on t resolved toward the lower global triangle id.
// The traversal itself lives in the Vulkan driver for the reference (traceRayEXT,
// pt_raygen_offline.rgen:198); there is no reference source to follow for it.
// ---------------------------------------------------------------------------------------------
struct RayPrep {
    V3 o, d; float tmin;
    int kx, ky, kz; float Sx, Sy, Sz;
    V3 idir;  // for slab test only (zero components replaced)
};
static inline RayPrep prepRay(V3 o, V3 d, float tmin) {
    RayPrep r; r.o = o; r.d = d; r.tmin = tmin;
    float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    int kz = 0; float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; m = az; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    if (get(d, kz) < 0.0f) std::swap(kx, ky);
    r.kx = kx; r.ky = ky; r.kz = kz;
    float dz = get(d, kz);
    r.Sx = get(d, kx) / dz; r.Sy = get(d, ky) / dz; r.Sz = 1.0f / dz;
    auto safe = [](float v) { return std::fabs(v) > 1e-20f ? v : (std::signbit(v) ? -1e-20f : 1e-20f); };
    r.idir = {1.0f / safe(d.x), 1.0f / safe(d.y), 1.0f / safe(d.z)};
    return r;
}
// returns true and fills t,u,v if the triangle is hit inside (tmin, tmax)
static inline bool intersectTri(const RayPrep& r, V3 p0, V3 p1, V3 p2, float tmax, float& t, float& bu, float& bv) {
    V3 A = p0 - r.o, B = p1 - r.o, C = p2 - r.o;
    float Akz = get(A, r.kz), Bkz = get(B, r.kz), Ckz = get(C, r.kz);
    // shear: ONE rounding per coordinate (IEEE fma); a per-vertex function, so shared vertices stay shared
    float Ax = std::fmaf(-r.Sx, Akz, get(A, r.kx)), Ay = std::fmaf(-r.Sy, Akz, get(A, r.ky));
    float Bx = std::fmaf(-r.Sx, Bkz, get(B, r.kx)), By = std::fmaf(-r.Sy, Bkz, get(B, r.ky));
    float Cx = std::fmaf(-r.Sx, Ckz, get(C, r.kx)), Cy = std::fmaf(-r.Sy, Ckz, get(C, r.ky));
    // edge functions: two rounded products and one subtraction, NEVER an fma — rn(a*b) - rn(c*d) changes sign exactly
    // when the shared edge is seen from the neighbouring triangle, which is what makes the test watertight
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = float(double(Cx) * double(By) - double(Cy) * double(Bx));
        V = float(double(Ax) * double(Cy) - double(Ay) * double(Cx));
        W = float(double(Bx) * double(Ay) - double(By) * double(Ax));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = (U + V) + W;
    if (det == 0.0f) return false;
    float T = std::fmaf(W, Ckz, std::fmaf(V, Bkz, U * Akz)) * r.Sz;
    float tt = T / det;
    if (!(tt > r.tmin && tt < tmax)) return false;
    t = tt; bu = V / det; bv = W / det;
    return true;
}

ohb_hit traceClosest(const Scene& s, V3 o, V3 d, float tmin, float tmax);
ohb_hit traceClosestBrute(const Scene& s, V3 o, V3 d, float tmin, float tmax);
bool traceAny(const Scene& s, V3 o, V3 d, float tmin, float tmax);

// env sampling (env_sampling.glsl:16-94)
void sampleEnvMap(const Scene& s, float u1, float u2, V3& dir, float& pdf);
float pdfEnvMap(const Scene& s, V3 dir);
// texture fetch: VK_FILTER_LINEAR + REPEAT on R8G8B8A8_UNORM layers / the RGBA32F env image
V4 sampleLayer(const Scene& s, uint32_t layer, V2 uv);
V4 sampleEnvTexture(const Scene& s, V2 uv);

}  // namespace orc

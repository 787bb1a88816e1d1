// oracle/oracle.cpp
//
// ============================  TEST INFRASTRUCTURE ONLY  ====================================
// CPU restatement of the reference's path-tracing hot path (OHAO @ c19e0d4), used ONLY as the
// checker by tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
// bench.py.  Nothing under ohao_engine_b200/ may include, link, import or execute this code.
//
// What is restated (one C++ function per reference shader function, cited at each definition):
//   shaders/rt/pt_raygen_offline.rgen   — whole offline integrator (Stage A/B/C, NEE, env MIS, RR)
//   shaders/rt/pt_closesthit.rchit      — surface fetch / material decode
//   shaders/rt/pt_miss.rmiss            — equirect env lookup + pdfEnvMap
//   shaders/includes/rt/*.glsl          — samplers, env CDF sampling, MIS
//   shaders/includes/material/ggx_aniso.glsl, shaders/rt/includes/pbr_unpack.glsl
//   ohao/render/rt/env_cdf.cpp          — EnvCDF::build
// Third-party arithmetic that is NOT in /root/reference: BVH build + ray/triangle test live in
// the Vulkan driver (vkCmdBuildAccelerationStructuresKHR / traceRayEXT; NVIDIA proprietary on
// the author's machine, Mesa lavapipe on CPU — unpinned).  The oracle uses its own binned-SAH
// BVH and the watertight test specified in oracle_scene.h.
//
// Pinning: tests/test_oracle_ref.py checks samplers and EnvCDF bit-for-bit against
// the reference's own env_cdf.cpp / sobol_generator.cpp / owen_scramble.cpp compiled into
// oracle/_ref (Makefile), the known-answer vectors of tests/renderer/sobol_test.cpp and
// env_cdf_test.cpp, and the offline integrator end-to-end against the reference's golden image
// tests/golden/cornell_box.png under the reference's own tolerance rule (render_golden.py).
// Traversal primitive IDs have no reference pin ("parity unpinned" for that sub-claim only).
// ============================================================================================
#include "oracle_scene.h"
#include <thread>
#include <atomic>
#include <mutex>
#include <string>

namespace orc {

struct Counters { uint64_t samples = 0, closest = 0, shadow = 0, hits = 0; };

struct RayRecorder { ohb_ray* rays = nullptr; ohb_hit* hits = nullptr; uint8_t* kinds = nullptr; uint32_t cap = 0; uint32_t n = 0; };
static RayRecorder g_rec;   // only honoured by single-threaded renders

// Per-frame constants == PTPushConstants (path_tracer.hpp:465-475) as filled by
// PathTracer::render (path_tracer_render.cpp:686-722).
struct Frame {
    M4 invView, invProj, prevViewProj;
    uint32_t W, H, sampleIndex, maxBounces;
    uint32_t flags, historyCount, viewChanged, envW;
    float fireflyClamp, envH, envIntegral, sss;
    float jitX, jitY, aniso, anisoRot;
    uint32_t samplerType, spf;
};

struct Payload {   // RayPayload (pt_closesthit.rchit:8-17)
    V3 color, attenuation, hitPos, hitNormal, hitAlbedo;
    float hitDist; uint32_t hitInstance; float envPdf;
};

struct Tracer {
    const Scene& sc; const Frame& fr; Counters cnt;
    Tracer(const Scene& s, const Frame& f) : sc(s), fr(f) {}

    // pt_miss.rmiss:52-82
    void miss(V3 rayDir, Payload& p) const {
        p.hitDist = -1.0f;
        bool haveEnv = sc.hasEnv();
        if (haveEnv) {
            V3 dir = normalize(rayDir);
            float phi = std::atan2(dir.z, dir.x);
            float theta = std::asin(clampf(dir.y, -1.0f, 1.0f));
            V2 uv{phi / 6.2831853f + 0.5f, theta / 3.1415926f + 0.5f};
            float envS = (sc.envIntensity > 0.0f) ? sc.envIntensity : 1.0f;
            V4 c = sampleEnvTexture(sc, uv);
            p.color = V3{c.x, c.y, c.z} * envS;
        } else {
            p.color = v3(0.0f);
        }
        if (haveEnv && fr.envW > 0u && fr.envH > 0.0f) p.envPdf = pdfEnvMap(sc, normalize(rayDir));
        else p.envPdf = 0.0f;
    }

    // pt_closesthit.rchit:37-162
    void closestHit(V3 o, V3 d, const ohb_hit& h, Payload& p) const {
        p.hitPos = o + d * h.t;
        p.hitDist = h.t;
        const Instance& in = sc.inst[sc.triInst[h.prim]];
        p.hitInstance = in.firstTri;
        float u = h.u, v = h.v, w = 1.0f - u - v;
        uint32_t i0 = sc.idx[size_t(h.prim) * 3], i1 = sc.idx[size_t(h.prim) * 3 + 1], i2 = sc.idx[size_t(h.prim) * 3 + 2];
        V2 t0 = sc.uv[i0], t1 = sc.uv[i1], t2 = sc.uv[i2];
        V2 texUV{w * t0.x + u * t1.x + v * t2.x, w * t0.y + u * t1.y + v * t2.y};
        V3 n0{sc.nrm[i0].x, sc.nrm[i0].y, sc.nrm[i0].z}, n1{sc.nrm[i1].x, sc.nrm[i1].y, sc.nrm[i1].z}, n2{sc.nrm[i2].x, sc.nrm[i2].y, sc.nrm[i2].z};
        V3 interp = w * n0 + u * n1 + v * n2;
        V3 worldN;
        bool thin = !(dot(interp, interp) > 0.0001f);
        if (!thin) {
            worldN = normalize(mulv(in.normalMat, normalize(interp)));
        } else {
            V3 hl{in.inv[0] * p.hitPos.x + in.inv[1] * p.hitPos.y + in.inv[2] * p.hitPos.z + in.inv[3],
                  in.inv[4] * p.hitPos.x + in.inv[5] * p.hitPos.y + in.inv[6] * p.hitPos.z + in.inv[7],
                  in.inv[8] * p.hitPos.x + in.inv[9] * p.hitPos.y + in.inv[10] * p.hitPos.z + in.inv[11]};
            V3 al = vabs(hl), ln;
            if (al.x >= al.y && al.x >= al.z) ln = {signf(hl.x), 0, 0};
            else if (al.y >= al.z) ln = {0, signf(hl.y), 0};
            else ln = {0, 0, signf(hl.z)};
            worldN = normalize(mulv(in.normalMat, ln));
            if (dot(worldN, d) > 0.0f) worldN = -worldN;
        }
        V3 nn0 = normalize(n0), nn1 = normalize(n1), nn2 = normalize(n2);
        float curv = (1.0f - dot(nn0, nn1)) + (1.0f - dot(nn1, nn2)) + (1.0f - dot(nn0, nn2));
        curv = clampf(curv * 8.0f, 0.0f, 1.0f);

        uint32_t matID = sc.matId[h.prim];
        V4 mc = sc.matColors[size_t(matID) * 3], mp = sc.matColors[size_t(matID) * 3 + 1], mp2 = sc.matColors[size_t(matID) * 3 + 2];
        uint32_t diffTex = f2u(mc.w), nrmTex = f2u(mp.z), emTex = f2u(mp.w), rmTex = f2u(mp2.x);
        V3 albedo{mc.x, mc.y, mc.z};
        if (diffTex != OHB_NO_TEXTURE) {
            V4 s = sampleLayer(sc, diffTex, texUV);
            albedo *= vpow(V3{s.x, s.y, s.z}, 2.2f);
        }
        if (nrmTex != OHB_NO_TEXTURE) {
            V4 s = sampleLayer(sc, nrmTex, texUV);
            V3 mapN = normalize(V3{s.x, s.y, s.z} * 2.0f - v3(1.0f));
            V3 T, B;
            if (worldN.z < -0.9999f) { T = {0, -1, 0}; B = {-1, 0, 0}; }
            else {
                float a = 1.0f / (1.0f + worldN.z);
                float dd = -worldN.x * worldN.y * a;
                T = {1.0f - worldN.x * worldN.x * a, dd, -worldN.x};
                B = {dd, 1.0f - worldN.y * worldN.y * a, -worldN.y};
            }
            worldN = normalize(T * mapN.x + B * mapN.y + worldN * mapN.z);
        }
        p.hitNormal = worldN;
        p.hitAlbedo = albedo;
        float rough = mp.x, metal = mp.y;
        if (rmTex != OHB_NO_TEXTURE) { V4 rm = sampleLayer(sc, rmTex, texUV); rough *= rm.y; metal *= rm.z; }
        rough = std::max(rough, 0.04f);
        V3 em = v3(0.0f);
        if (emTex != OHB_NO_TEXTURE) { V4 s = sampleLayer(sc, emTex, texUV); em = vpow(V3{s.x, s.y, s.z}, 2.2f); }
        p.color = em;
        p.attenuation = {rough, clampf(metal, 0.0f, 1.0f), curv};
    }

    void record(V3 o, float tmin, V3 d, float tmax, const ohb_hit& h, uint8_t kind) {
        if (g_rec.rays && g_rec.n < g_rec.cap) {
            g_rec.rays[g_rec.n] = {{o.x, o.y, o.z}, tmin, {d.x, d.y, d.z}, tmax};
            if (g_rec.hits) g_rec.hits[g_rec.n] = h;
            if (g_rec.kinds) g_rec.kinds[g_rec.n] = kind;
            g_rec.n++;
        }
    }
    // traceRayEXT(.., gl_RayFlagsOpaqueEXT, 0xFF, .., origin, 0.001, dir, 10000.0, ..)
    void trace(V3 o, V3 d, Payload& p) {
        cnt.closest++;
        ohb_hit h = traceClosest(sc, o, d, 0.001f, 10000.0f);
        record(o, 0.001f, d, 10000.0f, h, 0);
        if (h.prim == OHB_MISS) miss(d, p);
        else { cnt.hits++; closestHit(o, d, h, p); }
    }
    // TerminateOnFirstHit | SkipClosestHit: visible <=> the miss shader ran (which also refreshes
    // payload.color / envPdf — the env NEE reads payload.color after this call).
    bool shadow(V3 o, V3 d, float tmax, Payload& p) {
        cnt.shadow++;
        bool occ = traceAny(sc, o, d, 0.001f, tmax);
        ohb_hit h{occ ? 1.0f : -1.0f, 0, 0, occ ? 0u : OHB_MISS};
        record(o, 0.001f, d, tmax, h, 1);
        if (occ) { p.hitDist = 999.0f; return false; }
        miss(d, p);
        return true;
    }
};

// pbr_unpack.glsl:8-20
static void unpackHitPbr(V3 att, float& roughness, float& metallic) {
    if (att.x < 0.0f && att.y < 1e-4f) {
        roughness = -att.x; if (roughness >= 10.0f) roughness -= 10.0f;
        roughness = std::max(roughness, 0.01f); metallic = 1.0f;
    } else {
        roughness = std::fabs(att.x); if (roughness >= 10.0f) roughness -= 10.0f;
        roughness = std::max(roughness, 0.01f); metallic = clampf(att.y, 0.0f, 1.0f);
    }
}
// pt_raygen_offline.rgen:93-100
static V3 cosineHemisphere(V3 N, V2 u) {
    V3 up = std::fabs(N.y) < 0.999f ? V3{0, 1, 0} : V3{1, 0, 0};
    V3 T = normalize(cross(up, N));
    V3 B = cross(N, T);
    float r = std::sqrt(u.x);
    float phi = 6.2831853f * u.y;
    return normalize(T * r * std::cos(phi) + B * r * std::sin(phi) + N * std::sqrt(std::max(0.0f, 1.0f - r * r)));
}
static V3 ACES(V3 x) {   // pt_raygen_offline.rgen:103-105
    V3 n = x * (2.51f * x + v3(0.03f)), dd = x * (2.43f * x + v3(0.59f)) + v3(0.14f);
    V3 r = n / dd;
    return {clampf(r.x, 0.0f, 1.0f), clampf(r.y, 0.0f, 1.0f), clampf(r.z, 0.0f, 1.0f)};
}
static float misBalance(float a, float b) { return a / std::max(a + b, 1e-6f); }   // mis.glsl:7-9
// ggx_aniso.glsl:24-58
static void worldUpTangent(V3 n, V3& t, V3& b) {
    V3 up{0, 1, 0};
    V3 ref = std::fabs(dot(up, n)) > 0.97f ? V3{1, 0, 0} : up;
    t = normalize(ref - n * dot(ref, n));
    b = cross(n, t);
}
static float ggxD_anisoOrIso(V3 N, V3 H, float NdotH, float roughness, float anisotropy, float rotation) {
    if (anisotropy < 0.001f) {
        float a = roughness * roughness, a2 = a * a;
        float denom = NdotH * NdotH * (a2 - 1.0f) + 1.0f;
        return a2 / (3.14159265f * denom * denom + 0.0001f);
    }
    V3 T, B; worldUpTangent(N, T, B);
    float c = std::cos(rotation), s = std::sin(rotation);
    V3 Tr = T * c + B * s, Br = B * c - T * s;
    float r2 = roughness * roughness;
    float aspect = std::sqrt(1.0f - anisotropy * 0.9f);
    float rT = std::max(r2 / aspect, 0.001f), rB = std::max(r2 * aspect, 0.001f);
    float TdotH = dot(Tr, H), BdotH = dot(Br, H);
    float d = (TdotH * TdotH / rT) + (BdotH * BdotH / rB) + NdotH * NdotH;
    return 1.0f / (3.14159265f * rT * rB * d * d + 0.0001f);
}
static V3 schlick(V3 F0, float c) { return F0 + (v3(1.0f) - F0) * std::pow(1.0f - c, 5.0f); }

// One NEE light sample: pt_raygen_offline.rgen:311-389 (identical text again at :653-731, :961-1039).
struct LightSample { V3 L, Le; float shadowDist, weight; };
static void sampleSphereLightSolidAngle(V3 p, V3 center, float r, V2 u, V3& L, float& weight, float& shadowDist);   // oracle_realtime.inl
// `solidAngle` selects the realtime raygen's sphere-light sampler (pt_raygen_realtime.rgen:548-552); everything else is shared text.
static LightSample sampleLight(const Scene& sc, Sampler& sm, uint32_t& dimIdx, V3 hitPos, bool solidAngle = false) {
    uint32_t sel = uint32_t(sm.get1D(dimIdx) * float(sc.lightCount)); dimIdx += 1u;
    sel = std::min(sel, sc.lightCount - 1u);
    const GPULight& light = sc.lights[sel];
    float type = light.positionAndType.w;
    V3 center{light.positionAndType.x, light.positionAndType.y, light.positionAndType.z};
    V3 color{light.colorAndIntensity.x, light.colorAndIntensity.y, light.colorAndIntensity.z};
    float intensity = light.colorAndIntensity.w, radius = light.dirAndParam.w;
    V3 lightDir = normalize(V3{light.dirAndParam.x, light.dirAndParam.y, light.dirAndParam.z});
    float outerAngle = light.extra.w;
    float r = std::max(radius, 0.01f);
    float area = 4.0f * 3.14159f * r * r;
    LightSample ls;
    ls.Le = color * intensity / std::max(area, 0.01f);
    auto spherePoint = [&](V3& offset) {
        V2 u = sm.get2D(dimIdx); dimIdx += 2u;
        float cosT = 1.0f - 2.0f * u.x;
        float sinT = std::sqrt(std::max(0.0f, 1.0f - cosT * cosT));
        float phi = 6.2831853f * u.y;
        offset = V3{sinT * std::cos(phi), sinT * std::sin(phi), cosT} * r;
    };
    if (type < 0.5f && solidAngle) {
        V2 u = sm.get2D(dimIdx); dimIdx += 2u;
        sampleSphereLightSolidAngle(hitPos, center, r, u, ls.L, ls.weight, ls.shadowDist);
    } else if (type < 0.5f) {                // sphere: uniform surface point
        V3 offset; spherePoint(offset);
        V3 lp = center + offset, ln = normalize(offset);
        V3 toL = lp - hitPos; float dist = length(toL);
        ls.L = toL / dist;
        float lcos = std::max(dot(-ls.L, ln), 0.0f);
        ls.weight = lcos * area / (dist * dist);
        ls.shadowDist = dist - 0.02f;
    } else if (type < 1.5f) {                // directional
        ls.L = -lightDir; ls.weight = 1.0f; ls.shadowDist = 10000.0f;
    } else if (type < 2.5f) {                // spot
        V3 offset; spherePoint(offset);
        V3 lp = center + offset;
        V3 toL = lp - hitPos; float dist = length(toL);
        ls.L = toL / dist; ls.shadowDist = dist - 0.02f;
        float cosAngle = dot(-ls.L, lightDir);
        const float d2r = 0.017453292519943295f;
        float innerCos = std::cos(light.dirAndParam.w * d2r), outerCos = std::cos(outerAngle * d2r);
        float spot = clampf((cosAngle - outerCos) / (innerCos - outerCos + 0.001f), 0.0f, 1.0f);
        spot *= spot;
        ls.weight = area / (dist * dist) * spot;
    } else {                                 // rect area
        V3 e1{light.extra.x, light.extra.y, light.extra.z}, e2{light.extra2.x, light.extra2.y, light.extra2.z};
        float a = light.extra2.w;
        V2 u = sm.get2D(dimIdx); dimIdx += 2u;
        V3 lp = center + e1 * u.x + e2 * u.y;
        V3 ln = normalize(cross(e1, e2));
        V3 toL = lp - hitPos; float dist = length(toL);
        ls.L = toL / dist;
        float lcos = std::max(dot(-ls.L, ln), 0.0f);
        ls.weight = lcos * a / (dist * dist);
        ls.shadowDist = dist - 0.02f;
    }
    return ls;
}

static float specProbOf(V3 rayDir, V3 N, V3 F0, float rough, float metal) {
    float cosI = std::fabs(dot(normalize(rayDir), N));
    V3 fres = schlick(F0, cosI);
    float sp = maxcomp(fres) * (1.0f - rough * 0.9f);
    return mixf(sp, 1.0f, metal);
}
static void clampLum(V3& c, float cap) { float l = luminance(c); if (l > cap) c *= cap / l; }

struct SampleResult { V3 radiance; V4 albedoAOV, normalAOV; bool wroteAOV; };

struct Integrator {
    Tracer& tr; const Scene& sc; const Frame& fr; Sampler sm; uint32_t dimIdx = 0;
    bool clampOn; bool envOn;
    Integrator(Tracer& t) : tr(t), sc(t.sc), fr(t.fr) {
        clampOn = (fr.flags & OHB_FLAG_ENABLE_FIREFLY_CLAMP) && fr.fireflyClamp > 0.0f;
        envOn = fr.envW > 0u && fr.envH > 0.0f;
    }

    // Env-MIS NEE block: pt_raygen_offline.rgen:506-568 (bounce 0, D via ggxD_anisoOrIso) and
    // :776-831 / :1084-1139 (bounce>=1, inline isotropic D).  `thr` is vec3(1) at bounce 0; the
    // reference omits the multiply there, which is exact for 1.0.
    void envNEE(V3 hitPos, V3 N, V3 inDir, V3 albedo, V3 F0, float rough, float metal, bool bounce0, V3 thr, Payload& pl, V3& radiance) {
        V2 eu = sm.get2D(dimIdx); dimIdx += 2u;
        V3 envDir; float envPdf;
        sampleEnvMap(sc, eu.x, eu.y, envDir, envPdf);
        float NdotL = std::max(dot(N, envDir), 0.0f);
        if (!(NdotL > 0.0f && envPdf > 0.0f)) return;
        pl.hitDist = 999.0f;
        if (!tr.shadow(hitPos + N * 0.01f, envDir, 10000.0f, pl)) return;
        V3 envRad = pl.color;
        V3 V = normalize(-inDir), H = normalize(envDir + V);
        float NdotH = std::max(dot(N, H), 0.001f), NdotV = std::max(dot(N, V), 0.001f), VdotH = std::max(dot(V, H), 0.001f);
        float D;
        if (bounce0) D = ggxD_anisoOrIso(N, H, NdotH, rough, fr.aniso, fr.anisoRot);
        else { float a = rough * rough, a2 = a * a; float dn = NdotH * NdotH * (a2 - 1.0f) + 1.0f; D = a2 / (kPiEnv * dn * dn + 0.0001f); }
        V3 F = schlick(F0, VdotH);
        float k = (rough + 1.0f) * (rough + 1.0f) / 8.0f;
        float G = (NdotL / (NdotL * (1.0f - k) + k)) * (NdotV / (NdotV * (1.0f - k) + k));
        V3 spec = D * F * G / (4.0f * NdotV * NdotL + 0.001f);
        V3 kD = (v3(1.0f) - F) * (1.0f - metal);
        V3 diff = kD * albedo / kPiEnv;
        V3 brdf = diff + spec;
        float specProb = specProbOf(inDir, N, F0, rough, metal);
        float pdfDiff = NdotL / kPiEnv;
        float pdfSpec = D * NdotH / (4.0f * VdotH + 1e-4f);
        float bsdfPdf = mixf(pdfDiff, pdfSpec, specProb);
        float w = misBalance(envPdf, bsdfPdf);
        V3 c = bounce0 ? (envRad * brdf * NdotL * w / envPdf) : (thr * envRad * brdf * NdotL * w / envPdf);
        if (clampOn) clampLum(c, fr.fireflyClamp);
        radiance += c;
    }
    static constexpr float kPiEnv = 3.14159265358979f;   // OHAO_PI

    // NEE at bounce >= 1: pt_raygen_offline.rgen:653-773 / :961-1081
    void lightNEE(V3 hitPos, V3 N, V3 inDir, V3 albedo, V3 F0, float rough, float metal, V3 thr, Payload& pl, V3& radiance) {
        LightSample ls = sampleLight(sc, sm, dimIdx, hitPos);
        float NdotL = std::max(dot(N, ls.L), 0.0f);
        if (!(NdotL > 0.0f && ls.weight > 0.0f)) return;
        pl.hitDist = 999.0f;
        if (!tr.shadow(hitPos + N * 0.01f, ls.L, ls.shadowDist, pl)) return;
        V3 V = normalize(-inDir), H = normalize(ls.L + V);
        float NdotH = std::max(dot(N, H), 0.001f), NdotV = std::max(dot(N, V), 0.001f), VdotH = std::max(dot(V, H), 0.001f);
        float a = rough * rough, a2 = a * a;
        float dn = NdotH * NdotH * (a2 - 1.0f) + 1.0f;
        float D = a2 / (3.14159f * dn * dn + 0.0001f);
        V3 F = schlick(F0, VdotH);
        float k = (rough + 1.0f) * (rough + 1.0f) / 8.0f;
        float G = (NdotL / (NdotL * (1.0f - k) + k)) * (NdotV / (NdotV * (1.0f - k) + k));
        V3 spec = D * F * G / (4.0f * NdotV * NdotL + 0.001f);
        V3 kD = (v3(1.0f) - F) * (1.0f - metal);
        V3 diff = kD * albedo / 3.14159f;
        V3 c = thr * ls.Le * (diff + spec) * NdotL * ls.weight * float(sc.lightCount);
        if (clampOn) clampLum(c, fr.fireflyClamp);
        radiance += c;
    }

    // Bounce-0 analytic direct light incl. the skin/oil-spec extras: pt_raygen_offline.rgen:311-503
    void lightNEE0(V3 hitPos, V3 N, V3 rayDir, V3 albedo, V3 F0, float rough, float metal, float curvature, Payload& pl, V3& radiance) {
        LightSample ls = sampleLight(sc, sm, dimIdx, hitPos);
        float NdotL = std::max(dot(N, ls.L), 0.0f);
        if (!(NdotL > 0.0f && ls.weight > 0.0f)) return;
        pl.hitDist = 999.0f;
        if (!tr.shadow(hitPos + N * 0.01f, ls.L, ls.shadowDist, pl)) return;
        V3 V = normalize(-rayDir), H = normalize(ls.L + V);
        float NdotH = std::max(dot(N, H), 0.001f), NdotV = std::max(dot(N, V), 0.001f), VdotH = std::max(dot(V, H), 0.001f);
        float D = ggxD_anisoOrIso(N, H, NdotH, rough, fr.aniso, fr.anisoRot);
        V3 F = schlick(F0, VdotH);
        float k = (rough + 1.0f) * (rough + 1.0f) / 8.0f;
        float G = (NdotL / (NdotL * (1.0f - k) + k)) * (NdotV / (NdotV * (1.0f - k) + k));
        V3 spec = D * F * G / (4.0f * NdotV * NdotL + 0.001f);
        if (fr.sss > 0.001f && metal < 0.5f) {
            float rS = 0.3f, aS = rS * rS, a2S = aS * aS;
            float dS = NdotH * NdotH * (a2S - 1.0f) + 1.0f;
            float DS = a2S / (3.14159f * dS * dS + 1e-4f);
            float kS = (rS + 1.0f) * (rS + 1.0f) / 8.0f;
            float GS = (NdotL / (NdotL * (1.0f - kS) + kS)) * (NdotV / (NdotV * (1.0f - kS) + kS));
            V3 FS = v3(0.028f) + (v3(1.0f) - v3(0.028f)) * std::pow(1.0f - VdotH, 5.0f);
            V3 oil = DS * GS * FS / (4.0f * NdotV * NdotL + 1e-3f);
            spec += oil * (fr.sss * 0.4f);
        }
        V3 kD = (v3(1.0f) - F) * (1.0f - metal);
        V3 diff = kD * albedo / 3.14159f;
        float skinHint = clampf((albedo.x - albedo.z) * 3.0f, 0.0f, 1.0f);
        float sssStr = fr.sss * skinHint;
        V3 nlDiff = v3(NdotL);
        if (sssStr > 0.001f && metal < 0.5f) {
            float w = NdotL * 0.5f + 0.5f, d = 1.0f - w, d2 = d * d;
            float cs = mixf(1.0f, 0.3f, curvature);
            V3 wrap{std::exp(-d2 * 1.8f * cs), std::exp(-d2 * 6.0f * cs), std::exp(-d2 * 20.0f * cs)};
            V3 tint{1.0f, 0.45f, 0.30f};
            wrap *= mix(v3(1.0f), tint, smoothstep(0.7f, -0.4f, NdotL));
            nlDiff = mix(nlDiff, wrap, sssStr);
        }
        V3 c = ls.Le * (diff * nlDiff + spec * NdotL) * ls.weight * float(sc.lightCount);
        if (clampOn) clampLum(c, fr.fireflyClamp);
        radiance += c;
    }

    // One indirect chain: Stage B (:616-895, specChain=true) or Stage C (:923-1201).
    void chain(bool specChain, V3 origin, V3 dir, V3 thr, float lastPdf, bool lastDelta, V3& radiance) {
        Payload pl{};
        for (uint32_t bounce = 1u; bounce <= fr.maxBounces; bounce++) {
            pl.hitDist = -1.0f;
            tr.trace(origin, dir, pl);
            if (pl.hitDist < 0.0f) {
                float w = 1.0f;
                if (pl.envPdf > 0.0f && fr.envW > 0u && !lastDelta) w = misBalance(lastPdf, pl.envPdf);
                radiance += thr * pl.color * w;
                break;
            }
            V3 hp = pl.hitPos, N = pl.hitNormal, albedo = pl.hitAlbedo, em = pl.color;
            if (length(em) > 0.001f) radiance += thr * em;
            float rough, metal; unpackHitPbr(pl.attenuation, rough, metal);
            V3 F0 = mix(v3(0.04f), albedo, metal);
            if (sc.lightCount > 0u) lightNEE(hp, N, dir, albedo, F0, rough, metal, thr, pl, radiance);
            if (envOn) envNEE(hp, N, dir, albedo, F0, rough, metal, false, thr, pl, radiance);
            if (bounce > 1u) {
                float p = maxcomp(thr);
                float rr = sm.get1D(dimIdx); dimIdx += 1u;
                if (p < 0.01f || rr > p) break;
                thr /= p;
            }
            float specProb = specProbOf(dir, N, F0, rough, metal);
            float choice = sm.get1D(dimIdx); dimIdx += 1u;
            if (choice < specProb || rough < 0.05f) {
                V3 inDir = dir;
                V3 refl = reflect(dir, N);
                if (rough > 0.01f) {
                    V2 ju = sm.get2D(dimIdx); dimIdx += 2u;
                    refl = normalize(refl + cosineHemisphere(refl, ju) * rough);
                    if (dot(refl, N) < 0.0f) { V2 fu = sm.get2D(dimIdx); dimIdx += 2u; refl = cosineHemisphere(N, fu); }
                }
                dir = refl; origin = hp + N * 0.01f;
                // HEAD: Stage B mix(1,albedo,metal) (:864), Stage C albedo*(1-metal) (:1172).  The build that
                // rendered tests/golden/cornell_box.png still used the Stage-B form in Stage C (found by
                // diffing against the golden; DESIGN.md "Golden image"); OHB_FLAG_GOLDEN_COMPAT selects it.
                if (specChain || (fr.flags & OHB_FLAG_GOLDEN_COMPAT)) thr *= mix(v3(1.0f), albedo, metal);
                else           thr *= albedo * (1.0f - metal);
                thr /= std::max(specProb, 0.01f);
                if (rough < 0.05f) { lastPdf = 1.0f; lastDelta = true; }
                else {
                    V3 Hs = normalize(-inDir + dir);
                    float NdotH = std::max(dot(N, Hs), 0.001f), VdotH = std::max(dot(-inDir, Hs), 0.001f);
                    float as = rough * rough, as2 = as * as;
                    float dn = NdotH * NdotH * (as2 - 1.0f) + 1.0f;
                    float Ds = as2 / (kPiEnv * dn * dn + 1e-4f);
                    lastPdf = specProb * (Ds * NdotH / (4.0f * VdotH + 1e-4f));
                    lastDelta = false;
                }
            } else {
                V2 du = sm.get2D(dimIdx); dimIdx += 2u;
                dir = cosineHemisphere(N, du); origin = hp + N * 0.01f;
                thr *= albedo;
                thr /= std::max(1.0f - specProb, 0.01f);
                lastPdf = (1.0f - specProb) * std::max(dot(dir, N), 0.0f) / kPiEnv;
                lastDelta = false;
            }
        }
    }

    // pt_raygen_offline.rgen:129-1209 for one (pixel, sampleIdx)
    SampleResult sample(uint32_t px, uint32_t py, uint32_t sampleIdx) {
        tr.cnt.samples++;
        sm.init(fr.samplerType, px, py, sampleIdx);
        dimIdx = 0u;
        V2 j = sm.get2D(dimIdx); dimIdx += 2u;
        V2 uv{(float(px) + 0.5f + (j.x - 0.5f) + fr.jitX) / float(fr.W), (float(py) + 0.5f + (j.y - 0.5f) + fr.jitY) / float(fr.H)};
        V2 ndc{uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f};
        const float* iv = fr.invView.m;
        V3 camPos{iv[12], iv[13], iv[14]}, fwd{-iv[8], -iv[9], -iv[10]}, right{iv[0], iv[1], iv[2]}, up{iv[4], iv[5], iv[6]};
        float aspect = float(fr.W) / float(fr.H);
        float tanY = std::fabs(fr.invProj.m[5]), tanX = tanY * aspect;
        V3 rayDir = normalize(fwd + right * ndc.x * tanX - up * ndc.y * tanY);
        SampleResult out{}; out.radiance = v3(0.0f); out.wroteAOV = (fr.flags & OHB_FLAG_ENABLE_AOVS) != 0;
        Payload pl{}; pl.hitDist = -1.0f;
        tr.trace(camPos, rayDir, pl);
        if (pl.hitDist < 0.0f) {
            out.radiance = pl.color;
            out.albedoAOV = {pl.color.x, pl.color.y, pl.color.z, 1.0f}; out.normalAOV = {0, 0, 0, 0};
        } else {
            V3 hp = pl.hitPos, N = pl.hitNormal, albedo = pl.hitAlbedo, em = pl.color;
            out.albedoAOV = {albedo.x, albedo.y, albedo.z, 1.0f};
            out.normalAOV = {N.x * 0.5f + 0.5f, N.y * 0.5f + 0.5f, N.z * 0.5f + 0.5f, 1.0f};
            float rough, metal; unpackHitPbr(pl.attenuation, rough, metal);
            V3 F0 = mix(v3(0.04f), albedo, metal);
            float curvature = clampf(pl.attenuation.z, 0.0f, 1.0f);
            V3& rad = out.radiance;
            if (length(em) > 0.001f) rad += em;
            if (sc.lightCount > 0u) lightNEE0(hp, N, rayDir, albedo, F0, rough, metal, curvature, pl, rad);
            if (envOn) envNEE(hp, N, rayDir, albedo, F0, rough, metal, true, v3(1.0f), pl, rad);
            // Stage B set-up (:573-614)
            {
                V3 refl = reflect(rayDir, N);
                if (rough > 0.01f) {
                    V2 ju = sm.get2D(dimIdx); dimIdx += 2u;
                    refl = normalize(refl + cosineHemisphere(refl, ju) * rough);
                    if (dot(refl, N) < 0.0f) { V2 fu = sm.get2D(dimIdx); dimIdx += 2u; refl = cosineHemisphere(N, fu); }
                }
                V3 thr = mix(v3(1.0f), albedo, metal);
                float lastPdf; bool lastDelta;
                if (rough < 0.05f) { lastPdf = 1.0f; lastDelta = true; }
                else {
                    V3 Hs = normalize(-rayDir + refl);
                    float NdotH = std::max(dot(N, Hs), 0.001f), VdotH = std::max(dot(-rayDir, Hs), 0.001f);
                    float Ds = ggxD_anisoOrIso(N, Hs, NdotH, rough, fr.aniso, fr.anisoRot);
                    lastPdf = Ds * NdotH / (4.0f * VdotH + 1e-4f); lastDelta = false;
                }
                chain(true, hp + N * 0.01f, refl, thr, lastPdf, lastDelta, rad);
            }
            // Stage C set-up (:900-921)
            {
                V2 du = sm.get2D(dimIdx); dimIdx += 2u;
                V3 dd = cosineHemisphere(N, du);
                float lastPdf = std::max(dot(dd, N), 0.0f) / kPiEnv;
                chain(false, hp + N * 0.01f, dd, albedo, lastPdf, false, rad);
            }
        }
        if (clampOn) clampLum(out.radiance, fr.fireflyClamp);
        return out;
    }
};

static M4 toM4(const float* p) { M4 m; std::memcpy(m.m, p, 64); return m; }

static void tonemapStore(V3 acc, uint8_t* px) {   // pt_raygen_offline.rgen:1339-1340,1354
    V3 ldr = ACES(acc * 0.5f);
    ldr = vpow(ldr, 1.0f / 2.2f);
    auto q = [](float v) { return uint8_t(std::lrintf(clampf(v, 0.0f, 1.0f) * 255.0f)); };
    px[0] = q(ldr.x); px[1] = q(ldr.y); px[2] = q(ldr.z); px[3] = 255;
}

#include "oracle_realtime.inl"
#include "oracle_svgf.inl"
// rows [0, H) over nthreads host threads (dynamic)
template <class F> static void parallelRows(uint32_t H, int nthreads, F f) {
    std::atomic<uint32_t> next{0};
    auto work = [&]() { for (;;) { uint32_t y = next.fetch_add(1); if (y >= H) break; f(y); } };
    std::vector<std::thread> th; for (int t = 1; t < std::max(nthreads, 1); t++) th.emplace_back(work);
    work(); for (auto& t : th) t.join();
}
#include "oracle_hybrid.inl"

}  // namespace orc

using namespace orc;

// =============================================================================================
// C API (ctypes from tests/ and bench.py)
// =============================================================================================
extern "C" {

struct orc_scene_desc {
    const void* positions; uint64_t stride_bytes; uint32_t nverts;
    const uint32_t* indices; uint32_t ntris;
    const float* normals; const float* uvs; const uint32_t* mat_ids;
    const ohb_instance* instances; uint32_t ninstances;
    const float* mat_colors; uint32_t nmaterials;
    const uint8_t* textures; uint32_t tex_w, tex_h, tex_layers;
    const void* light_ssbo; uint64_t light_bytes;
    const float* env; uint32_t env_w, env_h;
};

void* orc_scene_create(const orc_scene_desc* d) {
    Scene* s = new Scene();
    s->pos.resize(d->nverts);
    for (uint32_t i = 0; i < d->nverts; i++) std::memcpy(&s->pos[i], (const char*)d->positions + size_t(i) * d->stride_bytes, 12);
    s->idx.assign(d->indices, d->indices + size_t(d->ntris) * 3);
    s->nrm.resize(d->nverts); std::memcpy(s->nrm.data(), d->normals, size_t(d->nverts) * 16);
    s->uv.resize(d->nverts); std::memcpy(s->uv.data(), d->uvs, size_t(d->nverts) * 8);
    s->matId.assign(d->mat_ids, d->mat_ids + d->ntris);
    s->matColors.resize(size_t(d->nmaterials) * 3); std::memcpy(s->matColors.data(), d->mat_colors, size_t(d->nmaterials) * 48);
    if (d->textures && d->tex_layers) {
        s->texW = d->tex_w; s->texH = d->tex_h; s->texLayers = d->tex_layers;
        s->tex.assign(d->textures, d->textures + size_t(d->tex_w) * d->tex_h * 4u * d->tex_layers);
    }
    if (d->light_ssbo && d->light_bytes >= 16) {
        const uint8_t* lb = (const uint8_t*)d->light_ssbo;
        std::memcpy(&s->lightCount, lb, 4); std::memcpy(&s->envMapTexIdx, lb + 4, 4); std::memcpy(&s->envIntensity, lb + 8, 4);
        uint32_t avail = uint32_t((d->light_bytes - 16) / 80);
        s->lightCount = std::min(s->lightCount, avail);
        s->lights.resize(s->lightCount);
        std::memcpy(s->lights.data(), lb + 16, size_t(s->lightCount) * 80);
    }
    if (d->env && d->env_w && d->env_h) {
        s->envW = d->env_w; s->envH = d->env_h;
        s->env.assign(d->env, d->env + size_t(d->env_w) * d->env_h * 4u);
        buildEnvCDF(s->env.data(), int(s->envW), int(s->envH), s->marg, s->cond, s->envIntegral);
    }
    // instances -> world-space triangles.  Arithmetic spec (shared with the device):
    // w = ((m0*x + m1*y) + m2*z) + m3, one rounding per op.
    s->triInst.assign(d->ntris, 0xFFFFFFFFu);
    s->wtri.assign(size_t(d->ntris) * 3, V3{0, 0, 0});
    for (uint32_t i = 0; i < d->ninstances; i++) {
        const ohb_instance& oi = d->instances[i];
        Instance in{}; in.firstTri = oi.first_tri; in.triCount = oi.tri_count; in.mask = oi.mask;
        std::memcpy(in.x, oi.xform, 48);
        M3 a; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a.r[r][c] = in.x[r * 4 + c];
        in.normalMat = inverse_transpose(a);
        // world->object = [A^-1 | -A^-1 t];  A^-1 = transpose(normalMat)
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) in.inv[r * 4 + c] = in.normalMat.r[c][r];
            in.inv[r * 4 + 3] = -(in.inv[r * 4 + 0] * in.x[3] + in.inv[r * 4 + 1] * in.x[7] + in.inv[r * 4 + 2] * in.x[11]);
        }
        s->inst.push_back(in);
        for (uint32_t t = oi.first_tri; t < oi.first_tri + oi.tri_count && t < d->ntris; t++) {
            s->triInst[t] = i;
            for (int k = 0; k < 3; k++) {
                V3 p = s->pos[s->idx[size_t(t) * 3 + k]];
                const float* m = in.x;
                s->wtri[size_t(t) * 3 + k] = {((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3],
                                              ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
                                              ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]};
            }
            s->activeTris.push_back(t);
        }
    }
    buildBvh(*s);
    return s;
}
void orc_scene_destroy(void* h) { delete (Scene*)h; }
// ohb_set_accel_mode: 0 flatten (world-space triangles), 1 two-level (object-space BLAS per instance, ray mapped on descent)
void orc_scene_set_accel_mode(void* h, int mode) {
    Scene* s = (Scene*)h; s->twoLevel = mode != 0;
    if (s->twoLevel && s->blas.empty()) buildTwoLevel(*s);
}
// ohb_update_instances: new transforms for the same instances
void orc_scene_update_instances(void* h, const ohb_instance* insts, uint32_t n) {
    Scene* s = (Scene*)h;
    for (uint32_t i = 0; i < n && i < s->inst.size(); i++) {
        Instance& in = s->inst[i];
        std::memcpy(in.x, insts[i].xform, 48);
        M3 a; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a.r[r][c] = in.x[r * 4 + c];
        in.normalMat = inverse_transpose(a);
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) in.inv[r * 4 + c] = in.normalMat.r[c][r];
            in.inv[r * 4 + 3] = -(in.inv[r * 4 + 0] * in.x[3] + in.inv[r * 4 + 1] * in.x[7] + in.inv[r * 4 + 2] * in.x[11]);
        }
        for (uint32_t t = in.firstTri; t < in.firstTri + in.triCount && size_t(t) * 3 + 2 < s->wtri.size(); t++)
            for (int k = 0; k < 3; k++) {
                V3 p = s->pos[s->idx[size_t(t) * 3 + k]];
                const float* m = in.x;
                s->wtri[size_t(t) * 3 + k] = {((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7], ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]};
            }
    }
    buildBvh(*s);
}

// Cheap edits mirroring updateRTMaterialParams / updateRTLightParams (no BVH rebuild).
void orc_scene_set_materials(void* h, const float* mat_colors, uint32_t nmat) {
    Scene* s = (Scene*)h; s->matColors.resize(size_t(nmat) * 3); std::memcpy(s->matColors.data(), mat_colors, size_t(nmat) * 48);
}
void orc_scene_set_lights(void* h, const void* ssbo, uint64_t bytes) {
    Scene* s = (Scene*)h; const uint8_t* lb = (const uint8_t*)ssbo;
    std::memcpy(&s->lightCount, lb, 4); std::memcpy(&s->envMapTexIdx, lb + 4, 4); std::memcpy(&s->envIntensity, lb + 8, 4);
    s->lightCount = std::min(s->lightCount, uint32_t((bytes - 16) / 80));
    s->lights.resize(s->lightCount); std::memcpy(s->lights.data(), lb + 16, size_t(s->lightCount) * 80);
}

void orc_env_cdf(const float* rgba, uint32_t w, uint32_t h, float* marg, float* cond, float* integral) {
    std::vector<float> m, c; float I;
    buildEnvCDF(rgba, int(w), int(h), m, c, I);
    std::memcpy(marg, m.data(), m.size() * 4); std::memcpy(cond, c.data(), c.size() * 4); *integral = I;
}
void orc_scene_env_cdf(void* h, float* marg, float* cond, float* integral) {
    Scene* s = (Scene*)h;
    std::memcpy(marg, s->marg.data(), s->marg.size() * 4); std::memcpy(cond, s->cond.data(), s->cond.size() * 4); *integral = s->envIntegral;
}
void orc_env_sample_batch(void* h, const float* u12, uint32_t n, float* dir_pdf, float* pdf_of_dir) {
    Scene* s = (Scene*)h;
    for (uint32_t i = 0; i < n; i++) {
        V3 d; float p; sampleEnvMap(*s, u12[2 * i], u12[2 * i + 1], d, p);
        dir_pdf[4 * i] = d.x; dir_pdf[4 * i + 1] = d.y; dir_pdf[4 * i + 2] = d.z; dir_pdf[4 * i + 3] = p;
        if (pdf_of_dir) pdf_of_dir[i] = pdfEnvMap(*s, d);
    }
}
// NRD front-end packing: shaders/includes/rt/nrd_frontend.glsl:11-41 and pt_raygen_offline.rgen:106-127, line by line.
static V3 nrdLinearToYCoCg(V3 c) { return {dot(c, V3{0.25f, 0.5f, 0.25f}), dot(c, V3{0.5f, 0.0f, -0.5f}), dot(c, V3{-0.25f, 0.5f, -0.25f})}; }     // :11-16
static V3 nrdYCoCgToLinear(V3 c) {                                                                                                              // :18-25
    float t = c.x - c.z; V3 r; r.y = c.x + c.z; r.x = t + c.y; r.z = t - c.y;
    return {std::max(r.x, 0.0f), std::max(r.y, 0.0f), std::max(r.z, 0.0f)};
}
static float nrdNormHitDist(float hitDist, float viewZ, float roughness) {                                                                      // :28-34
    float smc = roughness * roughness; smc = std::min(std::max(smc, 0.0f), 1.0f);
    float f = (3.0f + std::fabs(viewZ) * 0.1f) * (20.0f * (1.0f - smc) + 1.0f * smc);
    return std::min(std::max(hitDist / std::max(f, 1e-6f), 0.0f), 1.0f);
}
void orc_nrd_pack_batch(const float* in6, const float* nr4, uint32_t n, float* packedRad, float* packedNormal, float* unpackedRgb) {
    for (uint32_t i = 0; i < n; i++) {
        const float* a = in6 + size_t(i) * 6;
        V3 rad{std::max(a[0], 0.0f), std::max(a[1], 0.0f), std::max(a[2], 0.0f)};                                                                // :37-41
        V3 y = nrdLinearToYCoCg(rad);
        float hd = nrdNormHitDist(std::max(a[3], 0.0f), a[4], a[5]);
        packedRad[4 * i] = y.x; packedRad[4 * i + 1] = y.y; packedRad[4 * i + 2] = y.z; packedRad[4 * i + 3] = hd;
        V3 nn{nr4[4 * i], nr4[4 * i + 1], nr4[4 * i + 2]}; float rough = nr4[4 * i + 3];                                                        // rgen:111-127
        float l1 = std::fabs(nn.x) + std::fabs(nn.y) + std::fabs(nn.z);
        nn = {nn.x / l1, nn.y / l1, nn.z / l1};
        float ry = nn.y * 0.5f + 0.5f, rx = nn.x * 0.5f + ry; ry -= nn.x * 0.5f;
        rough = std::max(rough, 1.5f / 512.0f);
        float sgn = (nn.z < 0.0f) ? -rough : rough;
        packedNormal[4 * i] = rx; packedNormal[4 * i + 1] = ry; packedNormal[4 * i + 2] = sgn * 0.5f + 0.5f; packedNormal[4 * i + 3] = 0.0f;
        V3 back = nrdYCoCgToLinear(y);
        unpackedRgb[3 * i] = back.x; unpackedRgb[3 * i + 1] = back.y; unpackedRgb[3 * i + 2] = back.z;
    }
}
void orc_hybrid_shadow(void* h, uint32_t W, uint32_t H, const float* gPos, const float* gNrm, const orc_hybrid_shadow_params* p, uint8_t* mask, int nthreads) {
    hybridShadow(*(Scene*)h, W, H, gPos, gNrm, *p, mask, nthreads);
}
void orc_hybrid_gi(void* h, uint32_t W, uint32_t H, const float* gPos, const float* gNrm, const float* gAlbedo, const float* history, const float* instMat,
                   const orc_hybrid_gi_params* p, uint16_t* out, int nthreads) {
    hybridGi(*(Scene*)h, W, H, gPos, gNrm, gAlbedo, history, instMat, *p, out, nthreads);
}
void orc_env_pdf_batch(void* h, const float* dirs3, uint32_t n, float* pdf) {
    Scene* s = (Scene*)h;
    for (uint32_t i = 0; i < n; i++) pdf[i] = pdfEnvMap(*s, V3{dirs3[3 * i], dirs3[3 * i + 1], dirs3[3 * i + 2]});
}
float orc_sampler_1d(uint32_t sampler_type, uint32_t px, uint32_t py, uint32_t sample_idx, uint32_t dim) {
    Sampler sm; sm.init(sampler_type, px, py, sample_idx); return sm.get1D(dim);
}
float orc_sobol_raw(uint32_t index, uint32_t dim) { return float(sobolInt(index, dim) >> 8) * (1.0f / 16777216.0f); }
uint32_t orc_owen(uint32_t v, uint32_t seed) { return owenScramble(v, seed); }
const uint32_t* orc_sobol_dirs(void) { return &sobolTable().dirs[0][0]; }

void orc_trace_batch(void* h, const ohb_ray* rays, uint32_t n, ohb_hit* hits, int brute, int nthreads) {
    Scene* s = (Scene*)h;
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint32_t> next{0};
    auto work = [&]() {
        for (;;) {
            uint32_t b = next.fetch_add(1024); if (b >= n) break;
            for (uint32_t i = b; i < std::min(n, b + 1024u); i++) {
                V3 o{rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]}, d{rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]};
                hits[i] = brute ? traceClosestBrute(*s, o, d, rays[i].tmin, rays[i].tmax) : traceClosest(*s, o, d, rays[i].tmin, rays[i].tmax);
            }
        }
    };
    std::vector<std::thread> th; for (int t = 1; t < nthreads; t++) th.emplace_back(work);
    work(); for (auto& t : th) t.join();
}
void orc_occluded_batch(void* h, const ohb_ray* rays, uint32_t n, uint8_t* occ) {
    Scene* s = (Scene*)h;
    for (uint32_t i = 0; i < n; i++) {
        V3 o{rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]}, d{rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]};
        occ[i] = traceAny(*s, o, d, rays[i].tmin, rays[i].tmax) ? 1 : 0;
    }
}
// Record every ray the next SINGLE-THREADED render traces (kind 0 = closest, 1 = shadow).
void orc_set_ray_recorder(ohb_ray* rays, ohb_hit* hits, uint8_t* kinds, uint32_t cap) { g_rec = {rays, hits, kinds, cap, 0}; }
uint32_t orc_ray_recorder_count(void) { return g_rec.n; }

struct orc_render_args {
    float view[16], proj[16];
    uint32_t width, height;
    uint32_t first_sample_index;   /* PathTracer::m_sampleIndex at the first frame             */
    uint32_t history_count;        /* PathTracer::m_historyFrameCount at the first frame       */
    uint32_t nsamples;             /* number of consecutive render() calls to emulate          */
    uint32_t tile_x, tile_y, tile_w, tile_h;   /* 0,0,0,0 = full frame                          */
    ohb_settings settings;
    float* accum;                  /* W*H*4 RGBA32F, in/out (running mean, count in .w)        */
    uint8_t* ldr;                  /* W*H*4 or NULL                                            */
    float* albedo; float* normal;  /* W*H*4 or NULL                                            */
    float* sample_dump;            /* nsamples*W*H*4 or NULL                                   */
    int32_t nthreads;
    ohb_counters counters;         /* out                                                      */
};

int orc_render_offline(void* h, orc_render_args* a) {
    Scene* s = (Scene*)h;
    Frame fr{};
    M4 view = toM4(a->view), proj = toM4(a->proj);
    fr.invView = inverse(view); fr.invProj = inverse(proj); fr.prevViewProj = mulm(proj, view);
    fr.W = a->width; fr.H = a->height; fr.maxBounces = a->settings.max_bounces & 0xFFFFu;
    fr.flags = a->settings.flags; fr.viewChanged = 0;
    fr.envW = s->hasEnv() ? s->envW : 0u; fr.envH = s->hasEnv() ? float(s->envH) : 0.0f; fr.envIntegral = s->envIntegral;
    fr.fireflyClamp = a->settings.firefly_clamp_lum; fr.sss = a->settings.subsurface_strength;
    fr.jitX = fr.jitY = 0.0f; fr.aniso = a->settings.anisotropy_strength; fr.anisoRot = a->settings.anisotropy_rotation;
    fr.samplerType = a->settings.sampler_type; fr.spf = 1;
    uint32_t x0 = a->tile_x, y0 = a->tile_y, tw = a->tile_w ? a->tile_w : a->width, th = a->tile_h ? a->tile_h : a->height;
    int nthreads = std::max(1, a->nthreads);
    if (g_rec.rays) nthreads = 1;
    std::atomic<uint32_t> nextRow{0};
    std::mutex mu; Counters total;
    auto work = [&]() {
        Tracer tr(*s, fr);
        Integrator integ(tr);
        for (;;) {
            uint32_t row = nextRow.fetch_add(1); if (row >= th) break;
            uint32_t py = y0 + row;
            for (uint32_t px = x0; px < x0 + tw; px++) {
                size_t pi = size_t(py) * a->width + px;
                float* acc = a->accum + pi * 4;
                for (uint32_t k = 0; k < a->nsamples; k++) {
                    SampleResult r = integ.sample(px, py, a->first_sample_index + k);
                    // accumulate: pt_raygen_offline.rgen:1212-1262 (own-pixel running mean)
                    uint32_t hist = a->history_count + k;
                    if (hist == 0u) { acc[0] = r.radiance.x; acc[1] = r.radiance.y; acc[2] = r.radiance.z; acc[3] = 1.0f; }
                    else {
                        float cnt = acc[3] + 1.0f;
                        acc[0] = (acc[0] * acc[3] + r.radiance.x) / cnt;
                        acc[1] = (acc[1] * acc[3] + r.radiance.y) / cnt;
                        acc[2] = (acc[2] * acc[3] + r.radiance.z) / cnt;
                        acc[3] = cnt;
                    }
                    if (a->sample_dump) {
                        float* sd = a->sample_dump + (size_t(k) * a->width * a->height + pi) * 4;
                        sd[0] = r.radiance.x; sd[1] = r.radiance.y; sd[2] = r.radiance.z; sd[3] = 1.0f;
                    }
                    if (r.wroteAOV) {
                        if (a->albedo) std::memcpy(a->albedo + pi * 4, &r.albedoAOV, 16);
                        if (a->normal) std::memcpy(a->normal + pi * 4, &r.normalAOV, 16);
                    }
                }
                if (a->ldr) tonemapStore(V3{acc[0], acc[1], acc[2]}, a->ldr + pi * 4);
            }
        }
        std::lock_guard<std::mutex> lk(mu);
        total.samples += tr.cnt.samples; total.closest += tr.cnt.closest; total.shadow += tr.cnt.shadow; total.hits += tr.cnt.hits;
    };
    std::vector<std::thread> thv; for (int t = 1; t < nthreads; t++) thv.emplace_back(work);
    work(); for (auto& t : thv) t.join();
    a->counters = ohb_counters{};
    a->counters.samples = total.samples; a->counters.closest_rays = total.closest; a->counters.shadow_rays = total.shadow; a->counters.closest_hits = total.hits;
    return 0;
}

// One realtime frame (pt_raygen_realtime.rgen): pass 1 = everything up to the accumBuffer store, pass 2 = a-trous + tonemap.
struct orc_rt_args {
    float view[16], proj[16], prev_view_proj[16];
    uint32_t width, height, frame_index, history_count, view_changed;
    ohb_settings settings;
    const float* accum_prev; float* accum_curr; const float* surf_prev; float* surf_curr; const float* shad_prev; float* shad_curr;
    const float* res_prev[3]; float* res_curr[3];
    float* albedo; float* normal; float* radiance_dump; float* gi_dump; float* denoised; uint8_t* ldr;
    int32_t nthreads; ohb_counters counters;
};
int orc_render_realtime(void* h, orc_rt_args* a) {
    Scene* s = (Scene*)h;
    Frame fr{};
    M4 view = toM4(a->view), proj = toM4(a->proj);
    fr.invView = inverse(view); fr.invProj = inverse(proj); fr.prevViewProj = toM4(a->prev_view_proj);
    fr.W = a->width; fr.H = a->height; fr.sampleIndex = a->frame_index; fr.maxBounces = a->settings.max_bounces & 0xFFFFu;
    fr.flags = a->settings.flags; fr.historyCount = a->history_count; fr.viewChanged = a->view_changed;
    fr.envW = s->hasEnv() ? s->envW : 0u; fr.envH = s->hasEnv() ? float(s->envH) : 0.0f; fr.envIntegral = s->envIntegral;
    fr.fireflyClamp = a->settings.firefly_clamp_lum; fr.sss = a->settings.subsurface_strength;
    fr.jitX = fr.jitY = 0.0f; fr.aniso = a->settings.anisotropy_strength; fr.anisoRot = a->settings.anisotropy_rotation;
    fr.samplerType = a->settings.sampler_type;
    uint32_t spf = a->settings.samples_per_frame; fr.spf = spf < 1u ? 1u : (spf > 64u ? 64u : spf);
    RTImages im{a->accum_prev, a->accum_curr, a->surf_prev, a->surf_curr, a->shad_prev, a->shad_curr,
                {a->res_prev[0], a->res_prev[1], a->res_prev[2]}, {a->res_curr[0], a->res_curr[1], a->res_curr[2]},
                a->albedo, a->normal, a->radiance_dump, a->gi_dump};
    int nthreads = std::max(1, a->nthreads);
    std::atomic<uint32_t> nextRow{0};
    std::mutex mu; Counters total;
    auto work = [&]() {
        Tracer tr(*s, fr);
        RealtimeIntegrator integ(tr);
        for (;;) {
            uint32_t row = nextRow.fetch_add(1); if (row >= a->height) break;
            for (uint32_t px = 0; px < a->width; px++) integ.pixel(px, row, im);
        }
        std::lock_guard<std::mutex> lk(mu);
        total.samples += tr.cnt.samples; total.closest += tr.cnt.closest; total.shadow += tr.cnt.shadow; total.hits += tr.cnt.hits;
    };
    { std::vector<std::thread> thv; for (int t = 1; t < nthreads; t++) thv.emplace_back(work); work(); for (auto& t : thv) t.join(); }
    nextRow = 0;
    auto work2 = [&]() {
        for (;;) {
            uint32_t row = nextRow.fetch_add(1); if (row >= a->height) break;
            for (uint32_t px = 0; px < a->width; px++) {
                size_t pi = size_t(row) * a->width + px;
                realtimeDenoisePixel(fr, a->accum_curr, a->normal, int(px), int(row), a->ldr ? a->ldr + pi * 4 : nullptr, a->denoised ? a->denoised + pi * 4 : nullptr);
            }
        }
    };
    { std::vector<std::thread> thv; for (int t = 1; t < nthreads; t++) thv.emplace_back(work2); work2(); for (auto& t : thv) t.join(); }
    a->counters = ohb_counters{};
    a->counters.samples = total.samples; a->counters.closest_rays = total.closest; a->counters.shadow_rays = total.shadow; a->counters.closest_hits = total.hits;
    return 0;
}

// Guide AOVs of the SVGF denoiser from this frame's surface-history plane (firstHitPos, firstHitDist).
// currViewProj = inverse(invProj) * inverse(invView) and viewMat = inverse(invView), as the shader derives them.
void orc_svgf_guides(const float* surf, uint32_t W, uint32_t H, const float* view16, const float* proj16, const float* prevViewProj16, uint32_t frameIdx,
                     uint32_t* motion, float* depth) {
    M4 invView = inverse(toM4(view16)), invProj = inverse(toM4(proj16));
    M4 viewMat = inverse(invView), currViewProj = mulm(inverse(invProj), viewMat), prevVP = toM4(prevViewProj16);
    for (uint32_t y = 0; y < H; y++) for (uint32_t x = 0; x < W; x++) {
        size_t pi = size_t(y) * W + x;
        svgfGuidesPixel(currViewProj, prevVP, viewMat, W, H, frameIdx, V4{surf[pi * 4], surf[pi * 4 + 1], surf[pi * 4 + 2], surf[pi * 4 + 3]}, motion[pi], depth[pi]);
    }
}
struct orc_svgf_args {
    int32_t width, height, reset, nthreads;
    float sigma_l, sigma_normal, sigma_depth, _pad;
    uint8_t* beauty; const uint32_t* motion; const float* depth; const float* normal;
    const uint16_t* prev_color; const uint16_t* prev_moments; const uint16_t* prev_geom;
    uint16_t* cur_color; uint16_t* cur_moments; uint16_t* cur_geom;
};
int orc_svgf_dispatch(orc_svgf_args* a) {
    SvgfImages im{a->width, a->height, a->beauty, a->motion, a->depth, a->normal, a->prev_color, a->prev_moments, a->prev_geom, a->cur_color, a->cur_moments, a->cur_geom};
    svgfDispatch(im, a->reset, a->sigma_l, a->sigma_normal, a->sigma_depth, a->nthreads);
    return 0;
}
uint16_t orc_f2h(float f) { return halfBits(f); }
float orc_h2f(uint16_t h) { return halfToFloat(h); }

void orc_tonemap(const float* accum, uint32_t npix, uint8_t* ldr) {
    for (uint32_t i = 0; i < npix; i++) tonemapStore(V3{accum[4 * i], accum[4 * i + 1], accum[4 * i + 2]}, ldr + 4 * size_t(i));
}

}  // extern "C"

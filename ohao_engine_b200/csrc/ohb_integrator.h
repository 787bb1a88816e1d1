// ohb_integrator.h — wavefront form of the OFFLINE integrator (per-path code).
//
// The reference runs one megakernel thread per pixel (shaders/rt/pt_raygen_offline.rgen:129-1355)
// calling traceRayEXT up to 27 times.  B200 has no RT cores, so the same path tree is unrolled into
// a state machine: every wavefront iteration traces ONE closest-hit ray per live path
// (k_trace_closest), shades it (shadePath below: closest-hit/miss shader + the raygen's per-bounce
// body), emits at most two visibility rays whose contributions are parked in `pend` until the
// any-hit kernel has answered, and writes the next ray.  Shading is ONE kernel per iteration (k_shade): the
// closest-hit / miss SHADERS (geometry + texture gathers -> payload) and the raygen's per-bounce body (NEE, MIS, lobe
// sampling, state machine) back to back with the 64-B payload in registers, over a queue that k_sort_hits has split into
// hit-only and miss-only warps.  (The realtime profile keeps k_surface + k_bounce_rt with the payload in memory.)
// Stage B (specular chain) and Stage C
// (diffuse chain) of one sample run back-to-back on the same path slot because the Sobol dimension
// counter is shared: C's first dimension depends on how many B consumed.  Contributions are added
// in exactly the reference's order, so a sample's radiance differs from the megakernel only by
// libm-vs-CUDA intrinsic rounding.
#pragma once
#include "ohb_traverse.h"

namespace ohb {

enum { ST_PRIMARY = 0u, ST_CHAIN_B = 1u, ST_CHAIN_C = 2u, ST_DONE = 3u };
#define OHB_ST_STAGE(s)   ((s) & 3u)
#define OHB_ST_BOUNCE(s)  (((s) >> 2) & 15u)
#define OHB_ST_DELTA      (1u << 6)
#define OHB_ST_PEND_A     (1u << 7)
#define OHB_ST_PEND_B     (1u << 8)
#define OHB_ST_MAKE(stage, bounce) ((stage) | ((bounce) << 2))

// Queue entries carry the path index (OHB_MAX_PATHS <= 2^28) plus what the NEXT shading kernel needs to know to
// issue all of its path-state loads at once instead of chasing meta -> pend -> payload (k_bounce was stalled on
// that dependent chain: 60 % of its samples on the first uses of meta / pendA / pay0, profile r1d).
#define OHB_Q_PATH(e)   ((e) & 0x0FFFFFFFu)
#define OHB_Q_MISS      0x80000000u      // set by k_surface: the last closest-hit query missed
#define OHB_Q_PEND_A    0x40000000u      // set by k_bounce: a light-NEE contribution is parked in pendA
#define OHB_Q_PEND_B    0x20000000u      //                  an env-NEE contribution is parked in pendB
#define OHB_Q_PRIMARY   0x10000000u      // set by k_raygen: the path is at its camera ray
#define OHB_Q_NONE      0xFFFFFFFFu      // bouncePath: the path traces no further closest-hit ray

struct PathArrays {
    f4* rayO; f4* rayD;            // next closest-hit ray (origin.xyz | dir.xyz)
    ohb_hit* hit;                  // answer of k_trace_closest
    f4* thr;                       // throughput.xyz, lastBsdfPdf
    f4* rad;                       // radiance.xyz of this sample so far
    f4* pendA; f4* pendB;          // parked light-NEE / env-NEE contributions (zeroed by k_trace_shadow when occluded)
    u4* meta;                      // x = px | py << 16, y = sampleIdx, z = dimIdx, w = state bits
    f4* fh0; f4* fh1; f4* fh2;            // first hit kept for the Stage C set-up: (pos, rough) (N, metal) (albedo, -)
    f4* pay0; f4* pay1; f4* pay2; f4* pay3;   // (realtime profile only; the offline k_shade keeps it in registers) RayPayload of the last query: (hitPos|color, hitDist) (N, att.x | envPdf) (albedo, att.y) (emission, att.z)
    f4* shO; f4* shD;              // visibility-ray queue: (origin, tmax) (dir, bits(path << 1 | slot))
    uint32_t* shCount;
    uint32_t* queueIn; uint32_t* queueOut; uint32_t* countIn; uint32_t* countOut;
    uint8_t* hitFlag;              // per slot of queueIn: 1 = the closest-hit query hit (k_trace_closest -> k_sort_hits of the fused shading path)
    uint32_t* queueSorted; uint32_t* sortCount;   // k_surface re-emits queueIn two-ended: hits from the front, misses from the back; sortCount[0..1]
    unsigned long long* counters;  // [0] samples [1] closest rays [2] shadow rays [3] closest hits
    f4* albedoAOV; f4* normalAOV;  // W*H images (first hit of the last sample rendered)
    uint32_t numPixels;            // pixels in the tile
    uint32_t samplesInBatch;
    uint32_t firstSampleIndex;     // sample index of s = 0
    const u4* sobolTab;            // samplesInBatch entries: sobolQuad(firstSampleIndex + s)
    // STABLE octant binning of a ray queue before it is traced (k_oct_hist / scan / k_oct_scatter, OHB_OCT_BIN=1): octPerm lists
    // the queue slots grouped by ray octant, queue order kept inside each group — a warp then holds rays of one octant from
    // neighbouring paths, which descend the children of a node in the same order.  null = trace in queue order.
    uint32_t* octPerm; uint32_t* octHist; uint32_t* octScanTmp; uint32_t octBlocks;
};

struct Payload { f3 color, attenuation, hitPos, hitNormal, hitAlbedo; float hitDist, envPdf; };

// pt_miss.rmiss:52-82: dirToEquirect + bilinear fetch x envIntensity (dir normalised).  One shared copy: the miss shader
// and env NEE both call it (two inlined copies in the fused k_shade).
OHB_SHARED_FN f4 envRadianceSharedPhi(const f4* env, uint32_t envW, uint32_t envH, float envIntensity, float phi, float dy) {
    float theta = asinf(clampf(dy, -1.0f, 1.0f));
    f2 uv; uv.x = ediv(phi, 6.2831853f) + 0.5f; uv.y = ediv(theta, 3.1415926f) + 0.5f;
    float envS = (envIntensity > 0.0f) ? envIntensity : 1.0f;
    f3 c = xyz(sampleEnvTexel4(env, int(envW), int(envH), uv)) * envS;
    return mk4(c, 0.0f);
}
OHB_HD f3 envRadiance(const SceneDev& sc, f3 dir) { return xyz(envRadianceSharedPhi(sc.env, sc.envW, sc.envH, sc.envIntensity, ohb_atan2(dir.z, dir.x), dir.y)); }
OHB_HD void missShader(const SceneDev& sc, const FrameParams& fr, f3 rayDir, Payload& p) {
    p.hitDist = -1.0f;
    bool haveEnv = sc.envMapTexIdx != 0xFFFFFFFFu && sc.env != nullptr;
    if (haveEnv) {
        f3 dir = normalize(rayDir);
        const float phi = ohb_atan2(dir.z, dir.x);         // one atan2 for the lookup (dirToEquirect) and the pdf (pdfEnvMap)
        p.color = xyz(envRadianceSharedPhi(sc.env, sc.envW, sc.envH, sc.envIntensity, phi, dir.y));
        p.envPdf = (fr.envW > 0u && fr.envH > 0.0f) ? pdfEnvMapPhi(sc, dir.y, phi) : 0.0f;
    } else {
        p.color = mk3(0.0f); p.envPdf = 0.0f;
    }
}

// pt_closesthit.rchit:37-162 — 136 B of surface gathers + texels per hit.
OHB_HD void closestHitShader(const SceneDev& sc, f3 o, f3 d, const ohb_hit& h, Payload& p) {
    p.hitPos = o + d * h.t;
    p.hitDist = h.t;
    float u = h.u, v = h.v, w = 1.0f - u - v;
    const uint32_t* ip = sc.indices + size_t(h.prim) * 3u;
    uint32_t i0 = ldg(ip), i1 = ldg(ip + 1), i2 = ldg(ip + 2);
    f2 t0 = ldg(sc.uvs + i0), t1 = ldg(sc.uvs + i1), t2 = ldg(sc.uvs + i2);
    f2 texUV; texUV.x = w * t0.x + u * t1.x + v * t2.x; texUV.y = w * t0.y + u * t1.y + v * t2.y;
    f3 n0 = xyz(ld4(sc.normals + i0)), n1 = xyz(ld4(sc.normals + i1)), n2 = xyz(ld4(sc.normals + i2));
    f3 interp = w * n0 + u * n1 + v * n2;
    uint32_t inst = ldg(sc.triInst + h.prim);
    const f4* nm = sc.instNormalMat + size_t(inst) * 3u;
    f4 m0 = ld4(nm), m1 = ld4(nm + 1), m2 = ld4(nm + 2);
    f3 worldN;
    if (dot(interp, interp) > 0.0001f) {
        f3 ni = normalize(interp);
        worldN = normalize(mk3(dot(xyz(m0), ni), dot(xyz(m1), ni), dot(xyz(m2), ni)));
    } else {
        const f4* iv = sc.instInv + size_t(inst) * 3u;
        f4 r0 = ld4(iv), r1 = ld4(iv + 1), r2 = ld4(iv + 2);
        f3 hl = mk3(dot(xyz(r0), p.hitPos) + r0.w, dot(xyz(r1), p.hitPos) + r1.w, dot(xyz(r2), p.hitPos) + r2.w);
        f3 al = vabs(hl), ln;
        if (al.x >= al.y && al.x >= al.z) ln = mk3(signf(hl.x), 0, 0);
        else if (al.y >= al.z) ln = mk3(0, signf(hl.y), 0);
        else ln = mk3(0, 0, signf(hl.z));
        worldN = normalize(mk3(dot(xyz(m0), ln), dot(xyz(m1), ln), dot(xyz(m2), ln)));
        if (dot(worldN, d) > 0.0f) worldN = -worldN;
    }
    f3 nn0 = normalize(n0), nn1 = normalize(n1), nn2 = normalize(n2);
    float curv = (1.0f - dot(nn0, nn1)) + (1.0f - dot(nn1, nn2)) + (1.0f - dot(nn0, nn2));
    curv = clampf(curv * 8.0f, 0.0f, 1.0f);

    uint32_t matID = ldg(sc.matIds + h.prim);
    const f4* mp = sc.matColors + size_t(matID) * 3u;
    f4 mc = ld4(mp), mpar = ld4(mp + 1), mpar2 = ld4(mp + 2);
    uint32_t diffTex = f2u(mc.w), nrmTex = f2u(mpar.z), emTex = f2u(mpar.w), rmTex = f2u(mpar2.x);
    f3 albedo = xyz(mc);
    if (diffTex != OHB_NO_TEXTURE) albedo *= vpow22(xyz(sampleLayer(sc, diffTex, texUV)));
    if (nrmTex != OHB_NO_TEXTURE) {
        f3 mapN = normalize(xyz(sampleLayer(sc, nrmTex, texUV)) * 2.0f - mk3(1.0f));
        f3 T, B;
        if (worldN.z < -0.9999f) { T = mk3(0, -1, 0); B = mk3(-1, 0, 0); }
        else {
            float a = 1.0f / (1.0f + worldN.z);
            float dd = -worldN.x * worldN.y * a;
            T = mk3(1.0f - worldN.x * worldN.x * a, dd, -worldN.x);
            B = mk3(dd, 1.0f - worldN.y * worldN.y * a, -worldN.y);
        }
        worldN = normalize(T * mapN.x + B * mapN.y + worldN * mapN.z);
    }
    p.hitNormal = worldN; p.hitAlbedo = albedo;
    float rough = mpar.x, metal = mpar.y;
    if (rmTex != OHB_NO_TEXTURE) { f4 rm = sampleLayer(sc, rmTex, texUV); rough *= rm.y; metal *= rm.z; }
    rough = fmaxf(rough, 0.04f);
    f3 em = mk3(0.0f);
    if (emTex != OHB_NO_TEXTURE) em = vpow22(xyz(sampleLayer(sc, emTex, texUV)));
    p.color = em;
    p.attenuation = mk3(rough, clampf(metal, 0.0f, 1.0f), curv);
}

OHB_HD void unpackHitPbr(f3 att, float& roughness, float& metallic) {   // pbr_unpack.glsl:8-20
    if (att.x < 0.0f && att.y < 1e-4f) {
        roughness = -att.x; if (roughness >= 10.0f) roughness -= 10.0f;
        roughness = fmaxf(roughness, 0.01f); metallic = 1.0f;
    } else {
        roughness = fabsf(att.x); if (roughness >= 10.0f) roughness -= 10.0f;
        roughness = fmaxf(roughness, 0.01f); metallic = clampf(att.y, 0.0f, 1.0f);
    }
}
OHB_HD f3 cosineHemisphere(f3 N, f2 u) {   // pt_raygen_offline.rgen:93-100
    f3 up = fabsf(N.y) < 0.999f ? mk3(0, 1, 0) : mk3(1, 0, 0);
    f3 T = normalize(cross(up, N));
    f3 B = cross(N, T);
    float r = sqrtf(u.x);
    float sp, cp; ohb_sincos_turns(u.y, sp, cp);                     // phi = 6.2831853 * u.y
    return normalize(T * r * cp + B * r * sp + N * sqrtf(fmaxf(0.0f, 1.0f - r * r)));
}
#if defined(__CUDACC__)
static __device__ __host__ __noinline__ f3 cosineHemisphereShared(f3 N, f2 u) { return cosineHemisphere(N, u); }   // 4 call sites, one copy
#else
static inline f3 cosineHemisphereShared(f3 N, f2 u) { return cosineHemisphere(N, u); }
#endif
OHB_HD float misBalance(float a, float b) { return ediv(a, fmaxf(a + b, 1e-6f)); }   // mis.glsl:7-9
// anisotropic branch of ggxD_anisoOrIso: one shared, cold copy (anisotropy is 0 unless RTRenderSettings sets it)
OHB_SHARED_FN float ggxD_aniso(f3 N, f3 H, float NdotH, float roughness, float anisotropy, float rotation) {
    f3 up = mk3(0, 1, 0);
    f3 ref = fabsf(dot(up, N)) > 0.97f ? mk3(1, 0, 0) : up;
    f3 T = normalize(ref - N * dot(ref, N));
    f3 B = cross(N, T);
    float c, s; ohb_sincos(rotation, s, c);
    f3 Tr = T * c + B * s, Br = B * c - T * s;
    float r2 = roughness * roughness;
    float aspect = sqrtf(1.0f - anisotropy * 0.9f);
    float rT = fmaxf(r2 / aspect, 0.001f), rB = fmaxf(r2 * aspect, 0.001f);
    float TdotH = dot(Tr, H), BdotH = dot(Br, H);
    float dd = (TdotH * TdotH / rT) + (BdotH * BdotH / rB) + NdotH * NdotH;
    return 1.0f / (3.14159265f * rT * rB * dd * dd + 0.0001f);
}
OHB_HD float ggxD_anisoOrIso(f3 N, f3 H, float NdotH, float roughness, float anisotropy, float rotation) {   // ggx_aniso.glsl:24-58
    if (anisotropy < 0.001f) {
        float a = roughness * roughness, a2 = a * a;
        float denom = NdotH * NdotH * (a2 - 1.0f) + 1.0f;
        return ediv(a2, 3.14159265f * denom * denom + 0.0001f);
    }
    return ggxD_aniso(N, H, NdotH, roughness, anisotropy, rotation);
}
OHB_HD f3 schlick(f3 F0, float c) { return F0 + (mk3(1.0f) - F0) * pow5(1.0f - c); }   // pow(x, 5.0) as 3 multiplies (<= 2 ulp)
OHB_HD float specProbOf(f3 rayDir, f3 N, f3 F0, float rough, float metal) {
    float cosI = fabsf(dot(normalize(rayDir), N));
    float sp = maxcomp(schlick(F0, cosI)) * (1.0f - rough * 0.9f);
    return mixf(sp, 1.0f, metal);
}
OHB_HD void clampLum(f3& c, float cap) { float l = luminance(c); if (l > cap) c *= cap / l; }

// One NEE light sample: pt_raygen_offline.rgen:311-389 (== :653-731 == :961-1039).
struct LightSample { f3 L, Le; float shadowDist, weight; };
// pt_raygen_realtime.rgen:150-186 — subtended-cone sampling of a sphere light (realtime profile)
OHB_HD void sampleSphereLightSolidAngle(f3 p, f3 center, float r, f2 u, f3& L, float& weight, float& shadowDist) {
    f3 toCenter = center - p;
    float d2 = dot(toCenter, toCenter);
    float d = sqrtf(fmaxf(d2, 1e-8f));
    f3 axis = toCenter / d;
    float cosThetaMax = sqrtf(fmaxf(0.0f, 1.0f - (r * r) / fmaxf(d2, 1e-8f)));
    float cosTheta = 1.0f - u.x * (1.0f - cosThetaMax);
    float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta * cosTheta));
    f3 up = fabsf(axis.y) < 0.999f ? mk3(0, 1, 0) : mk3(1, 0, 0);
    f3 T = normalize(cross(up, axis));
    f3 B = cross(axis, T);
    float sp, cp; ohb_sincos_turns(u.y, sp, cp);                     // phi = 6.2831853 * u.y
    L = normalize(T * (sinTheta * cp) + B * (sinTheta * sp) + axis * cosTheta);
    weight = 6.2831853f * (1.0f - cosThetaMax);
    float b = dot(L, -toCenter);
    float c = d2 - r * r;
    float disc = b * b - c;
    float t;
    if (disc > 0.0f) { float sq = sqrtf(disc); t = -b - sq; if (t < 1e-3f) t = -b + sq; }
    else t = d;
    shadowDist = fmaxf(t - 0.02f, 1e-3f);
}
OHB_HD LightSample sampleLight(const SceneDev& sc, Sampler& sm, uint32_t& dimIdx, f3 hitPos, bool solidAngle = false) {
    uint32_t sel = uint32_t(sm.get1D(dimIdx) * float(sc.lightCount)); dimIdx += 1u;
    sel = sel < sc.lightCount - 1u ? sel : sc.lightCount - 1u;
    const f4* lp4 = reinterpret_cast<const f4*>(sc.lights + sel);
    f4 pt = ld4(lp4), ci = ld4(lp4 + 1), dp = ld4(lp4 + 2), ex = ld4(lp4 + 3), ex2 = ld4(lp4 + 4);
    float type = pt.w;
    f3 center = xyz(pt);
    float radius = dp.w;
    f3 lightDir = normalize(xyz(dp));
    float r = fmaxf(radius, 0.01f);
    float area = 4.0f * 3.14159f * r * r;
    LightSample ls;
    ls.Le = xyz(ci) * ci.w / fmaxf(area, 0.01f);
    if (type < 0.5f && solidAngle) {
        f2 u = sm.get2D(dimIdx); dimIdx += 2u;
        sampleSphereLightSolidAngle(hitPos, center, r, u, ls.L, ls.weight, ls.shadowDist);
    } else if (type < 0.5f || (type >= 1.5f && type < 2.5f)) {
        f2 u = sm.get2D(dimIdx); dimIdx += 2u;
        float cosT = 1.0f - 2.0f * u.x;
        float sinT = sqrtf(fmaxf(0.0f, 1.0f - cosT * cosT));
        float sp, cp; ohb_sincos_turns(u.y, sp, cp);                 // phi = 6.2831853 * u.y
        f3 offset = mk3(sinT * cp, sinT * sp, cosT) * r;
        f3 toL = (center + offset) - hitPos; float dist = length(toL);
        ls.L = toL / dist; ls.shadowDist = dist - 0.02f;
        if (type < 0.5f) {
            float lcos = fmaxf(dot(-ls.L, normalize(offset)), 0.0f);
            ls.weight = ediv(lcos * area, dist * dist);
        } else {
            float cosAngle = dot(-ls.L, lightDir);
            const float d2r = 0.017453292519943295f;
            float innerCos = ohb_cos(dp.w * d2r), outerCos = ohb_cos(ex.w * d2r);
            float spot = clampf((cosAngle - outerCos) / (innerCos - outerCos + 0.001f), 0.0f, 1.0f);
            spot *= spot;
            ls.weight = area / (dist * dist) * spot;
        }
    } else if (type < 1.5f) {
        ls.L = -lightDir; ls.weight = 1.0f; ls.shadowDist = 10000.0f;
    } else {
        f3 e1 = xyz(ex), e2 = xyz(ex2);
        f2 u = sm.get2D(dimIdx); dimIdx += 2u;
        f3 lp = center + e1 * u.x + e2 * u.y;
        f3 ln = normalize(cross(e1, e2));
        f3 toL = lp - hitPos; float dist = length(toL);
        ls.L = toL / dist;
        float lcos = fmaxf(dot(-ls.L, ln), 0.0f);
        ls.weight = ediv(lcos * ex2.w, dist * dist);
        ls.shadowDist = dist - 0.02f;
    }
    return ls;
}

struct ShadeCtx {
    const SceneDev& sc; const FrameParams& fr; const PathArrays& P;
    Sampler sm; uint32_t dimIdx; uint32_t path; uint32_t state;
    bool clampOn, envOn;
    OHB_HD ShadeCtx(const SceneDev& s, const FrameParams& f, const PathArrays& p) : sc(s), fr(f), P(p) {
        clampOn = (fr.flags & OHB_FLAG_ENABLE_FIREFLY_CLAMP) && fr.fireflyClamp > 0.0f;
        envOn = fr.envW > 0u && fr.envH > 0.0f;
    }
    OHB_HD void pushShadow(f3 o, f3 d, float tmax, uint32_t slot, f3 contribution) {
        uint32_t q = alloc_slot(P.shCount);
        P.shO[q] = mk4(o, tmax);
        P.shD[q] = mk4(d, u2f((path << 1) | slot));
        if (slot == 0u) { P.pendA[path] = mk4(contribution, 0.0f); state |= OHB_ST_PEND_A; }
        else            { P.pendB[path] = mk4(contribution, 0.0f); state |= OHB_ST_PEND_B; }
    }

    // light NEE: bounce 0 = pt_raygen_offline.rgen:311-503 (aniso D, skin extras, thr == 1),
    //            bounce >= 1 = :653-773 / :961-1081 (inline isotropic D).
    OHB_HD void lightNEE(f3 hp, f3 N, f3 inDir, f3 albedo, f3 F0, float rough, float metal, float curvature, bool bounce0, f3 thr) {
        LightSample ls = sampleLight(sc, sm, dimIdx, hp);
        float NdotL = fmaxf(dot(N, ls.L), 0.0f);
        if (!(NdotL > 0.0f && ls.weight > 0.0f)) return;
        f3 V = normalize(-inDir), H = normalize(ls.L + V);
        float NdotH = fmaxf(dot(N, H), 0.001f), NdotV = fmaxf(dot(N, V), 0.001f), VdotH = fmaxf(dot(V, H), 0.001f);
        float D;
        if (bounce0) D = ggxD_anisoOrIso(N, H, NdotH, rough, fr.aniso, fr.anisoRot);
        else { float a = rough * rough, a2 = a * a; float dn = NdotH * NdotH * (a2 - 1.0f) + 1.0f; D = ediv(a2, 3.14159f * dn * dn + 0.0001f); }
        f3 F = schlick(F0, VdotH);
        float k = (rough + 1.0f) * (rough + 1.0f) / 8.0f;
        float G = ediv(NdotL, NdotL * (1.0f - k) + k) * ediv(NdotV, NdotV * (1.0f - k) + k);
        f3 spec = D * F * G / (4.0f * NdotV * NdotL + 0.001f);
        f3 c;
        if (bounce0) {
            if (fr.sss > 0.001f && metal < 0.5f) {
                float rS = 0.3f, aS = rS * rS, a2S = aS * aS;
                float dS = NdotH * NdotH * (a2S - 1.0f) + 1.0f;
                float DS = a2S / (3.14159f * dS * dS + 1e-4f);
                float kS = (rS + 1.0f) * (rS + 1.0f) / 8.0f;
                float GS = (NdotL / (NdotL * (1.0f - kS) + kS)) * (NdotV / (NdotV * (1.0f - kS) + kS));
                f3 FS = mk3(0.028f) + (mk3(1.0f) - mk3(0.028f)) * pow5(1.0f - VdotH);
                spec += (DS * GS * FS / (4.0f * NdotV * NdotL + 1e-3f)) * (fr.sss * 0.4f);
            }
            f3 kD = (mk3(1.0f) - F) * (1.0f - metal);
            f3 diff = kD * albedo / 3.14159f;
            float sssStr = fr.sss * clampf((albedo.x - albedo.z) * 3.0f, 0.0f, 1.0f);
            f3 nlDiff = mk3(NdotL);
            if (sssStr > 0.001f && metal < 0.5f) {
                float w = NdotL * 0.5f + 0.5f, dd = 1.0f - w, d2 = dd * dd;
                float cs = mixf(1.0f, 0.3f, curvature);
                f3 wrap = mk3(expf(-d2 * 1.8f * cs), expf(-d2 * 6.0f * cs), expf(-d2 * 20.0f * cs));
                wrap *= mix(mk3(1.0f), mk3(1.0f, 0.45f, 0.30f), smoothstepf(0.7f, -0.4f, NdotL));
                nlDiff = mix(nlDiff, wrap, sssStr);
            }
            c = ls.Le * (diff * nlDiff + spec * NdotL) * ls.weight * float(sc.lightCount);
        } else {
            f3 kD = (mk3(1.0f) - F) * (1.0f - metal);
            f3 diff = kD * albedo / 3.14159f;
            c = thr * ls.Le * (diff + spec) * NdotL * ls.weight * float(sc.lightCount);
        }
        if (clampOn) clampLum(c, fr.fireflyClamp);
        pushShadow(hp + N * 0.01f, ls.L, ls.shadowDist, 0u, c);
    }

    // env NEE + balance-heuristic MIS: :506-568 (bounce 0) and :776-831 / :1084-1139.
    OHB_HD void envNEE(f3 hp, f3 N, f3 inDir, f3 albedo, f3 F0, float rough, float metal, bool bounce0, f3 thr) {
        f2 eu = sm.get2D(dimIdx); dimIdx += 2u;
        f3 envDir; float envPdf;
        sampleEnvMap(sc, eu.x, eu.y, envDir, envPdf);
        float NdotL = fmaxf(dot(N, envDir), 0.0f);
        if (!(NdotL > 0.0f && envPdf > 0.0f)) return;
        // payload.color of the visibility ray's miss shader == env radiance along envDir
        f3 envRad = envRadiance(sc, envDir);
        f3 V = normalize(-inDir), H = normalize(envDir + V);
        float NdotH = fmaxf(dot(N, H), 0.001f), NdotV = fmaxf(dot(N, V), 0.001f), VdotH = fmaxf(dot(V, H), 0.001f);
        float D;
        if (bounce0) D = ggxD_anisoOrIso(N, H, NdotH, rough, fr.aniso, fr.anisoRot);
        else { float a = rough * rough, a2 = a * a; float dn = NdotH * NdotH * (a2 - 1.0f) + 1.0f; D = ediv(a2, OHB_PI * dn * dn + 0.0001f); }
        f3 F = schlick(F0, VdotH);
        float k = (rough + 1.0f) * (rough + 1.0f) / 8.0f;
        float G = ediv(NdotL, NdotL * (1.0f - k) + k) * ediv(NdotV, NdotV * (1.0f - k) + k);
        f3 spec = D * F * G / (4.0f * NdotV * NdotL + 0.001f);
        f3 kD = (mk3(1.0f) - F) * (1.0f - metal);
        f3 diff = kD * albedo / OHB_PI;
        f3 brdf = diff + spec;
        float specProb = specProbOf(inDir, N, F0, rough, metal);
        float bsdfPdf = mixf(ediv(NdotL, OHB_PI), ediv(D * NdotH, 4.0f * VdotH + 1e-4f), specProb);
        float w = misBalance(envPdf, bsdfPdf);
        f3 c = bounce0 ? (envRad * brdf * NdotL * w / envPdf) : (thr * envRad * brdf * NdotL * w / envPdf);
        if (clampOn) clampLum(c, fr.fireflyClamp);
        pushShadow(hp + N * 0.01f, envDir, 10000.0f, 1u, c);
    }
};

// Camera ray: pt_raygen_offline.rgen:141-162.  Writes the primary ray of path p.
OHB_HD void raygenPath(const FrameParams& fr, const PathArrays& P, uint32_t p) {
    uint32_t s = p / P.numPixels, pix = p - s * P.numPixels;
    // 8x4 pixel blocks per warp keep primary rays of a warp coherent
    uint32_t tilesX = (fr.tileW + 7u) / 8u;
    uint32_t blk = pix / 32u, inb = pix & 31u;
    uint32_t lx = (blk % tilesX) * 8u + (inb & 7u), ly = (blk / tilesX) * 4u + (inb >> 3);
    u4 m; m.y = P.firstSampleIndex + s; m.z = 2u;
    if (lx >= fr.tileW || ly >= fr.tileH) {   // padding lane of a partial block
        m.x = 0xFFFFFFFFu; m.w = ST_DONE; P.meta[p] = m; P.rad[p] = mk4(0, 0, 0, 0);
        return;
    }
    uint32_t px = fr.tileX + lx, py = fr.tileY + ly;
    m.x = px | (py << 16);
    Sampler sm; sm.init(fr.samplerType, px, py, m.y, ldu4(P.sobolTab + s));
    f2 j = sm.get2D(0u);
    float uvx = (float(px) + 0.5f + (j.x - 0.5f) + fr.jitX) / float(fr.W), uvy = (float(py) + 0.5f + (j.y - 0.5f) + fr.jitY) / float(fr.H);
    float nx = uvx * 2.0f - 1.0f, ny = uvy * 2.0f - 1.0f;
    f3 dir = normalize(fr.fwd + fr.right * nx * fr.tanX - fr.up * ny * fr.tanY);
    m.w = OHB_ST_MAKE(ST_PRIMARY, 0u);
    if (fr.samplerType == OHB_SAMPLER_PCG) m.z = sm.pcg;   // PCG is stateful: carry the state instead of a dimension
    P.meta[p] = m;
    P.rayO[p] = mk4(fr.camPos, 0.0f); P.rayD[p] = mk4(dir, 0.0f);
    P.rad[p] = mk4(0, 0, 0, 0);
}

// The closest-hit / miss shader of path p's last query -> the four payload quads.  Returns true for a hit.
OHB_HD bool surfaceShade(const SceneDev& sc, const FrameParams& fr, const PathArrays& P, uint32_t p, f4& q0, f4& q1, f4& q2, f4& q3) {
    f3 o = xyz(P.rayO[p]), d = xyz(P.rayD[p]);
    ohb_hit h = P.hit[p];
    Payload pl;
    if (h.prim == OHB_MISS) {
        missShader(sc, fr, d, pl);
        q0 = mk4(pl.color, -1.0f); q1 = mk4(pl.envPdf, 0.0f, 0.0f, 0.0f); q2 = q3 = mk4(0.0f, 0.0f, 0.0f, 0.0f);
        return false;
    }
    closestHitShader(sc, o, d, h, pl);
    q0 = mk4(pl.hitPos, pl.hitDist); q1 = mk4(pl.hitNormal, pl.attenuation.x);
    q2 = mk4(pl.hitAlbedo, pl.attenuation.y); q3 = mk4(pl.color, pl.attenuation.z);
    return true;
}
// k_surface: payload record of path p to memory.
// Returns true for a hit (the caller sorts hits to the front of queueSorted, misses to the back).
OHB_HD bool surfacePath(const SceneDev& sc, const FrameParams& fr, const PathArrays& P, uint32_t p) {
    f4 q0, q1, q2, q3;
    const bool hit = surfaceShade(sc, fr, P, p, q0, q1, q2, q3);
    P.pay0[p] = q0; P.pay1[p] = q1;
    if (hit) { P.pay2[p] = q2; P.pay3[p] = q3; }
    return hit;
}

// reflect + roughness jitter + below-horizon re-sample: :573-592 (B set-up) == :843-862 == :1151-1170
OHB_HD f3 sampleSpecDir(ShadeCtx& cx, f3 d, f3 N, float rough) {
    f3 refl = reflect(d, N);
    if (rough > 0.01f) {
        f2 ju = cx.sm.get2D(cx.dimIdx); cx.dimIdx += 2u;
        refl = normalize(refl + cosineHemisphereShared(refl, ju) * rough);
        if (dot(refl, N) < 0.0f) { f2 fu = cx.sm.get2D(cx.dimIdx); cx.dimIdx += 2u; refl = cosineHemisphereShared(N, fu); }
    }
    return refl;
}

// The raygen's per-bounce body for the path of queue entry e, given the payload (q0..q3) of its last query.
// Returns the path's next queue entry, or OHB_Q_NONE if it traces no further closest-hit ray.
// PRIM: 1 = every entry of this launch is a camera ray (iteration 0 of a wavefront), 0 = none is (later iterations), 2 = read
// the queue flag.  The stage-specialised bodies drop the other stage's code (AOV writes, first-hit record, anisotropic GGX
// vs Russian roulette, lobe selection, chain bookkeeping) from kernels whose size is a first-class cost.
template <int PRIM = 2>
OHB_HD uint32_t bounceBody(const SceneDev& sc, const FrameParams& fr, const PathArrays& P, uint32_t e, f4 q0, f4 q1, f4 q2, f4 q3) {
    const uint32_t p = OHB_Q_PATH(e);
    const bool isMiss = (e & OHB_Q_MISS) != 0u, primary = PRIM == 2 ? (e & OHB_Q_PRIMARY) != 0u : PRIM == 1;
    // every load of the path record is issued here, before the first use
    const u4 m = P.meta[p];
    const f4 rad4 = P.rad[p], d4 = P.rayD[p];
    const f4 zero4 = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    const f4 pa4 = (e & OHB_Q_PEND_A) ? P.pendA[p] : zero4, pb4 = (e & OHB_Q_PEND_B) ? P.pendB[p] : zero4;
    const f4 t4 = primary ? mk4(1.0f, 1.0f, 1.0f, 0.0f) : P.thr[p];
    ShadeCtx cx(sc, fr, P);
    cx.path = p; cx.state = m.w;
    uint32_t px = m.x & 0xFFFFu, py = m.x >> 16;
    cx.sm.init(fr.samplerType, px, py, m.y, ldu4(P.sobolTab + (m.y - P.firstSampleIndex)));
    if (fr.samplerType == OHB_SAMPLER_PCG) { cx.sm.pcg = m.z; cx.dimIdx = 0u; } else cx.dimIdx = m.z;
    f3 rad = xyz(rad4);
    if (e & OHB_Q_PEND_A) rad += xyz(pa4);
    if (e & OHB_Q_PEND_B) rad += xyz(pb4);
    cx.state &= ~(OHB_ST_PEND_A | OHB_ST_PEND_B);
    uint32_t stage = OHB_ST_STAGE(cx.state), bounce = OHB_ST_BOUNCE(cx.state);
    const bool specChain = (stage == ST_CHAIN_B);
    f3 d = xyz(d4);

    f3 nextO = mk3(0.0f), nextD = mk3(0.0f), thr = xyz(t4); float lastPdf = t4.w; bool lastDelta = !primary && (cx.state & OHB_ST_DELTA) != 0u;
    bool startC = false, finished = false, chainEnds = false;
    f3 fhPos = mk3(0.0f), fhN = mk3(0.0f), fhAlbedo = mk3(0.0f);   // first-hit data for the Stage C set-up

    bool aov = false; size_t pi = 0;
    if (primary) {
        bool lastSample = (m.y == P.firstSampleIndex + P.samplesInBatch - 1u);
        aov = (fr.flags & OHB_FLAG_ENABLE_AOVS) && lastSample && P.albedoAOV;
        pi = size_t(py) * fr.W + px;
    }
    if (isMiss) {
        f3 color = xyz(q0); float envPdf = q1.x;
        if (primary) {
            rad = color;
            if (aov) { P.albedoAOV[pi] = mk4(color, 1.0f); P.normalAOV[pi] = mk4(0, 0, 0, 0); }
            finished = true;
        } else {
            float w = 1.0f;
            if (envPdf > 0.0f && fr.envW > 0u && !lastDelta) w = misBalance(lastPdf, envPdf);
            rad += thr * color * w;
            chainEnds = true;
        }
    } else {
        f3 hp = xyz(q0), N = xyz(q1), albedo = xyz(q2), em = xyz(q3);
        if (aov) { P.albedoAOV[pi] = mk4(albedo, 1.0f); P.normalAOV[pi] = mk4(N * 0.5f + mk3(0.5f), 1.0f); }
        float rough, metal; unpackHitPbr(mk3(q1.w, q2.w, q3.w), rough, metal);
        f3 F0 = mix(mk3(0.04f), albedo, metal);
        float curvature = primary ? clampf(q3.w, 0.0f, 1.0f) : 0.0f;
        if (length(em) > 0.001f) rad += primary ? em : thr * em;
        // NEE: one call site each for every bounce (bounce 0 differs inside, see ShadeCtx)
        if (sc.lightCount > 0u) cx.lightNEE(hp, N, d, albedo, F0, rough, metal, curvature, primary, thr);
        if (cx.envOn) cx.envNEE(hp, N, d, albedo, F0, rough, metal, primary, thr);
        bool killed = false; float specProb = 1.0f; bool takeSpec = true;
        if (!primary) {
            if (bounce > 1u) {   // Russian roulette :834-839 / :1142-1147
                float pr = maxcomp(thr);
                float rr = cx.sm.get1D(cx.dimIdx); cx.dimIdx += 1u;
                if (pr < 0.01f || rr > pr) killed = true; else thr /= pr;
            }
            if (!killed) {
                specProb = specProbOf(d, N, F0, rough, metal);
                float choice = cx.sm.get1D(cx.dimIdx); cx.dimIdx += 1u;
                takeSpec = (choice < specProb || rough < 0.05f);
            }
        }
        if (killed) chainEnds = true;
        else {
            nextO = hp + N * 0.01f;
            if (takeSpec) {
                nextD = sampleSpecDir(cx, d, N, rough);
                if (primary) thr = mix(mk3(1.0f), albedo, metal);                                                          // Stage B set-up :594
                else {
                    if (specChain || (fr.flags & OHB_FLAG_GOLDEN_COMPAT)) thr *= mix(mk3(1.0f), albedo, metal);            // :864
                    else                                                   thr *= albedo * (1.0f - metal);                 // :1172
                    thr /= fmaxf(specProb, 0.01f);
                }
                if (rough < 0.05f) { lastPdf = 1.0f; lastDelta = true; }
                else {
                    f3 Hs = normalize(-d + nextD);
                    float NdotH = fmaxf(dot(N, Hs), 0.001f), VdotH = fmaxf(dot(-d, Hs), 0.001f);
                    float Ds;
                    if (primary) Ds = ggxD_anisoOrIso(N, Hs, NdotH, rough, fr.aniso, fr.anisoRot);
                    else { float as = rough * rough, as2 = as * as; float dn = NdotH * NdotH * (as2 - 1.0f) + 1.0f; Ds = ediv(as2, OHB_PI * dn * dn + 1e-4f); }
                    float pdfH = ediv(Ds * NdotH, 4.0f * VdotH + 1e-4f);
                    lastPdf = primary ? pdfH : specProb * pdfH; lastDelta = false;
                }
            } else {
                f2 du = cx.sm.get2D(cx.dimIdx); cx.dimIdx += 2u;
                nextD = cosineHemisphereShared(N, du);
                thr *= albedo;
                thr /= fmaxf(1.0f - specProb, 0.01f);
                lastPdf = ediv((1.0f - specProb) * fmaxf(dot(nextD, N), 0.0f), OHB_PI); lastDelta = false;
            }
            if (primary) {
                P.fh0[p] = mk4(hp, rough); P.fh1[p] = mk4(N, metal); P.fh2[p] = mk4(albedo, 0.0f);
                if (fr.maxBounces >= 1u) { stage = ST_CHAIN_B; bounce = 1u; }
                else { startC = true; fhPos = hp; fhN = N; fhAlbedo = albedo; stage = ST_CHAIN_B; }
            } else if (bounce >= fr.maxBounces) chainEnds = true;
            else bounce += 1u;
        }
    }
    if (chainEnds) {
        if (specChain) {
            startC = true;
            f4 a0 = P.fh0[p], a1 = P.fh1[p], a2 = P.fh2[p];
            fhPos = xyz(a0); fhN = xyz(a1); fhAlbedo = xyz(a2);
        } else finished = true;
    }
    if (startC) {
        // Stage C set-up (:900-921): cosine direction around the FIRST hit's normal, throughput = albedo
        f2 du = cx.sm.get2D(cx.dimIdx); cx.dimIdx += 2u;
        nextD = cosineHemisphereShared(fhN, du);
        nextO = fhPos + fhN * 0.01f;
        thr = fhAlbedo;
        lastPdf = ediv(fmaxf(dot(nextD, fhN), 0.0f), OHB_PI); lastDelta = false;
        if (fr.maxBounces >= 1u) { stage = ST_CHAIN_C; bounce = 1u; } else finished = true;
    }
    P.rad[p] = mk4(rad, 0.0f);
    uint32_t keep = cx.state & (OHB_ST_PEND_A | OHB_ST_PEND_B);
    u4 mo = m;
    mo.z = (fr.samplerType == OHB_SAMPLER_PCG) ? cx.sm.pcg : cx.dimIdx;
    if (finished) { mo.w = ST_DONE | keep; P.meta[p] = mo; return OHB_Q_NONE; }
    mo.w = OHB_ST_MAKE(stage, bounce) | keep | (lastDelta ? OHB_ST_DELTA : 0u);
    P.meta[p] = mo;
    P.thr[p] = mk4(thr, lastPdf);
    P.rayO[p] = mk4(nextO, 0.0f); P.rayD[p] = mk4(nextD, 0.0f);
    return p | ((keep & OHB_ST_PEND_A) ? OHB_Q_PEND_A : 0u) | ((keep & OHB_ST_PEND_B) ? OHB_Q_PEND_B : 0u);
}
// k_shade: both shading stages in one kernel, the payload stays in registers (no 64-B payload round trip through HBM).
template <int PRIM = 2>
OHB_HD uint32_t shadePath(const SceneDev& sc, const FrameParams& fr, const PathArrays& P, uint32_t e) {
    f4 q0, q1, q2, q3;
    const bool hit = surfaceShade(sc, fr, P, OHB_Q_PATH(e), q0, q1, q2, q3);
    return bounceBody<PRIM>(sc, fr, P, hit ? e : (e | OHB_Q_MISS), q0, q1, q2, q3);
}

// ---------------------------------------------------------------------------------------------
// Film: pt_raygen_offline.rgen:1205-1262 (firefly clamp, own-pixel running mean) and :1339-1354
// (ACES(0.5 x) + gamma 2.2 -> RGBA8).  One thread per pixel walks its samples in index order so
// the float sequence equals `samplesInBatch` consecutive dispatches of the reference.
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// NRD front-end packing (shaders/includes/rt/nrd_frontend.glsl:11-41, pt_raygen_offline.rgen:106-127): what the raygen
// applies to its diffuse / specular radiance + hit distance and to the first-hit normal + roughness before they are
// written to the NRD AOV images.  AOV-only — the beauty path never reads them.
// ---------------------------------------------------------------------------------------------
OHB_HD f3 nrdLinearToYCoCg(f3 c) { return mk3(dot(c, mk3(0.25f, 0.5f, 0.25f)), dot(c, mk3(0.5f, 0.0f, -0.5f)), dot(c, mk3(-0.25f, 0.5f, -0.25f))); }
OHB_HD f3 nrdYCoCgToLinear(f3 c) { float t = c.x - c.z; return vmax(mk3(t + c.y, c.x + c.z, t - c.y), mk3(0.0f)); }
OHB_HD float nrdNormHitDist(float hitDist, float viewZ, float roughness) {      // saturate(hitDist / ((A + |viewZ| B) mix(C, 1, rough^2))), A=3 B=0.1 C=20
    float smc = clampf(roughness * roughness, 0.0f, 1.0f);
    float f = (3.0f + fabsf(viewZ) * 0.1f) * mixf(20.0f, 1.0f, smc);
    return clampf(hitDist / fmaxf(f, 1e-6f), 0.0f, 1.0f);
}
OHB_HD f4 nrdPackRadianceHitDist(f3 rad, float hitDist, float viewZ, float roughness) {
    return mk4(nrdLinearToYCoCg(vmax(rad, mk3(0.0f))), nrdNormHitDist(fmaxf(hitDist, 0.0f), viewZ, roughness));
}
OHB_HD f4 nrdPackNormalRoughness(f3 n, float roughness) {                        // rotated octahedron + sign of n.z smuggled into the roughness channel
    float l1 = fabsf(n.x) + fabsf(n.y) + fabsf(n.z);
    n = mk3(n.x / l1, n.y / l1, n.z / l1);
    f3 r;
    r.y = n.y * 0.5f + 0.5f; r.x = n.x * 0.5f + r.y; r.y -= n.x * 0.5f;
    roughness = fmaxf(roughness, 1.5f / 512.0f);
    float s = (n.z < 0.0f) ? -roughness : roughness;
    r.z = s * 0.5f + 0.5f;
    return mk4(r, 0.0f);
}

OHB_HD f3 ACES(f3 x) {
    f3 n = x * (2.51f * x + mk3(0.03f)), dd = x * (2.43f * x + mk3(0.59f)) + mk3(0.14f);
    f3 r = n / dd;
    return mk3(clampf(r.x, 0.0f, 1.0f), clampf(r.y, 0.0f, 1.0f), clampf(r.z, 0.0f, 1.0f));
}
OHB_HD uint32_t tonemapRGBA8(f3 acc) {
    f3 ldr = vpow(ACES(acc * 0.5f), 1.0f / 2.2f);
    uint32_t r = uint32_t(rintf(clampf(ldr.x, 0.0f, 1.0f) * 255.0f)), g = uint32_t(rintf(clampf(ldr.y, 0.0f, 1.0f) * 255.0f)), b = uint32_t(rintf(clampf(ldr.z, 0.0f, 1.0f) * 255.0f));
    return r | (g << 8) | (b << 16) | 0xFF000000u;
}
struct FilmArrays {
    f4* accum; uint32_t* ldr; float* sampleDump;   // sampleDump: [s][H][W][4] or null
    uint32_t historyCount;                          // PathTracer::m_historyFrameCount at s = 0
    int sumMode;                                    // 1: accum holds (sum.rgb, count) for sharded renders
};
OHB_HD void filmPixel(const FrameParams& fr, const PathArrays& P, const FilmArrays& F, uint32_t pix) {
    u4 m0 = P.meta[pix];
    if (m0.x == 0xFFFFFFFFu) return;
    uint32_t px = m0.x & 0xFFFFu, py = m0.x >> 16;
    size_t pi = size_t(py) * fr.W + px;
    bool clampOn = (fr.flags & OHB_FLAG_ENABLE_FIREFLY_CLAMP) && fr.fireflyClamp > 0.0f;
    f4 acc = F.accum[pi];
    for (uint32_t s = 0; s < P.samplesInBatch; s++) {
        uint32_t p = s * P.numPixels + pix;
        uint32_t st = P.meta[p].w;
        f3 rad = xyz(P.rad[p]);
        if (st & OHB_ST_PEND_A) rad += xyz(P.pendA[p]);
        if (st & OHB_ST_PEND_B) rad += xyz(P.pendB[p]);
        if (clampOn) clampLum(rad, fr.fireflyClamp);
        if (F.sampleDump) {
            float* sd = F.sampleDump + (size_t(s) * fr.W * fr.H + pi) * 4u;
            sd[0] = rad.x; sd[1] = rad.y; sd[2] = rad.z; sd[3] = 1.0f;
        }
        if (F.sumMode) {
            if (F.historyCount + s == 0u) acc = mk4(rad, 1.0f); else acc = mk4(acc.x + rad.x, acc.y + rad.y, acc.z + rad.z, acc.w + 1.0f);
        } else if (F.historyCount + s == 0u) acc = mk4(rad, 1.0f);
        else {
            float cnt = acc.w + 1.0f;
            acc = mk4((acc.x * acc.w + rad.x) / cnt, (acc.y * acc.w + rad.y) / cnt, (acc.z * acc.w + rad.z) / cnt, cnt);
        }
    }
    F.accum[pi] = acc;
    f3 mean = F.sumMode ? xyz(acc) / fmaxf(acc.w, 1.0f) : xyz(acc);
    F.ldr[pi] = tonemapRGBA8(mean);
}

}  // namespace ohb

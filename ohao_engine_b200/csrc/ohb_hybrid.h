// ohb_hybrid.h — the hybrid-RT techniques of the deferred path, the other two consumers of the TLAS:
//   hybridShadowPixel : shaders/rt/rt_shadow.rgen:41-111  (soft shadow mask from G-buffer position + normal, any-hit rays)
//   hybridGiPixel     : shaders/rt/rt_gi.rgen:59-134 + rt_gi.rchit:16-28 + rt_gi.rmiss (1-bounce diffuse GI, temporal blend)
// One thread per pixel, the rays traced inline (traceAny / traceClosest of ohb_traverse.h): a G-buffer pass launches
// W x H x sampleCount short rays once per frame, there is no path state to queue.  G-buffer images are caller-supplied
// (the raster path that produces them is out of scope): position RGBA32F (rgb world position, all-zero = sky), normal as
// the two octahedron channels in [0,1] (includes/common/encoding.glsl:15-22), albedo RGBA32F, GI history RGBA32F.
// Differences from the Vulkan path: per-ray cull masks do not exist here — every instance visible to mask 0xFF is traced;
// the GI closest-hit's "animated instance = miss" workaround (rt_gi.rchit:20-25, material alpha < 0.5) is kept.
#pragma once
#include "ohb_traverse.h"
#include "ohb_svgf.h"

namespace ohb {

struct HybridShadowParams { f3 lightDir; float lightRadius; f3 lightPos; float lightRange; uint32_t W, H, lightType, sampleCount; };
struct HybridGiParams { f3 lightPos; float lightIntensity; uint32_t W, H, sampleCount, frameIndex; };

OHB_HD float fractf(float x) { return xsub(x, floorf(x)); }
// The hash multiplies its rounding errors by ~10^3 (fract of a product of magnitude ~2000), so its operation order is part of
// the spec shared with the oracle: one rounding per operation, never contracted into an FMA.
OHB_HD float hybridHash(float px, float py) {                       // rt_shadow.rgen:25-29 == rt_gi.rgen:30-34
    f3 p3 = mk3(fractf(xmul(px, 0.1031f)), fractf(xmul(py, 0.1031f)), fractf(xmul(px, 0.1031f)));
    float d = xadd(xadd(xmul(p3.x, xadd(p3.y, 33.33f)), xmul(p3.y, xadd(p3.z, 33.33f))), xmul(p3.z, xadd(p3.x, 33.33f)));
    p3 = mk3(xadd(p3.x, d), xadd(p3.y, d), xadd(p3.z, d));
    return fractf(xmul(xadd(p3.x, p3.y), p3.z));
}
OHB_HD f3 decodeNormalOctahedron(float ex, float ey) {              // encoding.glsl:15-22
    float fx = ex * 2.0f - 1.0f, fy = ey * 2.0f - 1.0f;
    f3 n = mk3(fx, fy, 1.0f - fabsf(fx) - fabsf(fy));
    float t = clampf(-n.z, 0.0f, 1.0f);
    n.x += (n.x >= 0.0f) ? -t : t;
    n.y += (n.y >= 0.0f) ? -t : t;
    return normalize(n);
}
OHB_HD void hybridBasis(f3 N, f3& T, f3& B) {                       // buildBasis, rt_shadow.rgen:32-36
    f3 up = fabsf(N.y) < 0.999f ? mk3(0.0f, 1.0f, 0.0f) : mk3(1.0f, 0.0f, 0.0f);
    T = normalize(cross(up, N)); B = cross(N, T);
}
// returns the R8_UNORM texel the reference's imageStore(shadowMask, ...) writes
OHB_HD uint8_t hybridShadowPixel(const SceneDev& sc, const HybridShadowParams& pc, const f4* gPos, const f2* gNrm, uint32_t x, uint32_t y) {
    const size_t pi = size_t(y) * pc.W + x;
    const f4 ps = gPos[pi];
    if (ps.x == 0.0f && ps.y == 0.0f && ps.z == 0.0f && ps.w == 0.0f) return 255;          // sky
    const f3 worldPos = xyz(ps), N = decodeNormalOctahedron(gNrm[pi].x, gNrm[pi].y);
    const f3 origin = worldPos + N * 0.05f;
    const uint32_t sampleCount = pc.sampleCount > 1u ? pc.sampleCount : 1u;
    float visibility = 0.0f;
    for (uint32_t s = 0; s < sampleCount; s++) {
        const float fs = float(s);
        const float r1 = hybridHash(xadd(float(x), xmul(fs, 7.13f)), xadd(float(y), xmul(fs, 13.37f))), r2 = hybridHash(xadd(float(x), xmul(fs, 31.17f)), xadd(float(y), xmul(fs, 47.53f)));
        f3 L; float tMax;
        if (pc.lightType == 0u) {
            f3 lightDir = normalize(-pc.lightDir), T, B; hybridBasis(lightDir, T, B);
            float angle = pc.lightRadius * sqrtf(r1), phi = 6.2831853f * r2;
            L = normalize(lightDir + T * (angle * cosf(phi)) + B * (angle * sinf(phi)));
            tMax = 10000.0f;
        } else {
            float theta = 6.2831853f * r1, phi = acosf(1.0f - 2.0f * r2);
            f3 offset = mk3(sinf(phi) * cosf(theta), sinf(phi) * sinf(theta), cosf(phi)) * pc.lightRadius;
            f3 toLight = (pc.lightPos + offset) - worldPos;
            float dist = length(toLight);
            L = toLight / dist; tMax = dist;
        }
        if (dot(N, L) <= 0.0f) continue;
        // shadowPayload: 0 on any hit, 1 from the miss shader (rt_shadow.rmiss)
        if (!traceAny(sc, origin, L, 0.001f, tMax)) visibility += 1.0f;
    }
    visibility /= float(sampleCount);
    return uint8_t(rintf(clampf(visibility, 0.0f, 1.0f) * 255.0f));
}
OHB_HD f3 hybridCosineHemisphere(float ux, float uy, f3 N) {       // rt_gi.rgen:41-56
    f3 T, B; hybridBasis(N, T, B);
    float r = sqrtf(ux), phi = 6.2831853f * uy;
    float cx = r * cosf(phi), cy = r * sinf(phi), cz = sqrtf(fmaxf(0.0f, 1.0f - ux));
    return normalize(T * cx + B * cy + N * cz);
}
// writes the RGBA16F texel of giOutput (as fp16 bits)
OHB_HD h4 hybridGiPixel(const SceneDev& sc, const HybridGiParams& pc, const f4* gPos, const f2* gNrm, const f4* gAlbedo, const f4* history,
                        const f4* instMaterials, uint32_t x, uint32_t y) {
    const size_t pi = size_t(y) * pc.W + x;
    const f4 ps = gPos[pi];
    if (ps.x == 0.0f && ps.y == 0.0f && ps.z == 0.0f && ps.w == 0.0f) return packH4(0.0f, 0.0f, 0.0f, 0.0f);
    const f3 worldPos = xyz(ps), N = decodeNormalOctahedron(gNrm[pi].x, gNrm[pi].y), albedo = xyz(gAlbedo[pi]);
    const f3 origin = worldPos + N * 0.05f;
    const uint32_t sampleCount = pc.sampleCount > 1u ? pc.sampleCount : 1u;
    f3 indirect = mk3(0.0f);
    for (uint32_t s = 0; s < sampleCount; s++) {
        const float fs = float(s), hx = xadd(float(x), xadd(xmul(fs, 7.13f), xmul(float(pc.frameIndex), 1.618f))), hy = xadd(float(y), xmul(fs, 13.37f));
        const f3 dir = hybridCosineHemisphere(hybridHash(hx, hy), hybridHash(xadd(hx, 127.1f), xadd(hy, 311.7f)), N);
        const ohb_hit h = traceClosest(sc, origin, dir, 0.01f, 100.0f);
        if (h.prim == OHB_MISS) continue;
        const f4 m = instMaterials[sc.triInst[h.prim]];                 // materials[gl_InstanceID]
        if (m.w < 0.5f) continue;                                         // animated instance: treated as a miss (rt_gi.rchit:20-25)
        const f3 hitPos = origin + dir * h.t;
        const f3 toLight = pc.lightPos - hitPos;
        const float lightDist = length(toLight), falloff = pc.lightIntensity / (1.0f + lightDist * lightDist);
        const float hitNdotL = fmaxf(dot(normalize(toLight), -dir), 0.0f);
        indirect += xyz(m) * falloff * hitNdotL;
    }
    indirect = indirect / float(sampleCount);
    indirect = indirect * albedo;
    const f3 hist = xyz(history[pi]);
    const float blend = pc.frameIndex == 0u ? 1.0f : 0.3f;
    const f3 acc = hist * (1.0f - blend) + indirect * blend;              // mix(history, indirect, blend)
    return packH4(acc.x, acc.y, acc.z, 1.0f);
}

}  // namespace ohb

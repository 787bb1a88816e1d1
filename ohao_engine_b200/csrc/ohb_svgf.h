// ohb_svgf.h — SVGF denoiser of DenoiseMode::Atrous (per-pixel code of the two compute passes).
//
// Reference: ohao/render/rt/denoise/atrous_denoise.cpp:392-537 (dispatch schedule, sigmas :28-40),
// shaders/rt/rt_svgf_temporal.comp:45-126 (pass 1: reprojection, disocclusion test, EMA of colour and
// luminance moments, variance with 3x3 spatial bootstrap) and shaders/rt/rt_svgf_atrous.comp:54-115
// (pass 2: five variance-guided 5x5 B3-spline a-trous iterations, step 1,2,4,8,16; the last one writes the
// RGBA8 beauty).  Inputs come from the realtime raygen: RGBA8 beauty (binding 2), normal AOV (N*0.5+0.5),
// linear view-Z AOV (binding 20, 1e30 = background) and pixel-space motion vectors (binding 19, RG16F,
// pt_raygen_realtime.rgen:412-455).  History and scratch images are RGBA16F / R16F in the reference: the same
// fp16 storage is used here (round-to-nearest-even on every store), so the filter sees the same quantisation.
//
// Bandwidth-bound stencil: per iteration 25 taps x (8 B colour + 16 B normal + 4 B depth + 2 B variance), all
// L1/L2 hits after the first touch; DRAM traffic is one read + one write of the 14 B/pixel colour+variance planes.
#pragma once
#include "ohb_scene.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace ohb {

// ---- fp16 storage --------------------------------------------------------------------------------
OHB_HD uint16_t f2h(float f) {
#if OHB_DEVICE_CODE
    return __half_as_ushort(__float2half_rn(f));
#else
    uint32_t x = f2u(f), sign = (x >> 16) & 0x8000u, em = x & 0x7FFFFFFFu;
    if (em >= 0x7F800000u) return uint16_t(sign | (em > 0x7F800000u ? 0x7E00u : 0x7C00u));        // NaN / inf
    if (em >= 0x477FF000u) return uint16_t(sign | 0x7C00u);                                         // rounds to inf (>= 65520)
    if (em < 0x33000001u) return uint16_t(sign);                                                    // <= 2^-25: rounds to zero
    int e = int(em >> 23) - 127; uint32_t man = (em & 0x7FFFFFu) | 0x800000u;
    int shift = e < -14 ? (13 + (-14 - e)) : 13;                                                     // subnormal halves lose more bits
    uint32_t q = man >> shift, rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    uint32_t h = e < -14 ? q : ((uint32_t(e + 15) << 10) + (q - 0x400u));                            // carry out of the mantissa bumps the exponent
    return uint16_t(sign | h);
#endif
}
OHB_HD float h2f(uint16_t h) {
#if OHB_DEVICE_CODE
    return __half2float(__ushort_as_half(h));
#else
    uint32_t sign = uint32_t(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 0x3FFu;
    if (e == 31u) return u2f(sign | 0x7F800000u | (m << 13));
    if (e == 0u) { float v = float(m) * 5.9604644775390625e-8f; return (h & 0x8000u) ? -v : v; }      // m * 2^-24
    return u2f(sign | ((e + 112u) << 23) | (m << 13));
#endif
}
struct alignas(8) h4 { uint16_t x, y, z, w; };
OHB_HD h4 packH4(float a, float b, float c, float d) { h4 r; r.x = f2h(a); r.y = f2h(b); r.z = f2h(c); r.w = f2h(d); return r; }
OHB_HD f4 unpackH4(h4 v) { return mk4(h2f(v.x), h2f(v.y), h2f(v.z), h2f(v.w)); }
OHB_HD f3 unorm8rgb(uint32_t p) { return mk3(float(p & 255u) / 255.0f, float((p >> 8) & 255u) / 255.0f, float((p >> 16) & 255u) / 255.0f); }
OHB_HD uint32_t packUnorm8(f3 c) {
    uint32_t r = uint32_t(rintf(clampf(c.x, 0.0f, 1.0f) * 255.0f)), g = uint32_t(rintf(clampf(c.y, 0.0f, 1.0f) * 255.0f)), b = uint32_t(rintf(clampf(c.z, 0.0f, 1.0f) * 255.0f));
    return r | (g << 8) | (b << 16) | 0xFF000000u;
}
OHB_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- AOVs the denoiser consumes, written by k_rt_pixel for the pixel's first hit (pt_raygen_realtime.rgen:412-455)
// motion = currPix - prevPix of the first hit under proj*view of this and of the previous frame (zero on a miss, on
// frame 0, or when either clip w <= 0); depth = -(view * pos).z, 1e30 on a miss.
OHB_HD void svgfGuides(const float* currVP, const float* prevVP, const float* viewRow2, uint32_t W, uint32_t H, uint32_t frameIdx,
                       bool hit, f3 p, uint32_t& motionOut, float& depthOut) {
    float mx = 0.0f, my = 0.0f;
    if (hit && frameIdx > 0u) {
        float cx = currVP[0] * p.x + currVP[4] * p.y + currVP[8] * p.z + currVP[12], cy = currVP[1] * p.x + currVP[5] * p.y + currVP[9] * p.z + currVP[13];
        float cw = currVP[3] * p.x + currVP[7] * p.y + currVP[11] * p.z + currVP[15];
        float qx = prevVP[0] * p.x + prevVP[4] * p.y + prevVP[8] * p.z + prevVP[12], qy = prevVP[1] * p.x + prevVP[5] * p.y + prevVP[9] * p.z + prevVP[13];
        float qw = prevVP[3] * p.x + prevVP[7] * p.y + prevVP[11] * p.z + prevVP[15];
        if (cw > 0.0f && qw > 0.0f) {
            float cpx = (cx / cw * 0.5f + 0.5f) * float(W), cpy = (cy / cw * 0.5f + 0.5f) * float(H);
            float ppx = (qx / qw * 0.5f + 0.5f) * float(W), ppy = (qy / qw * 0.5f + 0.5f) * float(H);
            mx = cpx - ppx; my = cpy - ppy;
        }
    }
    motionOut = uint32_t(f2h(mx)) | (uint32_t(f2h(my)) << 16);
    depthOut = hit ? -(viewRow2[0] * p.x + viewRow2[1] * p.y + viewRow2[2] * p.z + viewRow2[3]) : 1e30f;
}

struct SvgfTemporalArgs {
    const uint32_t* beauty; const uint32_t* motion; const float* depth; const f4* normal;
    const h4* prevColor; const h4* prevMoments; const h4* prevGeom;
    h4* outColor; h4* outMoments; uint16_t* outVariance; h4* outGeom;
    int W, H, reset;
};
OHB_HD float svgfLuma(f3 c) { return dot(c, mk3(0.2126f, 0.7152f, 0.0722f)); }

// rt_svgf_temporal.comp:61-126
OHB_HD void svgfTemporalPixel(const SvgfTemporalArgs& a, int px, int py) {
    const size_t pi = size_t(py) * a.W + px;
    f3 cur = unorm8rgb(a.beauty[pi]);
    float curL = svgfLuma(cur);
    float curZ = a.depth[pi];
    f3 curN = xyz(a.normal[pi]) * 2.0f - mk3(1.0f);
    curN = dot(curN, curN) > 1e-8f ? normalize(curN) : mk3(0.0f, 0.0f, 1.0f);
    float curZc = fminf(curZ, 1.0e4f);
    uint32_t mvp = a.motion[pi];
    float mvx = h2f(uint16_t(mvp & 0xFFFFu)), mvy = h2f(uint16_t(mvp >> 16));
    int qx = int(floorf(float(px) - mvx + 0.5f)), qy = int(floorf(float(py) - mvy + 0.5f));
    bool valid = a.reset == 0;
    if (qx < 0 || qy < 0 || qx >= a.W || qy >= a.H) valid = false;
    const size_t qi = valid ? size_t(qy) * a.W + qx : 0;
    if (valid) {
        f4 g = unpackH4(a.prevGeom[qi]);
        float zrel = fabsf(curZc - g.x) / fmaxf(fmaxf(fabsf(curZc), fabsf(g.x)), 1e-3f);
        float ndot = dot(curN, mk3(g.y, g.z, g.w));
        if (zrel > 0.1f || ndot < 0.9f) valid = false;
    }
    float prevLen = 0.0f; f3 prevCol = cur; float prevM1 = curL, prevM2 = curL * curL;
    if (valid) {
        f4 hc = unpackH4(a.prevColor[qi]), hm = unpackH4(a.prevMoments[qi]);
        prevCol = xyz(hc); prevM1 = hm.x; prevM2 = hm.y; prevLen = hm.z;
    }
    float newLen = valid ? fminf(prevLen + 1.0f, 32.0f) : 1.0f;
    float alpha = fmaxf(1.0f / newLen, 0.05f);
    f3 accumCol = mix(prevCol, cur, alpha);
    float m1 = mixf(prevM1, curL, alpha), m2 = mixf(prevM2, curL * curL, alpha);
    float variance = fmaxf(0.0f, m2 - m1 * m1);
    if (newLen < 4.0f) {
        float s = 0.0f, s2 = 0.0f;
        for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
            int x = clampi(px + dx, 0, a.W - 1), y = clampi(py + dy, 0, a.H - 1);
            float lq = svgfLuma(unorm8rgb(a.beauty[size_t(y) * a.W + x]));
            s += lq; s2 += lq * lq;
        }
        float mm = s / 9.0f;
        variance = fmaxf(variance, fmaxf(0.0f, s2 / 9.0f - mm * mm));
    }
    a.outColor[pi] = packH4(accumCol.x, accumCol.y, accumCol.z, 1.0f);
    a.outMoments[pi] = packH4(m1, m2, newLen, 0.0f);
    a.outVariance[pi] = f2h(variance);
    a.outGeom[pi] = packH4(curZc, curN.x, curN.y, curN.z);
}

struct SvgfAtrousArgs {
    const h4* inColor; h4* outColor16; const f4* normal; const float* depth; const uint16_t* inVar; uint16_t* outVar; uint32_t* outLDR;
    int W, H, stepSize, isFinal; float sigmaL, sigmaNormal, sigmaDepth;
};
// rt_svgf_atrous.comp:54-115
OHB_HD void svgfAtrousPixel(const SvgfAtrousArgs& a, int px, int py) {
    const size_t pi = size_t(py) * a.W + px;
    const float kern[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
    const float gk[3] = {0.25f, 0.5f, 0.25f};
    f3 cColor = xyz(unpackH4(a.inColor[pi]));
    f3 cN = xyz(a.normal[pi]) * 2.0f - mk3(1.0f);
    float cD = a.depth[pi], cL = svgfLuma(cColor);
    float gVar = 0.0f, gW = 0.0f;
    for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
        int x = clampi(px + dx, 0, a.W - 1), y = clampi(py + dy, 0, a.H - 1);
        float w = gk[dx + 1] * gk[dy + 1];
        gVar += w * h2f(a.inVar[size_t(y) * a.W + x]); gW += w;
    }
    float centerVar = gW > 0.0f ? gVar / gW : h2f(a.inVar[pi]);
    float sqrtVar = sqrtf(fmaxf(centerVar, 1e-8f));
    // the three edge-stop denominators are loop invariants: one reciprocal each instead of 75 IEEE divisions per pixel
    // (x * (1/d) vs x / d differ in the last ulp of the exponent argument, the same class as expf vs libm exp)
    const float invL = -1.0f / (sqrtVar * a.sigmaL + 1e-6f), invN = -1.0f / (a.sigmaNormal * a.sigmaNormal + 1e-4f), invD = -1.0f / (a.sigmaDepth * a.sigmaDepth + 1e-4f);
    f3 sum = mk3(0.0f); float weightSum = 0.0f, varSum = 0.0f;
    for (int dy = -2; dy <= 2; dy++) for (int dx = -2; dx <= 2; dx++) {
        int x = clampi(px + dx * a.stepSize, 0, a.W - 1), y = clampi(py + dy * a.stepSize, 0, a.H - 1);
        size_t qi = size_t(y) * a.W + x;
        f3 sColor = xyz(unpackH4(a.inColor[qi]));
        f3 sN = xyz(a.normal[qi]) * 2.0f - mk3(1.0f);
        float sD = a.depth[qi], sVar = h2f(a.inVar[qi]);
        float w = kern[dx + 2] * kern[dy + 2];
        w *= expf(fabsf(cL - svgfLuma(sColor)) * invL);
        w *= expf(fmaxf(1.0f - dot(cN, sN), 0.0f) * invN);
        w *= expf(fabsf(cD - sD) * invD);
        sum += sColor * w; weightSum += w; varSum += w * w * sVar;
    }
    f3 outC = weightSum > 1e-6f ? mk3(sum.x / weightSum, sum.y / weightSum, sum.z / weightSum) : cColor;
    float outV = weightSum > 1e-6f ? varSum / (weightSum * weightSum) : centerVar;
    if (a.isFinal) a.outLDR[pi] = packUnorm8(outC);
    else a.outColor16[pi] = packH4(outC.x, outC.y, outC.z, 1.0f);
    a.outVar[pi] = f2h(outV);
}

#define OHB_SVGF_ITERATIONS 5
#define OHB_SVGF_SIGMA_L 0.4f
#define OHB_SVGF_SIGMA_NORMAL 0.30f
#define OHB_SVGF_SIGMA_DEPTH 2.0f

}  // namespace ohb

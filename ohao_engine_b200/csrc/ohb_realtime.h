// ohb_realtime.h — wavefront form of the REALTIME integrator (shaders/rt/pt_raygen_realtime.rgen:277-1925).
//
// One frame = N decorrelated path trees per pixel (N = samplesPerFrame) + ReSTIR GI + reprojected EMA
// accumulation + in-shader a-trous.  The path trees reuse the offline machinery (queues, k_trace_*, k_surface):
//   stage PRIMARY  -> Stage A: emissive, light NEE (solid-angle sphere lights), env NEE with the VNDF MIS pdf
//   stage CHAIN_B  -> Stage B: VNDF first glossy bounce, then the crude lobe chain, RR clamped to [0.1, 0.95]
//   stage GI       -> ReSTIR GI initial sample: one cosine bounce from the FIRST hit to x_s, Lo = emission +
//                     light NEE + env NEE at x_s (their visibility answers are parked in pendA/pendB like any NEE)
// and a per-pixel kernel (pixelRT) then walks the pixel's N finished samples IN ORDER — radiance mean, RIS
// streaming of the N candidates — and runs temporal + spatial resampling (<= 5 visibility rays, traced inline),
// shading, history writes and the EMA.  denoiseRT is the a-trous + tonemap pass.
//
// Deliberate differences from the reference, shared with the oracle (oracle/oracle_realtime.inl): history images are
// double-buffered (the reference races on accumBuffer inside one dispatch, quirk Q12, and never flips its reservoir
// images, Q2) and the a-trous pass runs after the whole frame has accumulated.  NRD/DLSS AOVs are not produced.
#pragma once
#include "ohb_integrator.h"
#include "ohb_svgf.h"

namespace ohb {

#define ST_GI 2u                       // reuses the stage code of the offline Stage C
#define OHB_ST_GI_HIT   (1u << 9)      // finished path carries a reservoir candidate (x_s in pay0/pay1, Lo base in pay3)
#define OHB_ST_GI_MISS  (1u << 10)     // finished path's cosine ray escaped: env colour in pay0.xyz

// ggx_aniso.glsl:66-126
OHB_HD float ggxDiso(float NdotH, float alpha) { float a2 = alpha * alpha; float dn = NdotH * NdotH * (a2 - 1.0f) + 1.0f; return ediv(a2, 3.14159265f * dn * dn + 1e-8f); }
OHB_HD float smithLambdaGGX(float c, float alpha) { float c2 = c * c; float tan2 = ediv(fmaxf(0.0f, 1.0f - c2), fmaxf(c2, 1e-8f)); return 0.5f * (-1.0f + sqrtf(1.0f + alpha * alpha * tan2)); }
OHB_HD float smithG1GGX(float c, float alpha) { return ediv(1.0f, 1.0f + smithLambdaGGX(c, alpha)); }
OHB_HD float smithG2overG1GGX(float NdotV, float NdotL, float alpha) { float lv = smithLambdaGGX(NdotV, alpha), ll = smithLambdaGGX(NdotL, alpha); return ediv(1.0f + lv, 1.0f + lv + ll + 1e-8f); }
OHB_HD f3 sampleGGXVNDF(f3 Ve, float ax, float ay, f2 u) {
    f3 Vh = normalize(mk3(ax * Ve.x, ay * Ve.y, Ve.z));
    float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    f3 T1 = lensq > 0.0f ? mk3(-Vh.y, Vh.x, 0.0f) * (1.0f / sqrtf(lensq)) : mk3(1, 0, 0);
    f3 T2 = cross(Vh, T1);
    float r = sqrtf(u.x);
    float sp, cp; ohb_sincos_turns(u.y, sp, cp);                     // phi = 2 * 3.14159265 * u.y
    float t1 = r * cp, t2 = r * sp;
    float s = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - s) * sqrtf(fmaxf(0.0f, 1.0f - t1 * t1)) + s * t2;
    f3 Nh = t1 * T1 + t2 * T2 + sqrtf(fmaxf(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalize(mk3(ax * Nh.x, ay * Nh.y, fmaxf(0.0f, Nh.z)));
}

// ReSTIR GI helpers: pt_raygen_realtime.rgen:213-275
struct GIReservoir { f3 xs, ns, Lo; float wSum, M, W; };
OHB_HD float giTargetPHat(f3 albedoD, f3 n1, f3 x1, f3 xs, f3 Lo) {
    f3 d = xs - x1; float len = length(d);
    if (len < 1e-5f) return 0.0f;
    d /= len;
    float cosT = fmaxf(dot(n1, d), 0.0f);
    if (cosT <= 0.0f) return 0.0f;
    return luminance((albedoD / OHB_PI) * Lo * cosT);
}
OHB_HD float giSpatialJacobian(f3 ns, f3 xs, f3 x1r, f3 x1q) {
    f3 toR = x1r - xs; float dr2 = dot(toR, toR);
    f3 toQ = x1q - xs; float dq2 = dot(toQ, toQ);
    float dr = sqrtf(fmaxf(dr2, 1e-12f)), dq = sqrtf(fmaxf(dq2, 1e-12f));
    if (dr < 1e-3f || dq < 1e-3f) return 0.0f;
    f3 nsN = normalize(ns);
    float cosR = fabsf(dot(nsN, toR / dr)), cosQ = fabsf(dot(nsN, toQ / dq));
    if (cosQ < 1e-4f) return 0.0f;
    return clampf((cosR * dq2) / fmaxf(cosQ * dr2, 1e-8f), 1e-3f, 1e3f);
}
OHB_HD f3 giShade(f3 albedoD, f3 n1, f3 x1, f3 xs, f3 Lo, float W) {
    f3 d = xs - x1; float len = length(d);
    if (len < 1e-5f) return mk3(0.0f);
    d /= len;
    return (albedoD / OHB_PI) * Lo * fmaxf(dot(n1, d), 0.0f) * W;
}

// BRDF of the NEE blocks (bounce 0: ggxD_anisoOrIso, later: inline isotropic D with the block's own pi literal).
struct BrdfEval { f3 diff, spec; float D, NdotH, VdotH, NdotV, NdotL; };
OHB_HD BrdfEval evalBrdfRT(const FrameParams& fr, f3 N, f3 V, f3 L, f3 albedo, f3 F0, float rough, float kdScale, bool bounce0, float piDiff) {
    BrdfEval e;
    f3 H = normalize(L + V);
    e.NdotL = fmaxf(dot(N, L), 0.0f);
    e.NdotH = fmaxf(dot(N, H), 0.001f); e.NdotV = fmaxf(dot(N, V), 0.001f); e.VdotH = fmaxf(dot(V, H), 0.001f);
    if (bounce0) e.D = ggxD_anisoOrIso(N, H, e.NdotH, rough, fr.aniso, fr.anisoRot);
    else { float a = rough * rough, a2 = a * a; float dn = e.NdotH * e.NdotH * (a2 - 1.0f) + 1.0f; e.D = ediv(a2, piDiff * dn * dn + 0.0001f); }
    f3 F = schlick(F0, e.VdotH);
    float k = (rough + 1.0f) * (rough + 1.0f) / 8.0f;
    float G = ediv(e.NdotL, e.NdotL * (1.0f - k) + k) * ediv(e.NdotV, e.NdotV * (1.0f - k) + k);
    e.spec = e.D * F * G / (4.0f * e.NdotV * e.NdotL + 0.001f);
    e.diff = ((mk3(1.0f) - F) * kdScale) * albedo / piDiff;
    return e;
}
// light NEE (:528-660 bounce 0, :830-950 chain); gi = the x_s variant (:1401-1486): no throughput, no clamp, legacy metal flag
OHB_HD void lightNEE_RT(ShadeCtx& cx, f3 hp, f3 N, f3 inDir, f3 albedo, f3 F0, float rough, float kdScale, bool bounce0, bool gi, f3 thr) {
    LightSample ls = sampleLight(cx.sc, cx.sm, cx.dimIdx, hp, true);
    float NdotL = fmaxf(dot(N, ls.L), 0.0f);
    if (!(NdotL > 0.0f && ls.weight > 0.0f)) return;
    BrdfEval e = evalBrdfRT(cx.fr, N, normalize(-inDir), ls.L, albedo, F0, rough, kdScale, bounce0, 3.14159f);
    f3 c = (bounce0 || gi) ? (ls.Le * (e.diff + e.spec) * NdotL * ls.weight * float(cx.sc.lightCount))
                           : (thr * ls.Le * (e.diff + e.spec) * NdotL * ls.weight * float(cx.sc.lightCount));
    if (!gi && cx.clampOn) clampLum(c, cx.fr.fireflyClamp);
    cx.pushShadow(hp + N * 0.01f, ls.L, ls.shadowDist, 0u, c);
}
// env NEE + MIS (:668-716 bounce 0 with the VNDF pdf, :955-1000 chain, :1489-1520 at x_s with the cosine pdf only)
OHB_HD void envNEE_RT(ShadeCtx& cx, f3 hp, f3 N, f3 inDir, f3 albedo, f3 F0, float rough, float metal, float kdScale, bool bounce0, bool gi, f3 thr) {
    f2 eu = cx.sm.get2D(cx.dimIdx); cx.dimIdx += 2u;
    f3 envDir; float envPdf;
    sampleEnvMap(cx.sc, eu.x, eu.y, envDir, envPdf);
    float NdotL = fmaxf(dot(N, envDir), 0.0f);
    if (!(NdotL > 0.0f && envPdf > 0.0f)) return;
    f3 envRad = envRadiance(cx.sc, envDir);
    BrdfEval e = evalBrdfRT(cx.fr, N, normalize(-inDir), envDir, albedo, F0, rough, kdScale, bounce0, OHB_PI);
    float pdfDiff = ediv(NdotL, OHB_PI);
    float bsdfPdf;
    if (gi) bsdfPdf = pdfDiff;
    else {
        float specProb = specProbOf(inDir, N, F0, rough, metal);
        float pdfSpec = bounce0 ? ediv(smithG1GGX(e.NdotV, rough * rough) * e.D, 4.0f * e.NdotV + 1e-4f) : ediv(e.D * e.NdotH, 4.0f * e.VdotH + 1e-4f);
        bsdfPdf = mixf(pdfDiff, pdfSpec, specProb);
    }
    float w = misBalance(envPdf, bsdfPdf);
    f3 c = (bounce0 || gi) ? (envRad * (e.diff + e.spec) * NdotL * w / envPdf) : (thr * envRad * (e.diff + e.spec) * NdotL * w / envPdf);
    if (!gi && cx.clampOn) clampLum(c, cx.fr.fireflyClamp);
    cx.pushShadow(hp + N * 0.01f, envDir, 10000.0f, 1u, c);
}

// Camera ray (:337-356): the jitter comes from sample index frameIdx (dims 0,1), the path from frameIdx*N + s (dims 2..).
OHB_HD void raygenPathRT(const FrameParams& fr, const PathArrays& P, uint32_t p) {
    uint32_t s = p / P.numPixels, pix = p - s * P.numPixels;
    uint32_t tilesX = (fr.tileW + 7u) / 8u;
    uint32_t blk = pix / 32u, inb = pix & 31u;
    uint32_t lx = (blk % tilesX) * 8u + (inb & 7u), ly = (blk / tilesX) * 4u + (inb >> 3);
    u4 m; m.y = P.firstSampleIndex + s; m.z = 2u;
    if (lx >= fr.tileW || ly >= fr.tileH) { m.x = 0xFFFFFFFFu; m.w = ST_DONE; P.meta[p] = m; P.rad[p] = mk4(0, 0, 0, 0); return; }
    uint32_t px = fr.tileX + lx, py = fr.tileY + ly;
    m.x = px | (py << 16);
    Sampler sm; sm.init(fr.samplerType, px, py, fr.frameIdx, fr.jitterSobol);
    f2 j = sm.get2D(0u);
    float uvx = (float(px) + 0.5f + (j.x - 0.5f) + fr.jitX) / float(fr.W), uvy = (float(py) + 0.5f + (j.y - 0.5f) + fr.jitY) / float(fr.H);
    float nx = uvx * 2.0f - 1.0f, ny = uvy * 2.0f - 1.0f;
    f3 dir = normalize(fr.fwd + fr.right * nx * fr.tanX - fr.up * ny * fr.tanY);
    m.w = OHB_ST_MAKE(ST_PRIMARY, 0u);
    if (fr.samplerType == OHB_SAMPLER_PCG) { Sampler ps; ps.init(fr.samplerType, px, py, m.y, fr.jitterSobol); m.z = ps.pcg; }
    P.meta[p] = m;
    P.rayO[p] = mk4(fr.camPos, 0.0f); P.rayD[p] = mk4(dir, 0.0f);
    P.rad[p] = mk4(0, 0, 0, 0);
}

// k_bounce_rt: per-bounce body of the realtime raygen for path p (after k_surface).  Returns true if the path traces again.
// FUSED: the closest-hit / miss shader runs here and its 64-B payload stays in registers (k_shade_rt), like the offline k_shade;
// only the payload pixelRT reads later — the ReSTIR GI sample's (x_s, n_s, Lo) — still goes to memory.  FUSED = false: the payload
// was written by k_surface (host emulator, A/B).
template <bool FUSED = false>
OHB_HD bool bouncePathRT(const SceneDev& sc, const FrameParams& fr, const PathArrays& P, uint32_t p) {
    ShadeCtx cx(sc, fr, P);
    u4 m = P.meta[p];
    cx.path = p; cx.state = m.w;
    uint32_t px = m.x & 0xFFFFu, py = m.x >> 16;
    cx.sm.init(fr.samplerType, px, py, m.y, ldu4(P.sobolTab + (m.y - P.firstSampleIndex)));
    if (fr.samplerType == OHB_SAMPLER_PCG) { cx.sm.pcg = m.z; cx.dimIdx = 0u; } else cx.dimIdx = m.z;
    f3 rad = xyz(P.rad[p]);
    if (cx.state & OHB_ST_PEND_A) rad += xyz(P.pendA[p]);
    if (cx.state & OHB_ST_PEND_B) rad += xyz(P.pendB[p]);
    cx.state &= ~(OHB_ST_PEND_A | OHB_ST_PEND_B);
    uint32_t stage = OHB_ST_STAGE(cx.state), bounce = OHB_ST_BOUNCE(cx.state);
    const bool primary = (stage == ST_PRIMARY);
    f3 d = xyz(P.rayD[p]);
    f4 q0, q1, q2, q3;
    if (FUSED) surfaceShade(sc, fr, P, p, q0, q1, q2, q3);
    else { q0 = P.pay0[p]; q1 = P.pay1[p]; }
    const bool isMiss = q0.w < 0.0f;
    const bool legacy = (fr.flags & OHB_FLAG_RESTIRGI_LEGACY) != 0u;

    f3 nextO = mk3(0.0f), nextD = mk3(0.0f), thr = mk3(1.0f); float lastPdf = 0.0f; bool lastDelta = false;
    if (!primary) { f4 t4 = P.thr[p]; thr = xyz(t4); lastPdf = t4.w; lastDelta = (cx.state & OHB_ST_DELTA) != 0u; }
    bool startGI = false, finished = false, chainEnds = false; uint32_t doneFlags = 0u;

    if (stage == ST_GI) {
        // ---- ReSTIR GI initial sample at x_s (:1377-1531); the RIS update itself happens in pixelRT, in sample order ----
        if (FUSED) { P.pay0[p] = q0; P.pay1[p] = q1; if (!isMiss) P.pay3[p] = q3; }      // what pixelRT reads of the GI sample
        if (isMiss) doneFlags = OHB_ST_GI_MISS;
        else {
            if (!FUSED) q2 = P.pay2[p];
            f3 xs = xyz(q0), ns = xyz(q1), sAlbedo = xyz(q2);
            float sPacked = q1.w;
            bool sIsMetal = sPacked < 0.0f;                                   // legacy encoding, always false today (quirk Q6)
            float sRough = fabsf(sPacked); if (sRough >= 10.0f) sRough -= 10.0f; sRough = fmaxf(sRough, 0.01f);
            f3 sF0 = sIsMetal ? sAlbedo : mk3(0.04f);
            float kd = sIsMetal ? 0.0f : 1.0f;
            if (sc.lightCount > 0u) lightNEE_RT(cx, xs, ns, d, sAlbedo, sF0, sRough, kd, false, true, mk3(1.0f));
            if (cx.envOn) envNEE_RT(cx, xs, ns, d, sAlbedo, sF0, sRough, 0.0f, kd, false, true, mk3(1.0f));
            doneFlags = OHB_ST_GI_HIT;
        }
        finished = true;
    } else if (isMiss) {
        f3 color = xyz(q0); float envPdf = q1.x;
        if (primary) {
            rad = color;
            P.fh2[p] = mk4(0.0f, 0.0f, 0.0f, -1.0f);                          // firstHitDist = -1
            if ((fr.flags & OHB_FLAG_ENABLE_AOVS) && m.y == P.firstSampleIndex && P.albedoAOV) {
                size_t pi = size_t(py) * fr.W + px; P.albedoAOV[pi] = mk4(color, 1.0f); P.normalAOV[pi] = mk4(0, 0, 0, 0);
            }
            finished = true;
        } else {
            float w = 1.0f;
            if (envPdf > 0.0f && fr.envW > 0u && !lastDelta) w = misBalance(lastPdf, envPdf);
            rad += thr * color * w;
            chainEnds = true;
        }
    } else {
        if (!FUSED) { q2 = P.pay2[p]; q3 = P.pay3[p]; }
        f3 hp = xyz(q0), N = xyz(q1), albedo = xyz(q2), em = xyz(q3);
        float rough, metal; unpackHitPbr(mk3(q1.w, q2.w, q3.w), rough, metal);
        f3 F0 = mix(mk3(0.04f), albedo, metal);
        if (primary && (fr.flags & OHB_FLAG_ENABLE_AOVS) && m.y == P.firstSampleIndex && P.albedoAOV) {
            size_t pi = size_t(py) * fr.W + px; P.albedoAOV[pi] = mk4(albedo, 1.0f); P.normalAOV[pi] = mk4(N * 0.5f + mk3(0.5f), 1.0f);
        }
        if (length(em) > 0.001f) rad += primary ? em : thr * em;
        if (sc.lightCount > 0u) lightNEE_RT(cx, hp, N, d, albedo, F0, rough, 1.0f - metal, primary, false, thr);
        if (cx.envOn) envNEE_RT(cx, hp, N, d, albedo, F0, rough, metal, 1.0f - metal, primary, false, thr);
        nextO = hp + N * 0.01f;
        if (primary) {
            // ---- Stage B set-up: GGX VNDF first glossy bounce (:722-801) ----
            f3 V = normalize(-d);
            float NdotV = fmaxf(dot(N, V), 1e-4f);
            float alpha = rough * rough;
            if (rough < 0.02f) {
                nextD = reflect(d, N);
                thr = F0 + (mk3(1.0f) - F0) * pow5(1.0f - NdotV);
                lastPdf = 1.0f; lastDelta = true;
            } else {
                f3 up = fabsf(N.y) < 0.999f ? mk3(0, 1, 0) : mk3(1, 0, 0);
                f3 T = normalize(cross(up, N)), B = cross(N, T);
                f3 Vloc = mk3(dot(V, T), dot(V, B), dot(V, N));
                f2 u = cx.sm.get2D(cx.dimIdx); cx.dimIdx += 2u;
                f3 Hloc = sampleGGXVNDF(Vloc, alpha, alpha, u);
                f3 Hh = normalize(Hloc.x * T + Hloc.y * B + Hloc.z * N);
                f3 refl = reflect(-V, Hh);
                if (dot(refl, N) <= 0.0f) { refl = reflect(d, N); Hh = N; }
                refl = normalize(refl);
                nextD = refl;
                float NdotL = fmaxf(dot(N, refl), 1e-4f), NdotH = fmaxf(dot(N, Hh), 1e-4f), VdotH = fmaxf(dot(V, Hh), 1e-4f);
                f3 F = F0 + (mk3(1.0f) - F0) * pow5(1.0f - VdotH);
                thr = F * smithG2overG1GGX(NdotV, NdotL, alpha);
                lastPdf = ediv(smithG1GGX(NdotV, alpha) * ggxDiso(NdotH, alpha), 4.0f * NdotV) + 1e-6f; lastDelta = false;
            }
            P.fh0[p] = mk4(hp, rough); P.fh1[p] = mk4(N, metal); P.fh2[p] = mk4(albedo, q0.w);
            if (fr.maxBounces >= 1u) { stage = ST_CHAIN_B; bounce = 1u; } else chainEnds = true;
        } else {
            bool killed = false;
            if (bounce > 1u) {   // Russian roulette (:1007-1013)
                float pr = clampf(maxcomp(thr), 0.1f, 0.95f);
                float rr = cx.sm.get1D(cx.dimIdx); cx.dimIdx += 1u;
                if (rr > pr) killed = true; else thr /= pr;
            }
            if (killed) chainEnds = true;
            else {
                float specProb = specProbOf(d, N, F0, rough, metal);
                float choice = cx.sm.get1D(cx.dimIdx); cx.dimIdx += 1u;
                if (choice < specProb || rough < 0.05f) {
                    nextD = sampleSpecDir(cx, d, N, rough);
                    thr *= mix(mk3(1.0f), albedo, metal);
                    thr /= fmaxf(specProb, 0.01f);
                    if (rough < 0.05f) { lastPdf = 1.0f; lastDelta = true; }
                    else {
                        f3 Hs = normalize(-d + nextD);
                        float NdotH = fmaxf(dot(N, Hs), 0.001f), VdotH = fmaxf(dot(-d, Hs), 0.001f);
                        float as = rough * rough, as2 = as * as; float dn = NdotH * NdotH * (as2 - 1.0f) + 1.0f;
                        float Ds = ediv(as2, OHB_PI * dn * dn + 1e-4f);
                        lastPdf = specProb * ediv(Ds * NdotH, 4.0f * VdotH + 1e-4f); lastDelta = false;
                    }
                } else {
                    f2 du = cx.sm.get2D(cx.dimIdx); cx.dimIdx += 2u;
                    nextD = cosineHemisphereShared(N, du);
                    thr *= albedo;
                    thr /= fmaxf(1.0f - specProb, 0.01f);
                    lastPdf = ediv((1.0f - specProb) * fmaxf(dot(nextD, N), 0.0f), OHB_PI); lastDelta = false;
                }
                if (bounce >= fr.maxBounces) chainEnds = true; else bounce += 1u;
            }
        }
    }
    if (chainEnds) { if (legacy) finished = true; else startGI = true; }
    if (startGI) {
        // cosine bounce from the FIRST hit (:1369-1375)
        f4 a0 = P.fh0[p], a1 = P.fh1[p];
        f3 fhPos = xyz(a0), fhN = xyz(a1);
        f2 du = cx.sm.get2D(cx.dimIdx); cx.dimIdx += 2u;
        nextD = cosineHemisphereShared(fhN, du);
        nextO = fhPos + fhN * 0.01f;
        thr = mk3(0.0f); lastPdf = fmaxf(dot(fhN, nextD), 0.0f);   // thr.w carries cosAtX1 to pixelRT
        lastDelta = false; stage = ST_GI; bounce = 1u;
    }
    P.rad[p] = mk4(rad, 0.0f);
    uint32_t keep = cx.state & (OHB_ST_PEND_A | OHB_ST_PEND_B);
    m.z = (fr.samplerType == OHB_SAMPLER_PCG) ? cx.sm.pcg : cx.dimIdx;
    if (finished) { m.w = ST_DONE | keep | doneFlags; P.meta[p] = m; return false; }
    m.w = OHB_ST_MAKE(stage, bounce) | keep | (lastDelta ? OHB_ST_DELTA : 0u);
    P.meta[p] = m;
    P.thr[p] = mk4(thr, lastPdf);
    P.rayO[p] = mk4(nextO, 0.0f); P.rayD[p] = mk4(nextD, 0.0f);
    return true;
}

// Per-frame images of the realtime profile (all RGBA32F, W*H).
struct RTImagesDev {
    const f4* accumPrev; f4* accumCurr;
    const f4* surfPrev; f4* surfCurr; const f4* shadPrev; f4* shadCurr;
    const f4* res0Prev; const f4* res1Prev; const f4* res2Prev; f4* res0Curr; f4* res1Curr; f4* res2Curr;
    float* radianceDump; float* giDump;     // optional parity dumps (W*H*4 floats)
    uint32_t* motionAOV; float* depthAOV;   // RG16F motion vectors / R32F linear view Z for the SVGF denoiser (null when it is off)
    unsigned long long* counters;
};
OHB_HD bool prevPixelOf(const FrameParams& fr, f3 p, int& qx, int& qy) {
    const float* M = fr.prevViewProj;   // column-major
    float cx = M[0] * p.x + M[4] * p.y + M[8] * p.z + M[12], cy = M[1] * p.x + M[5] * p.y + M[9] * p.z + M[13], cw = M[3] * p.x + M[7] * p.y + M[11] * p.z + M[15];
    if (!(cw > 0.0f)) return false;
    float ux = cx / cw * 0.5f + 0.5f, uy = 1.0f - (cy / cw * 0.5f + 0.5f);
    if (!(ux >= 0.0f && ux < 1.0f && uy >= 0.0f && uy < 1.0f)) return false;
    int ix = int(ux * float(fr.W)), iy = int(uy * float(fr.H));
    qx = ix < 0 ? 0 : (ix > int(fr.W) - 1 ? int(fr.W) - 1 : ix); qy = iy < 0 ? 0 : (iy > int(fr.H) - 1 ? int(fr.H) - 1 : iy);
    return true;
}
OHB_HD bool historyGate(f4 ps, f4 ph, f3 pos, f3 nrm, float dist, float rough, float posAbs, float posRel, float nThr, float rThr) {
    if (!(ps.w > 0.0f)) return false;
    bool ok = length(xyz(ps) - pos) <= fmaxf(posAbs, posRel * dist);
    if (ok && ph.w > 0.0f) ok = dot(normalize(xyz(ph)), normalize(nrm)) >= nThr && fabsf(ph.w - rough) <= rThr;
    return ok;
}

// k_rt_pixel: one thread per pixel, after the wavefront (:1533-1848).
OHB_HD void pixelRT(const SceneDev& sc, const FrameParams& fr, const PathArrays& P, const RTImagesDev& im, uint32_t pix) {
    u4 m0 = P.meta[pix];
    if (m0.x == 0xFFFFFFFFu) return;
    const uint32_t px = m0.x & 0xFFFFu, py = m0.x >> 16, N = P.samplesInBatch;
    const size_t pi = size_t(py) * fr.W + px;
    const bool clampOn = (fr.flags & OHB_FLAG_ENABLE_FIREFLY_CLAMP) && fr.fireflyClamp > 0.0f;
    const bool giOff = (fr.flags & OHB_FLAG_RESTIRGI_OFF) != 0u, giLegacy = (fr.flags & OHB_FLAG_RESTIRGI_LEGACY) != 0u, giNoSpatial = (fr.flags & OHB_FLAG_RESTIRGI_NOSPATIAL) != 0u;
    const bool viewChanged = fr.viewChanged != 0u;
    // first hit of the pixel (identical for all N samples: the camera ray is fixed); the reference keeps the last sample's
    const uint32_t pl = (N - 1u) * P.numPixels + pix;
    f4 a0 = P.fh0[pl], a1 = P.fh1[pl], a2 = P.fh2[pl];
    const float firstHitDist = a2.w;
    const bool hit = firstHitDist > 0.0f;
    const f3 firstHitPos = hit ? xyz(a0) : mk3(0.0f), firstHitNormal = hit ? xyz(a1) : mk3(0, 0, 1);
    const float firstHitRoughness = hit ? a0.w : 1.0f;
    const f3 giAlbedoD = hit ? xyz(a2) * (1.0f - a1.w) : mk3(0.0f);

    f3 radianceTotal = mk3(0.0f), giEnvMissTotal = mk3(0.0f);
    GIReservoir cur; cur.xs = mk3(0.0f); cur.ns = mk3(0, 0, 1); cur.Lo = mk3(0.0f); cur.wSum = 0.0f; cur.M = 0.0f; cur.W = 0.0f;
    Sampler sm; uint32_t dimIdx = 2u;
    for (uint32_t s = 0; s < N; s++) {
        uint32_t p = s * P.numPixels + pix;
        u4 m = P.meta[p];
        f3 rad = xyz(P.rad[p]);
        bool giHit = (m.w & OHB_ST_GI_HIT) != 0u;
        f3 pendSum = mk3(0.0f);
        if (m.w & OHB_ST_PEND_A) pendSum += xyz(P.pendA[p]);
        if (m.w & OHB_ST_PEND_B) pendSum += xyz(P.pendB[p]);
        sm.init(fr.samplerType, px, py, m.y, ldu4(P.sobolTab + (m.y - P.firstSampleIndex)));
        if (fr.samplerType == OHB_SAMPLER_PCG) { sm.pcg = m.z; dimIdx = 0u; } else dimIdx = m.z;
        if (giHit) {
            f3 xs = xyz(P.pay0[p]), ns = xyz(P.pay1[p]);
            f3 Lo = xyz(P.pay3[p]);
            if (m.w & OHB_ST_PEND_A) Lo += xyz(P.pendA[p]);
            if (m.w & OHB_ST_PEND_B) Lo += xyz(P.pendB[p]);
            float cosAtX1 = P.thr[p].w;
            float pHat = giTargetPHat(giAlbedoD, firstHitNormal, firstHitPos, xs, Lo);
            float p_i = cosAtX1 / OHB_PI;
            float w_i = (p_i > 0.0f) ? pHat / p_i : 0.0f;
            float rSel = sm.get1D(dimIdx); dimIdx += 1u;
            cur.wSum += w_i; cur.M += 1.0f;
            if (w_i > 0.0f && rSel * cur.wSum < w_i) { cur.xs = xs; cur.ns = ns; cur.Lo = Lo; }
        } else {
            rad += pendSum;
            if (m.w & OHB_ST_GI_MISS) giEnvMissTotal += giAlbedoD * xyz(P.pay0[p]);
        }
        radianceTotal += rad;
    }
    float invSamples = 1.0f / float(N);
    f3 radiance = radianceTotal * invSamples;
    if (clampOn) { float lum = luminance(radiance), cap = fr.fireflyClamp * 0.75f; if (lum > cap) radiance *= cap / lum; }
    if (im.radianceDump) { float* o = im.radianceDump + pi * 4u; o[0] = radiance.x; o[1] = radiance.y; o[2] = radiance.z; o[3] = 1.0f; }

    uint32_t visRays = 0u;
    f3 giDiffuse = mk3(0.0f);
    if (!giLegacy) {
        f3 giEnvMiss = giEnvMissTotal * invSamples;
        GIReservoir merged = cur;
        bool giReuse = !giOff && fr.historyCount > 0u && !viewChanged && hit && cur.M > 0.0f;
        int qx = 0, qy = 0;
        if (giReuse && prevPixelOf(fr, firstHitPos, qx, qy)) {
            size_t qi = size_t(qy) * fr.W + size_t(qx);
            if (historyGate(im.surfPrev[qi], im.shadPrev[qi], firstHitPos, firstHitNormal, firstHitDist, firstHitRoughness, 0.03f, 0.02f, 0.9f, 0.12f)) {
                f4 r0 = im.res0Prev[qi], r1 = im.res1Prev[qi], r2 = im.res2Prev[qi];
                if (r2.w > 0.5f && r0.w > 0.0f && r1.w > 0.0f) {
                    f3 pxs = xyz(r0), pLo = xyz(r2);
                    float pHatPrev = giTargetPHat(giAlbedoD, firstHitNormal, firstHitPos, pxs, pLo);
                    if (pHatPrev > 0.0f) {
                        f3 toXs = pxs - firstHitPos; float distXs = length(toXs);
                        if (distXs > 0.05f) { visRays++; if (traceAny(sc, firstHitPos + firstHitNormal * 0.01f, toXs / distXs, 0.001f, distXs - 0.02f)) pHatPrev = 0.0f; }
                    }
                    float mClamped = fminf(r0.w, 20.0f * cur.M);
                    float wPrev = pHatPrev * r1.w * mClamped;
                    merged.wSum += wPrev; merged.M += mClamped;
                    float rMerge = sm.get1D(dimIdx); dimIdx += 1u;
                    if (wPrev > 0.0f && rMerge * merged.wSum < wPrev) { merged.xs = pxs; merged.ns = xyz(r1); merged.Lo = pLo; }
                }
            }
        }
        float pHatHeld = giTargetPHat(giAlbedoD, firstHitNormal, firstHitPos, merged.xs, merged.Lo);
        merged.W = (pHatHeld > 0.0f && merged.M > 0.0f) ? merged.wSum / (merged.M * pHatHeld) : 0.0f;
        im.res0Curr[pi] = mk4(merged.xs, merged.M); im.res1Curr[pi] = mk4(merged.ns, merged.W); im.res2Curr[pi] = mk4(merged.Lo, hit ? 1.0f : 0.0f);

        GIReservoir sp = merged;
        bool spatialOn = !giOff && !giNoSpatial && fr.historyCount > 0u && !viewChanged && hit && merged.M > 0.0f;
        if (spatialOn) {
            int bx = int(px), by = int(py);
            { int tx, ty; if (prevPixelOf(fr, firstHitPos, tx, ty)) { bx = tx; by = ty; } }
            f3 n1r = normalize(firstHitNormal);
            for (int k = 0; k < 4; k++) {
                f2 du = sm.get2D(dimIdx); dimIdx += 2u;
                float rr = 20.0f * (1.0f - 0.15f * float(k)) * sqrtf(du.x);
                float sth, cth; ohb_sincos_turns(du.y, sth, cth);     // th = 6.2831853 * du.y
                int ox = int(nearbyintf(rr * cth)), oy = int(nearbyintf(rr * sth));
                if (ox == 0 && oy == 0) { ox = 1; oy = 0; }
                int sx = bx + ox, sy = by + oy;
                sx = sx < 0 ? 0 : (sx > int(fr.W) - 1 ? int(fr.W) - 1 : sx); sy = sy < 0 ? 0 : (sy > int(fr.H) - 1 ? int(fr.H) - 1 : sy);
                size_t qi = size_t(sy) * fr.W + size_t(sx);
                f4 qSurf = im.surfPrev[qi];
                if (qSurf.w <= 0.0f) continue;
                f3 x1q = xyz(qSurf);
                if (fabsf(dot(x1q - firstHitPos, n1r)) > fmaxf(0.05f, 0.1f * firstHitDist)) continue;
                f4 qShad = im.shadPrev[qi];
                if (qShad.w > 0.0f && dot(normalize(xyz(qShad)), n1r) < 0.9f) continue;
                f4 q0 = im.res0Prev[qi], q1 = im.res1Prev[qi], q2 = im.res2Prev[qi];
                if (q2.w <= 0.5f || q0.w <= 0.0f || q1.w <= 0.0f) continue;
                f3 qxs = xyz(q0), qns = xyz(q1), qLo = xyz(q2);
                float pHatR = giTargetPHat(giAlbedoD, firstHitNormal, firstHitPos, qxs, qLo);
                if (pHatR <= 0.0f) continue;
                float J = giSpatialJacobian(qns, qxs, firstHitPos, x1q);
                if (J <= 0.0f) continue;
                f3 toXs = qxs - firstHitPos; float dXs = length(toXs);
                if (dXs > 0.05f) { visRays++; if (traceAny(sc, firstHitPos + n1r * 0.01f, toXs / dXs, 0.001f, dXs - 0.02f)) continue; }
                float mNb = fminf(q0.w, 20.0f);
                float wNb = pHatR * q1.w * mNb * J;
                sp.wSum += wNb; sp.M += mNb;
                float rMerge = sm.get1D(dimIdx); dimIdx += 1u;
                if (wNb > 0.0f && rMerge * sp.wSum < wNb) { sp.xs = qxs; sp.ns = qns; sp.Lo = qLo; }
            }
        }
        float pHatS = giTargetPHat(giAlbedoD, firstHitNormal, firstHitPos, sp.xs, sp.Lo);
        sp.W = (pHatS > 0.0f && sp.M > 0.0f) ? sp.wSum / (sp.M * pHatS) : 0.0f;
        if (hit) giDiffuse = giShade(giAlbedoD, firstHitNormal, firstHitPos, sp.xs, sp.Lo, sp.W);
        giDiffuse += giEnvMiss;
        if (clampOn) clampLum(giDiffuse, fr.fireflyClamp);
        radiance += giDiffuse;
    }
    if (im.giDump) { float* o = im.giDump + pi * 4u; o[0] = giDiffuse.x; o[1] = giDiffuse.y; o[2] = giDiffuse.z; o[3] = 1.0f; }
    if (im.motionAOV) svgfGuides(fr.currViewProj, fr.prevViewProj, fr.viewRow2, fr.W, fr.H, fr.frameIdx, hit, firstHitPos, im.motionAOV[pi], im.depthAOV[pi]);
    im.surfCurr[pi] = hit ? mk4(firstHitPos, firstHitDist) : mk4(0.0f, 0.0f, 0.0f, -1.0f);
    im.shadCurr[pi] = hit ? mk4(firstHitNormal, firstHitRoughness) : mk4(0.0f, 0.0f, 1.0f, -1.0f);

    // reprojected EMA (:1773-1848), history from the previous frame's buffers
    f3 acc = radiance; float count = 1.0f;
    if (fr.historyCount > 0u) {
        bool useReprojection = false; int qx = int(px), qy = int(py);
        if (hit) { int tx, ty; if (prevPixelOf(fr, firstHitPos, tx, ty)) { useReprojection = true; qx = tx; qy = ty; } }
        size_t qi = size_t(qy) * fr.W + size_t(qx);
        f4 history = im.accumPrev[useReprojection ? qi : pi];
        bool historyValid = !viewChanged;
        if (useReprojection) historyValid = hit && historyGate(im.surfPrev[qi], im.shadPrev[qi], firstHitPos, firstHitNormal, firstHitDist, firstHitRoughness, 0.03f, 0.02f, 0.9f, 0.12f);
        else if (!viewChanged && hit) historyValid = historyGate(im.surfPrev[pi], im.shadPrev[pi], firstHitPos, firstHitNormal, firstHitDist, firstHitRoughness, 0.02f, 0.01f, 0.93f, 0.08f);
        float alpha = useReprojection ? 0.90f : 0.70f;
        if (viewChanged) alpha = useReprojection ? 0.72f : 0.45f;
        if (historyValid) {
            acc = mix(radiance, xyz(history), alpha);
            count = fminf(history.w + 1.0f, useReprojection ? (viewChanged ? 6.0f : 12.0f) : (viewChanged ? 2.0f : 4.0f));
        }
    }
    im.accumCurr[pi] = mk4(acc, count);
    if (visRays) count_add(im.counters + 2, visRays);
}

// k_rt_denoise: in-shader a-trous + tonemap (:1850-1925) over the finished frame.
// exp(-d2 / 4) for the squared tap distances of a 5x5 footprint (0, 1, 2, 4, 5, 8), correctly rounded
OHB_HD float atrousSpatialWeight(int d2) {
    return d2 == 0 ? 1.0f : (d2 == 1 ? 0.778800783f : (d2 == 2 ? 0.606530660f : (d2 == 4 ? 0.367879441f : (d2 == 5 ? 0.286504797f : 0.135335283f))));
}
// x^48 = x^32 * x^16 by squaring (six roundings; the shader's pow(normalSim, 48.0))
OHB_HD float pow48(float x) { float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8; return (x16 * x16) * x16; }
#ifndef OHB_RT_ATROUS_PASSES
#define OHB_RT_ATROUS_PASSES 1     // 3 = evaluate the two identity passes as well (A/B builds, tests/emul)
#endif
template <int PASSES = OHB_RT_ATROUS_PASSES, bool FASTW = true>
OHB_HD void denoiseRT(const FrameParams& fr, const f4* accum, const f4* normalAOV, uint32_t* ldr, float* denoisedDump, uint32_t pi) {
    const int W = int(fr.W), H = int(fr.H);
    const int px = int(pi % fr.W), py = int(pi / fr.W);
    f3 acc = xyz(accum[pi]);
    f3 den = acc;
    if ((fr.flags & OHB_FLAG_ENABLE_INTERNAL_DENOISE) && (fr.flags & OHB_FLAG_ENABLE_AOVS)) {
        f3 centerN = xyz(normalAOV[pi]);
        f3 mean = mk3(0.0f), meanSq = mk3(0.0f); int n = 0;
        for (int vy = -1; vy <= 1; vy++) for (int vx = -1; vx <= 1; vx++) {
            int x = px + vx, y = py + vy;
            if (x >= 0 && y >= 0 && x < W && y < H) { f3 s = xyz(accum[size_t(y) * W + x]); mean += s; meanSq += s * s; n++; }
        }
        mean /= float(n);
        f3 var = vmax(meanSq / float(n) - mean * mean, mk3(0.0f));
        float noise = dot(var, mk3(0.333f));
        if (noise > 0.00005f) {
            // Passes 1 and 2 of the shader (:1872-1899, steps 2 and 4) take the CENTRE value as every tap's colour
            // (`(pass == 0) ? imageLoad(accumBuffer, sp).rgb : denoised`): colorDiff = 0, wc = 1 and the pass returns
            // denoised * sum(w) / sum(w) — the identity up to ~25 roundings (or nothing at all when sum(w) <= 0.001).  The
            // oracle evaluates them literally; this kernel runs pass 0 only and skips their 50 normal taps per pixel
            // (k_rt_denoise 0.78 -> 0.32 ms per 1080p frame; table / squaring weights instead of expf / powf: -> 0.17 ms,
            // profiles/r2ae_sweep_rt_denoise.txt; tests/test_gpu_realtime.py bounds the deviation against the oracle).
            const float sigmaC = fmaxf(noise * 3.0f, 0.001f);
            for (int pass = 0; pass < PASSES; pass++) {
                int step = pass == 0 ? 1 : (pass == 1 ? 2 : 4);
                f3 sum = mk3(0.0f); float wSum = 0.0f;
                for (int dy = -2; dy <= 2; dy++) for (int dx = -2; dx <= 2; dx++) {
                    int x = px + dx * step, y = py + dy * step;
                    if (x < 0 || y < 0 || x >= W || y >= H) continue;
                    size_t si = size_t(y) * W + x;
                    f3 sc = pass == 0 ? xyz(accum[si]) : den;
                    f3 sn = xyz(normalAOV[si]);
                    float ws = FASTW ? atrousSpatialWeight(dx * dx + dy * dy) : expf(-float(dx * dx + dy * dy) / 4.0f);
                    float wn = FASTW ? pow48(fmaxf(dot(centerN, sn), 0.0f)) : ohb_pow(fmaxf(dot(centerN, sn), 0.0f), 48.0f);
                    f3 cd = den - sc;
                    float wc = expf(ediv(-dot(cd, cd), sigmaC + 0.0001f));
                    float w = ws * wn * wc;
                    sum += sc * w; wSum += w;
                }
                if (wSum > 0.001f) den = sum / wSum;
            }
        }
        float accLum = luminance(acc), denLum = luminance(den);
        if (denLum > 0.001f && accLum > 0.001f) den *= accLum / denLum;
    }
    if (denoisedDump) { float* o = denoisedDump + size_t(pi) * 4u; o[0] = den.x; o[1] = den.y; o[2] = den.z; o[3] = 1.0f; }
    ldr[pi] = tonemapRGBA8(den);
}

}  // namespace ohb

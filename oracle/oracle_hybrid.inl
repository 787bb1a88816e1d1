// oracle_hybrid.inl — CPU restatement of the hybrid-RT raygen shaders (TEST INFRASTRUCTURE, included by oracle.cpp):
//   shaders/rt/rt_shadow.rgen:25-111 (+ rt_shadow.rmiss: payload 1 on miss), shaders/rt/rt_gi.rgen:30-134, rt_gi.rchit:16-28,
//   rt_gi.rmiss, includes/common/encoding.glsl:15-22.  G-buffer texture() fetches happen at texel centres of same-size images,
//   i.e. they return the texel.  Parity unpinned against the reference (no test vectors exist for the techniques).
static inline float fractF(float x) { return x - std::floor(x); }
static float hyHash(float px, float py) {                                   // hash(), rt_shadow.rgen:25-29
    V3 p3{fractF(px * 0.1031f), fractF(py * 0.1031f), fractF(px * 0.1031f)};
    float d = dot(p3, V3{p3.y + 33.33f, p3.z + 33.33f, p3.x + 33.33f});
    p3 = V3{p3.x + d, p3.y + d, p3.z + d};
    return fractF((p3.x + p3.y) * p3.z);
}
static V3 hyDecodeOct(float ex, float ey) {                                 // decodeNormalOctahedron, encoding.glsl:15-22
    float fx = ex * 2.0f - 1.0f, fy = ey * 2.0f - 1.0f;
    V3 n{fx, fy, 1.0f - std::fabs(fx) - std::fabs(fy)};
    float t = clampf(-n.z, 0.0f, 1.0f);
    n.x += (n.x >= 0.0f) ? -t : t;
    n.y += (n.y >= 0.0f) ? -t : t;
    return normalize(n);
}
static void hyBasis(V3 N, V3& T, V3& B) {                                   // buildBasis, rt_shadow.rgen:32-36
    V3 up = std::fabs(N.y) < 0.999f ? V3{0, 1, 0} : V3{1, 0, 0};
    T = normalize(cross(up, N)); B = cross(N, T);
}
struct orc_hybrid_shadow_params { float light_dir[3]; float light_radius; float light_pos[3]; float light_range; uint32_t light_type, sample_count, _pad[2]; };
struct orc_hybrid_gi_params { float light_pos[3]; float light_intensity; uint32_t sample_count, frame_index, _pad[2]; };

static void hybridShadow(const Scene& sc, uint32_t W, uint32_t H, const float* gPos, const float* gNrm, const orc_hybrid_shadow_params& pc, uint8_t* mask, int nthreads) {
    parallelRows(H, nthreads, [&](uint32_t y) {
        for (uint32_t x = 0; x < W; x++) {
            size_t pi = size_t(y) * W + x;
            const float* ps = gPos + pi * 4;
            if (ps[0] == 0.0f && ps[1] == 0.0f && ps[2] == 0.0f && ps[3] == 0.0f) { mask[pi] = 255; continue; }          // :49-53
            V3 worldPos{ps[0], ps[1], ps[2]}, N = hyDecodeOct(gNrm[pi * 2], gNrm[pi * 2 + 1]);
            V3 origin = worldPos + N * 0.05f;                                                                              // :60
            uint32_t sampleCount = std::max(pc.sample_count, 1u);
            float visibility = 0.0f;
            for (uint32_t s = 0; s < sampleCount; s++) {
                float fs = float(s);
                float r1 = hyHash(float(x) + fs * 7.13f, float(y) + fs * 13.37f), r2 = hyHash(float(x) + fs * 31.17f, float(y) + fs * 47.53f);   // :70-71
                V3 L; float tMax;
                if (pc.light_type == 0u) {                                                                                 // :76-85
                    V3 lightDir = normalize(V3{-pc.light_dir[0], -pc.light_dir[1], -pc.light_dir[2]}), T, B; hyBasis(lightDir, T, B);
                    float angle = pc.light_radius * std::sqrt(r1), phi = 6.2831853f * r2;
                    L = normalize(lightDir + T * (angle * std::cos(phi)) + B * (angle * std::sin(phi)));
                    tMax = 10000.0f;
                } else {                                                                                                   // :86-98
                    float theta = 6.2831853f * r1, phi = std::acos(1.0f - 2.0f * r2);
                    V3 offset = V3{std::sin(phi) * std::cos(theta), std::sin(phi) * std::sin(theta), std::cos(phi)} * pc.light_radius;
                    V3 toLight = (V3{pc.light_pos[0], pc.light_pos[1], pc.light_pos[2]} + offset) - worldPos;
                    float dist = length(toLight);
                    L = toLight / dist; tMax = dist;
                }
                if (dot(N, L) <= 0.0f) continue;                                                                           // :101
                if (!traceAny(sc, origin, L, 0.001f, tMax)) visibility += 1.0f;                                            // :103-108
            }
            visibility /= float(sampleCount);
            mask[pi] = uint8_t(std::lrintf(clampf(visibility, 0.0f, 1.0f) * 255.0f));                                      // R8_UNORM store
        }
    });
}
static V3 hyCosineHemisphere(float ux, float uy, V3 N) {                                                                  // rt_gi.rgen:41-56
    V3 T, B; hyBasis(N, T, B);
    float r = std::sqrt(ux), phi = 6.2831853f * uy;
    float cx = r * std::cos(phi), cy = r * std::sin(phi), cz = std::sqrt(std::max(0.0f, 1.0f - ux));
    return normalize(T * cx + B * cy + N * cz);
}
static void hybridGi(const Scene& sc, uint32_t W, uint32_t H, const float* gPos, const float* gNrm, const float* gAlbedo, const float* history, const float* instMat,
                     const orc_hybrid_gi_params& pc, uint16_t* out, int nthreads) {
    parallelRows(H, nthreads, [&](uint32_t y) {
        for (uint32_t x = 0; x < W; x++) {
            size_t pi = size_t(y) * W + x;
            const float* ps = gPos + pi * 4; uint16_t* o = out + pi * 4;
            if (ps[0] == 0.0f && ps[1] == 0.0f && ps[2] == 0.0f && ps[3] == 0.0f) { o[0] = o[1] = o[2] = o[3] = 0; continue; }   // :67-71
            V3 worldPos{ps[0], ps[1], ps[2]}, N = hyDecodeOct(gNrm[pi * 2], gNrm[pi * 2 + 1]), albedo{gAlbedo[pi * 4], gAlbedo[pi * 4 + 1], gAlbedo[pi * 4 + 2]};
            V3 origin = worldPos + N * 0.05f;
            uint32_t sampleCount = std::max(pc.sample_count, 1u);
            V3 indirect{0, 0, 0};
            for (uint32_t s = 0; s < sampleCount; s++) {
                float fs = float(s), hx = float(x) + (fs * 7.13f + float(pc.frame_index) * 1.618f), hy = float(y) + fs * 13.37f;   // :88
                V3 dir = hyCosineHemisphere(hyHash(hx, hy), hyHash(hx + 127.1f, hy + 311.7f), N);
                ohb_hit h = traceClosest(sc, origin, dir, 0.01f, 100.0f);                                                   // :94-99
                if (h.prim == OHB_MISS) continue;
                const float* m = instMat + size_t(sc.triInst[h.prim]) * 4;                                                  // rt_gi.rchit:17
                if (m[3] < 0.5f) continue;                                                                                  // rt_gi.rchit:22-25
                V3 hitPos = origin + dir * h.t;
                V3 toLight = V3{pc.light_pos[0], pc.light_pos[1], pc.light_pos[2]} - hitPos;
                float lightDist = length(toLight), falloff = pc.light_intensity / (1.0f + lightDist * lightDist);
                float hitNdotL = std::max(dot(normalize(toLight), -dir), 0.0f);
                indirect += V3{m[0], m[1], m[2]} * falloff * hitNdotL;                                                      // :118
            }
            indirect = indirect / float(sampleCount);
            indirect = indirect * albedo;
            V3 hist{history[pi * 4], history[pi * 4 + 1], history[pi * 4 + 2]};
            float blend = pc.frame_index == 0u ? 1.0f : 0.3f;
            V3 acc = hist * (1.0f - blend) + indirect * blend;                                                              // mix(), :131
            o[0] = halfBits(acc.x); o[1] = halfBits(acc.y); o[2] = halfBits(acc.z); o[3] = halfBits(1.0f);                  // RGBA16F store
        }
    });
}

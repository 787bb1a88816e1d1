// oracle/oracle_svgf.inl — TEST INFRASTRUCTURE ONLY (see oracle.cpp header; included from it inside namespace orc).
//
// CPU restatement of the SVGF denoiser behind DenoiseMode::Atrous (OHAO @ c19e0d4):
//   * the guide AOVs the realtime raygen writes for it — pixel-space motion vectors (RG16F, binding 19) and linear
//     view Z (R32F, binding 20): shaders/rt/pt_raygen_realtime.rgen:412-455;
//   * pass 1, shaders/rt/rt_svgf_temporal.comp:61-126; pass 2, shaders/rt/rt_svgf_atrous.comp:54-115;
//   * the dispatch schedule, history ping-pong and sigmas of ohao/render/rt/denoise/atrous_denoise.cpp:28-40,392-537.
// Images the reference stores as RGBA16F / R16F / RG16F are kept as raw fp16 bits and converted with the compiler's
// _Float16 (IEEE round-to-nearest-even), RGBA8 loads are c / 255, RGBA8 stores round to nearest.
//
// PARITY UNPINNED against the reference: it holds no test vectors for this denoiser (tests/ has none for
// atrous_denoise.cpp) and the shaders only run inside a Vulkan compute pipeline.  Pinned only to the extent that
// the arithmetic below is a line-by-line restatement; the CUDA kernels are checked against THIS.

static inline uint16_t halfBits(float f) { _Float16 h = (_Float16)f; uint16_t u; std::memcpy(&u, &h, 2); return u; }
static inline float halfToFloat(uint16_t u) { _Float16 h; std::memcpy(&h, &u, 2); return (float)h; }
struct H4 { uint16_t v[4]; };
static inline V4 ldH4(const uint16_t* img, int W, int x, int y) { const uint16_t* p = img + (size_t(y) * W + x) * 4; return {halfToFloat(p[0]), halfToFloat(p[1]), halfToFloat(p[2]), halfToFloat(p[3])}; }
static inline void stH4(uint16_t* img, int W, int x, int y, V4 c) { uint16_t* p = img + (size_t(y) * W + x) * 4; p[0] = halfBits(c.x); p[1] = halfBits(c.y); p[2] = halfBits(c.z); p[3] = halfBits(c.w); }
static inline V3 ldRGBA8(const uint8_t* img, int W, int x, int y) { const uint8_t* p = img + (size_t(y) * W + x) * 4; return {float(p[0]) / 255.0f, float(p[1]) / 255.0f, float(p[2]) / 255.0f}; }
static inline void stRGBA8(uint8_t* img, int W, int x, int y, V3 c) {
    uint8_t* p = img + (size_t(y) * W + x) * 4;
    p[0] = uint8_t(std::nearbyint(clampf(c.x, 0.0f, 1.0f) * 255.0f)); p[1] = uint8_t(std::nearbyint(clampf(c.y, 0.0f, 1.0f) * 255.0f));
    p[2] = uint8_t(std::nearbyint(clampf(c.z, 0.0f, 1.0f) * 255.0f)); p[3] = 255;
}

// pt_raygen_realtime.rgen:412-455.  surf = this frame's surface history plane (firstHitPos, firstHitDist; w <= 0: miss).
static void svgfGuidesPixel(const M4& currViewProj, const M4& prevViewProj, const M4& viewMat, uint32_t W, uint32_t H, uint32_t frameIdx,
                            V4 surf, uint32_t& motionOut, float& depthOut) {
    V3 firstHitPos{surf.x, surf.y, surf.z}; float firstHitDist = surf.w;
    float mx = 0.0f, my = 0.0f;
    if (firstHitDist > 0.0f && frameIdx > 0u) {
        V4 currClip = mulv(currViewProj, V4{firstHitPos.x, firstHitPos.y, firstHitPos.z, 1.0f});
        V4 prevClip = mulv(prevViewProj, V4{firstHitPos.x, firstHitPos.y, firstHitPos.z, 1.0f});
        if (currClip.w > 0.0f && prevClip.w > 0.0f) {
            float cnx = currClip.x / currClip.w, cny = currClip.y / currClip.w, pnx = prevClip.x / prevClip.w, pny = prevClip.y / prevClip.w;
            float cpx = (cnx * 0.5f + 0.5f) * float(W), cpy = (cny * 0.5f + 0.5f) * float(H);
            float ppx = (pnx * 0.5f + 0.5f) * float(W), ppy = (pny * 0.5f + 0.5f) * float(H);
            mx = cpx - ppx; my = cpy - ppy;
        }
    }
    motionOut = uint32_t(halfBits(mx)) | (uint32_t(halfBits(my)) << 16);
    float firstHitViewZ = 1e30f;
    if (firstHitDist > 0.0f) { V4 viewPos = mulv(viewMat, V4{firstHitPos.x, firstHitPos.y, firstHitPos.z, 1.0f}); firstHitViewZ = -viewPos.z; }
    depthOut = firstHitViewZ;
}

struct SvgfImages {
    int W, H;
    uint8_t* beauty;                 // RGBA8, in and final out
    const uint32_t* motion; const float* depth; const float* normal;   // RG16F bits, R32F, RGBA32F (N*0.5+0.5)
    const uint16_t* prevColor; const uint16_t* prevMoments; const uint16_t* prevGeom;      // RGBA16F history of the previous frame
    uint16_t* curColor; uint16_t* curMoments; uint16_t* curGeom;                           // history written by this frame
};

// rt_svgf_temporal.comp:61-126
static void svgfTemporalPixel(const SvgfImages& im, int reset, int x, int y, uint16_t* outColor, uint16_t* outVariance) {
    const int W = im.W, H = im.H;
    V3 cur = ldRGBA8(im.beauty, W, x, y);
    float curL = luminance(cur);
    float curZ = im.depth[size_t(y) * W + x];
    const float* np = im.normal + (size_t(y) * W + x) * 4;
    V3 curN = V3{np[0], np[1], np[2]} * 2.0f - v3(1.0f);
    curN = dot(curN, curN) > 1e-8f ? normalize(curN) : V3{0.0f, 0.0f, 1.0f};
    float curZc = std::min(curZ, 1.0e4f);
    uint32_t mvBits = im.motion[size_t(y) * W + x];
    float mvx = halfToFloat(uint16_t(mvBits & 0xFFFFu)), mvy = halfToFloat(uint16_t(mvBits >> 16));
    float prevPixX = float(x) - mvx, prevPixY = float(y) - mvy;
    int prevX = int(std::floor(prevPixX + 0.5f)), prevY = int(std::floor(prevPixY + 0.5f));
    bool valid = reset == 0;
    if (prevX < 0 || prevY < 0 || prevX >= W || prevY >= H) valid = false;
    if (valid) {
        V4 g = ldH4(im.prevGeom, W, prevX, prevY);
        float prevZ = g.x; V3 prevN{g.y, g.z, g.w};
        float zrel = std::fabs(curZc - prevZ) / std::max(std::max(std::fabs(curZc), std::fabs(prevZ)), 1e-3f);
        float ndot = dot(curN, prevN);
        if (zrel > 0.1f || ndot < 0.9f) valid = false;
    }
    float prevLen = 0.0f; V3 prevCol = cur; float prevM1 = curL, prevM2 = curL * curL;
    if (valid) {
        V4 hc = ldH4(im.prevColor, W, prevX, prevY), hm = ldH4(im.prevMoments, W, prevX, prevY);
        prevCol = V3{hc.x, hc.y, hc.z}; prevM1 = hm.x; prevM2 = hm.y; prevLen = hm.z;
    }
    float newLen = valid ? std::min(prevLen + 1.0f, 32.0f) : 1.0f;
    float alpha = std::max(1.0f / newLen, 0.05f);
    V3 accumCol = mix(prevCol, cur, alpha);
    float m1 = mixf(prevM1, curL, alpha), m2 = mixf(prevM2, curL * curL, alpha);
    float variance = std::max(0.0f, m2 - m1 * m1);
    if (newLen < 4.0f) {
        float s = 0.0f, s2 = 0.0f; int cnt = 0;
        for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
            int qx = std::min(std::max(x + dx, 0), W - 1), qy = std::min(std::max(y + dy, 0), H - 1);
            float lq = luminance(ldRGBA8(im.beauty, W, qx, qy));
            s += lq; s2 += lq * lq; ++cnt;
        }
        float mm = s / float(cnt);
        variance = std::max(variance, std::max(0.0f, s2 / float(cnt) - mm * mm));
    }
    stH4(outColor, W, x, y, V4{accumCol.x, accumCol.y, accumCol.z, 1.0f});
    stH4(im.curMoments, W, x, y, V4{m1, m2, newLen, 0.0f});
    outVariance[size_t(y) * W + x] = halfBits(variance);
    stH4(im.curGeom, W, x, y, V4{curZc, curN.x, curN.y, curN.z});
}

// rt_svgf_atrous.comp:54-115
static void svgfAtrousPixel(const SvgfImages& im, int x, int y, int stepSize, float sigmaL, float sigmaNormal, float sigmaDepth, bool isFinal,
                            const uint16_t* inColor, uint16_t* outColor16, const uint16_t* inVar, uint16_t* outVar) {
    const int W = im.W, H = im.H;
    static const float kernel[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
    static const float gk[3] = {0.25f, 0.5f, 0.25f};
    auto nrm = [&](int qx, int qy) { const float* p = im.normal + (size_t(qy) * W + qx) * 4; return V3{p[0], p[1], p[2]} * 2.0f - v3(1.0f); };
    V4 c4 = ldH4(inColor, W, x, y); V3 cColor{c4.x, c4.y, c4.z};
    V3 cN = nrm(x, y);
    float cD = im.depth[size_t(y) * W + x], cL = luminance(cColor);
    float gVar = 0.0f, gW = 0.0f;
    for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
        int qx = std::min(std::max(x + dx, 0), W - 1), qy = std::min(std::max(y + dy, 0), H - 1);
        float w = gk[dx + 1] * gk[dy + 1];
        gVar += w * halfToFloat(inVar[size_t(qy) * W + qx]); gW += w;
    }
    float centerVar = gW > 0.0f ? gVar / gW : halfToFloat(inVar[size_t(y) * W + x]);
    float sqrtVar = std::sqrt(std::max(centerVar, 1e-8f));
    V3 sum = v3(0.0f); float weightSum = 0.0f, varSum = 0.0f;
    for (int dy = -2; dy <= 2; ++dy) for (int dx = -2; dx <= 2; ++dx) {
        int qx = std::min(std::max(x + dx * stepSize, 0), W - 1), qy = std::min(std::max(y + dy * stepSize, 0), H - 1);
        V4 s4 = ldH4(inColor, W, qx, qy); V3 sColor{s4.x, s4.y, s4.z};
        V3 sN = nrm(qx, qy);
        float sD = im.depth[size_t(qy) * W + qx], sVar = halfToFloat(inVar[size_t(qy) * W + qx]);
        float w = kernel[dx + 2] * kernel[dy + 2];
        float sL = luminance(sColor);
        w *= std::exp(-std::fabs(cL - sL) / (sqrtVar * sigmaL + 1e-6f));
        float normalDist = std::max(1.0f - dot(cN, sN), 0.0f);
        w *= std::exp(-normalDist / (sigmaNormal * sigmaNormal + 1e-4f));
        float depthDist = std::fabs(cD - sD);
        w *= std::exp(-depthDist / (sigmaDepth * sigmaDepth + 1e-4f));
        sum += sColor * w; weightSum += w; varSum += w * w * sVar;
    }
    V3 outC = weightSum > 1e-6f ? V3{sum.x / weightSum, sum.y / weightSum, sum.z / weightSum} : cColor;
    float outV = weightSum > 1e-6f ? varSum / (weightSum * weightSum) : centerVar;
    if (isFinal) stRGBA8(im.beauty, W, x, y, outC);
    else stH4(outColor16, W, x, y, V4{outC.x, outC.y, outC.z, 1.0f});
    outVar[size_t(y) * W + x] = halfBits(outV);
}

// AtrousDenoiser::dispatch (atrous_denoise.cpp:392-537): temporal pass, then 5 a-trous iterations (step 1,2,4,8,16);
// iteration 0 writes next frame's colour history, the last one the RGBA8 beauty.
static void svgfDispatch(const SvgfImages& im, int reset, float sigmaL, float sigmaNormal, float sigmaDepth, int nthreads) {
    const size_t n = size_t(im.W) * im.H;
    std::vector<uint16_t> colorA(n * 4), colorB(n * 4), varA(n), varB(n);
    auto rows = [&](auto&& fn) {
        std::atomic<int> next{0};
        auto work = [&]() { for (;;) { int y = next.fetch_add(1); if (y >= im.H) break; for (int x = 0; x < im.W; x++) fn(x, y); } };
        std::vector<std::thread> th; for (int t = 1; t < std::max(1, nthreads); t++) th.emplace_back(work);
        work(); for (auto& t : th) t.join();
    };
    rows([&](int x, int y) { svgfTemporalPixel(im, reset, x, y, colorA.data(), varA.data()); });
    // the final iteration reads the beauty image nowhere (inputs are the fp16 colour planes), so writing it in place is safe
    uint16_t* A = colorA.data(); uint16_t* B = colorB.data(); uint16_t* HC = im.curColor;
    const uint16_t* inColor[5] = {A, HC, B, A, B}; uint16_t* outColor[5] = {HC, B, A, B, A};
    const uint16_t* inVar[5] = {varA.data(), varB.data(), varA.data(), varB.data(), varA.data()};
    uint16_t* outVar[5] = {varB.data(), varA.data(), varB.data(), varA.data(), varB.data()};
    for (int it = 0; it < 5; it++)
        rows([&](int x, int y) { svgfAtrousPixel(im, x, y, 1 << it, sigmaL, sigmaNormal, sigmaDepth, it == 4, inColor[it], outColor[it], inVar[it], outVar[it]); });
}

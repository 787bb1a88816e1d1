// tests/emul/emul.cpp — HOST EMULATOR OF THE CUDA KERNELS.  TEST TOOLING ONLY.
//
// The build container has no GPU, so the per-thread code of the product kernels
// (ohao_engine_b200/csrc/ohb_{bvh,traverse,integrator}.h — the very same source nvcc compiles for
// sm_100a) is compiled here with g++ and driven by single-threaded loops that mimic the kernel
// launch sequence of ohb_kernels.cu.  This lets `pytest -m "not gpu"` check the LBVH builder, the
// traversal and the wavefront state machine against the oracle before any GPU minute is spent.
// It is NOT a fallback: nothing in ohao_engine_b200/ links or loads this file, it is slow, and the
// radix sort / scan / queue kernels (device-only code) are replaced by std::stable_sort here.
// Host stand-in for the warp's active-lane count: with EMUL_WARP_NOISE set the traversal sees a pseudo-random
// number of active lanes, which drives the triangle-postponing and pause/resume paths that a single host
// thread would otherwise never take.
static unsigned g_emulWarpNoise = 0, g_emulWarpState = 12345u;
static inline int emulWarpActive() {
    if (!g_emulWarpNoise) return 32;
    g_emulWarpState = g_emulWarpState * 1664525u + 1013904223u;
    return int(g_emulWarpState >> 27) + 1;
}
#define OHB_WARP_ACTIVE() emulWarpActive()
static unsigned long long g_statNodes = 0, g_statTris = 0;
#define OHB_STAT_NODE() (g_statNodes++)
#define OHB_STAT_TRI() (g_statTris++)
#include "../../ohao_engine_b200/csrc/ohb_bvh.h"
#include "../../ohao_engine_b200/csrc/ohb_integrator.h"
#include "../../ohao_engine_b200/csrc/ohb_realtime.h"
#include "../../ohao_engine_b200/csrc/ohb_hybrid.h"
#include <vector>
#include <algorithm>
#include <numeric>
#include <cstring>
#include <cmath>
#include <cstdlib>

using namespace ohb;

struct EmulScene {
    std::vector<uint8_t> positions; uint64_t stride = 0; uint32_t nverts = 0, ntris = 0;
    std::vector<uint32_t> indices, matIds, triInst, active;
    std::vector<f4> normals, matColors, instXform, instNormalMat, instInv, env;
    std::vector<f2> uvs;
    std::vector<uint8_t> tex; uint32_t texW = 0, texH = 0, texLayers = 0;
    std::vector<GPULight> lights; uint32_t lightCount = 0, envMapTexIdx = 0xFFFFFFFFu; float envIntensity = 1.0f;
    uint32_t envW = 0, envH = 0; std::vector<float> marg, cond, margTop, condTop; float envIntegral = 0;
    std::vector<u4> wnodes; std::vector<f4> tris; uint32_t numActive = 0, numNodes = 0, levels = 0; float sah = 0;
    // two-level structure + what MODE_UPDATE needs
    std::vector<ohb_instance> instances; int accelMode = 0;
    struct Range { uint32_t inst, first, count; }; std::vector<Range> ranges;
    std::vector<u4> tlasNodes, blasInfo; std::vector<f4> tlasLeaves, blasLo, blasHi; std::vector<uint32_t> instOfPrim;
    struct HostBuilder* mainB = nullptr; struct HostBuilder* tlasB = nullptr;
    SceneDev dev() const {
        SceneDev s; memset(&s, 0, sizeof(s));
        s.wnodes = wnodes.data(); s.tris = tris.data(); s.numTris = numActive; s.numWideNodes = accelMode ? 0u : numNodes;
        s.twoLevel = accelMode ? 1u : 0u; s.tlasNodes = tlasNodes.data(); s.tlasLeaves = tlasLeaves.data(); s.blasInfo = blasInfo.data();
        s.indices = indices.data(); s.normals = normals.data(); s.uvs = uvs.data(); s.matIds = matIds.data(); s.triInst = triInst.data();
        s.instNormalMat = instNormalMat.data(); s.instInv = instInv.data(); s.matColors = matColors.data();
        s.tex = tex.data(); s.texW = texW; s.texH = texH; s.texLayers = texLayers;
        s.lights = lights.data(); s.lightCount = lightCount; s.envMapTexIdx = envW ? envMapTexIdx : 0xFFFFFFFFu; s.envIntensity = envIntensity;
        s.env = envW ? env.data() : nullptr; s.envW = envW; s.envH = envH; s.marg = marg.data(); s.cond = cond.data(); s.envIntegral = envIntegral;
        s.margTop = condTop.empty() ? nullptr : margTop.data(); s.condTop = condTop.empty() ? nullptr : condTop.data();
        return s;
    }
};


// ---- host stand-in for the builder's launch sequences (ohb_kernels.cu launchBuild / launchRefit / launchBuildBlas / launchBuildTlas) ----
struct HostBuilder {
    std::vector<f4> wtri, primLo, primHi, nodeLo, nodeHi; std::vector<uint32_t> bounds, vals, visit, wideCounters; std::vector<uint64_t> keys;
    std::vector<int32_t> left, right, pin, pleaf; std::vector<float> sah; std::vector<WideItem> qa, qb;
    BuildArrays arrays(uint32_t n) {
        wtri.resize(size_t(n) * 3); primLo.resize(n); primHi.resize(n); nodeLo.resize(n); nodeHi.resize(n); bounds.assign(6, 0); vals.resize(n); visit.assign(n, 0); wideCounters.assign(4, 0);
        keys.resize(n); left.resize(n); right.resize(n); pin.assign(n, -1); pleaf.assign(n, -1); sah.assign(2, 0.0f); qa.resize(n / 4 + 2); qb.resize(n / 4 + 2);
        return view(n);
    }
    BuildArrays view(uint32_t n) {
        BuildArrays b{}; b.n = n; b.wtri = wtri.data(); b.primLo = primLo.data(); b.primHi = primHi.data(); b.boundsBits = bounds.data();
        b.keys = keys.data(); b.vals = vals.data(); b.left = left.data(); b.right = right.data(); b.parentInner = pin.data(); b.parentLeaf = pleaf.data();
        b.nodeLo = nodeLo.data(); b.nodeHi = nodeHi.data(); b.visit = visit.data(); b.wideCounters = wideCounters.data(); b.sah = sah.data();
        return b;
    }
    void initBuild(BuildArrays& b) { std::fill(visit.begin(), visit.end(), 0u); bounds[0] = bounds[1] = bounds[2] = 0xFFFFFFFFu; bounds[3] = bounds[4] = bounds[5] = 0u; sah[0] = sah[1] = 0.0f; (void)b; }
    void collapse(BuildArrays& b) {                                   // launchCollapse
        qa[0] = WideItem{0, 0u, 0u, 0u}; wideCounters[0] = 1; wideCounters[1] = 1; wideCounters[2] = 0; wideCounters[3] = 0;
        WideItem* in = qa.data(); WideItem* out = qb.data(); uint32_t* cin = &wideCounters[1]; uint32_t* cout = &wideCounters[2];
        for (int level = 0; level <= OHB_MAX_LEVELS; level++) {
            for (uint32_t i = 0; i < *cin; i++) emitWideNode(b, in[i], out, cout);
            *cin = 0; if (*cout) wideCounters[3]++;
            std::swap(in, out); std::swap(cin, cout);
        }
    }
    void fromPrims(BuildArrays& b) {                                  // launchBuildFromPrims
        const uint32_t n = b.n;
        for (uint32_t i = 0; i < n; i++) buildMorton(b, i);
        {   // stand-in for radixSort64 (device-only): stable sort by key
            std::vector<uint32_t> order(n); std::iota(order.begin(), order.end(), 0u);
            std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return keys[x] < keys[y]; });
            std::vector<uint64_t> k2(n); std::vector<uint32_t> v2(n);
            for (uint32_t i = 0; i < n; i++) { k2[i] = keys[order[i]]; v2[i] = vals[order[i]]; }
            std::copy(k2.begin(), k2.end(), keys.begin()); std::copy(v2.begin(), v2.end(), vals.begin());
        }
        if (n >= 2) {
            for (uint32_t i = 0; i + 1 < n; i++) buildHierarchyNode(b, int(i));
            for (uint32_t i = 0; i < n; i++) sweepFromLeaf(b, i, 0u);
            const char* tp = getenv("OHB_TREELET_PASSES"); uint32_t passes = tp ? uint32_t(atoi(tp)) : 3u;
            for (uint32_t pass = 0, gamma = OHB_TREELET_LEAVES; pass < passes; pass++, gamma *= 2u) {
                std::fill(visit.begin(), visit.end(), 0u);
                for (uint32_t i = 0; i < n; i++) sweepFromLeaf(b, i, gamma);
            }
        }
        collapse(b);
    }
    void build(BuildArrays& b) { initBuild(b); for (uint32_t i = 0; i < b.n; i++) buildWorldTri(b, i); fromPrims(b); }                 // launchBuild
    void refit(BuildArrays& b) {                                                                                                       // launchRefit
        initBuild(b); for (uint32_t i = 0; i < b.n; i++) buildWorldTri(b, i);
        if (b.n >= 2) for (uint32_t i = 0; i < b.n; i++) sweepFromLeaf(b, i, 0u);
        collapse(b);
    }
    void tlas(BuildArrays& b, const f4* lo, const f4* hi, const uint32_t* iop, bool refitOnly) {                                       // launchBuildTlas
        initBuild(b); for (uint32_t i = 0; i < b.n; i++) buildTlasPrim(b, lo, hi, iop, i);
        if (!refitOnly) { fromPrims(b); return; }
        if (b.n >= 2) for (uint32_t i = 0; i < b.n; i++) sweepFromLeaf(b, i, 0u);
        collapse(b);
    }
};
static void emulInstanceTables(EmulScene* s) {           // instanceTables of ohb_api.cu
    uint32_t ni = uint32_t(s->instances.size());
    s->instXform.assign(size_t(ni) * 3 + 3, f4{}); s->instNormalMat.assign(size_t(ni) * 3 + 3, f4{}); s->instInv.assign(size_t(ni) * 3 + 3, f4{});
    for (uint32_t i = 0; i < ni; i++) {
        const float* a = s->instances[i].xform;
        memcpy(&s->instXform[size_t(i) * 3], a, 48);
        float co[3][3];
        co[0][0] = a[5] * a[10] - a[6] * a[9];  co[0][1] = a[6] * a[8] - a[4] * a[10]; co[0][2] = a[4] * a[9] - a[5] * a[8];
        co[1][0] = a[2] * a[9] - a[1] * a[10];  co[1][1] = a[0] * a[10] - a[2] * a[8]; co[1][2] = a[1] * a[8] - a[0] * a[9];
        co[2][0] = a[1] * a[6] - a[2] * a[5];   co[2][1] = a[2] * a[4] - a[0] * a[6];  co[2][2] = a[0] * a[5] - a[1] * a[4];
        float det = a[0] * co[0][0] + a[1] * co[0][1] + a[2] * co[0][2], id = 1.0f / det;
        float nm[12], iv[12];
        for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) nm[r * 4 + k] = co[r][k] * id; nm[r * 4 + 3] = 0; }
        for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) iv[r * 4 + k] = nm[k * 4 + r]; iv[r * 4 + 3] = -(iv[r * 4] * a[3] + iv[r * 4 + 1] * a[7] + iv[r * 4 + 2] * a[11]); }
        memcpy(&s->instNormalMat[size_t(i) * 3], nm, 48); memcpy(&s->instInv[size_t(i) * 3], iv, 48);
    }
}
static BuildArrays emulMainArrays(EmulScene* s, uint32_t n) {
    BuildArrays b = s->mainB->arrays(n);
    b.positions = s->positions.data(); b.posStride = s->stride; b.indices = s->indices.data(); b.triInst = s->triInst.data(); b.instXform = s->instXform.data();
    b.activeTris = s->active.data(); b.wnodes = s->wnodes.data(); b.tris = s->tris.data();
    return b;
}
// ohb_build_accel, both modes
static void emulBuildAccel(EmulScene* s) {
    if (!s->mainB) s->mainB = new HostBuilder();
    if (!s->tlasB) s->tlasB = new HostBuilder();
    uint32_t ni = uint32_t(s->instances.size());
    s->triInst.assign(s->ntris, 0xFFFFFFFFu); s->active.clear(); s->ranges.clear();
    emulInstanceTables(s);
    for (uint32_t i = 0; i < ni; i++) {
        const ohb_instance& in = s->instances[i];
        if ((in.mask & 0xFFu) == 0u) continue;
        uint32_t first = uint32_t(s->active.size());
        for (uint32_t t = in.first_tri; t < in.first_tri + in.tri_count && t < s->ntris; t++) { if (s->triInst[t] == 0xFFFFFFFFu) s->active.push_back(t); s->triInst[t] = i; }
        if (s->active.size() > first) s->ranges.push_back({i, first, uint32_t(s->active.size()) - first});
    }
    uint32_t n = uint32_t(s->active.size()); s->numActive = n; s->numNodes = 0; s->levels = 0; s->sah = 0;
    if (n == 0) return;
    s->wnodes.assign(size_t(n) * OHB_WNODE_VECS, u4{0, 0, 0, 0}); s->tris.assign(size_t(n) * 3, f4{});
    if (!s->accelMode) {
        BuildArrays b = emulMainArrays(s, n);
        s->mainB->build(b);
        s->numNodes = s->mainB->wideCounters[0]; s->levels = s->mainB->wideCounters[3] + 1;
        float ra = n >= 2 ? boxArea(xyz(s->mainB->nodeLo[0]), xyz(s->mainB->nodeHi[0])) : 0.0f;
        s->sah = ra > 0 ? (s->mainB->sah[0] + s->mainB->sah[1]) / ra : float(n);
        return;
    }
    s->blasLo.assign(ni + 1, f4{}); s->blasHi.assign(ni + 1, f4{}); s->blasInfo.assign(ni + 1, u4{0, 0, 0, 0}); s->instOfPrim.clear();
    uint32_t blasLevels = 0;
    for (const EmulScene::Range& r : s->ranges) {
        BuildArrays b = emulMainArrays(s, r.count);
        b.objectSpace = 1u; b.activeTris = s->active.data() + r.first; b.wnodes = s->wnodes.data() + size_t(r.first) * OHB_WNODE_VECS; b.tris = s->tris.data() + size_t(r.first) * 3u;
        s->mainB->build(b);
        storeBlasRootBox(b, s->blasLo.data(), s->blasHi.data(), r.inst, &blasLevels);
        s->blasInfo[r.inst] = u4{r.first, r.first, r.count, 0u}; s->instOfPrim.push_back(r.inst);
    }
    uint32_t np = uint32_t(s->ranges.size());
    s->tlasNodes.assign(size_t(np + 1) * OHB_WNODE_VECS, u4{0, 0, 0, 0}); s->tlasLeaves.assign(size_t(np + 1) * 3, f4{});
    BuildArrays tb = s->tlasB->arrays(np); tb.instXform = s->instXform.data(); tb.wnodes = s->tlasNodes.data(); tb.tris = s->tlasLeaves.data();
    s->tlasB->tlas(tb, s->blasLo.data(), s->blasHi.data(), s->instOfPrim.data(), false);
    s->numNodes = s->tlasB->wideCounters[0]; s->levels = s->tlasB->wideCounters[3] + 1 + blasLevels + 1;
}
// ohb_update_instances
static void emulUpdateInstances(EmulScene* s, const ohb_instance* inst, uint32_t n) {
    for (uint32_t i = 0; i < n && i < s->instances.size(); i++) memcpy(s->instances[i].xform, inst[i].xform, 48);
    emulInstanceTables(s);
    if (!s->numActive) return;
    if (s->accelMode) {
        uint32_t np = uint32_t(s->ranges.size());
        BuildArrays tb = s->tlasB->view(np); tb.instXform = s->instXform.data(); tb.wnodes = s->tlasNodes.data(); tb.tris = s->tlasLeaves.data();
        s->tlasB->tlas(tb, s->blasLo.data(), s->blasHi.data(), s->instOfPrim.data(), true);
        s->numNodes = s->tlasB->wideCounters[0];
    } else {
        BuildArrays b = s->mainB->view(s->numActive);
        b.positions = s->positions.data(); b.posStride = s->stride; b.indices = s->indices.data(); b.triInst = s->triInst.data(); b.instXform = s->instXform.data();
        b.activeTris = s->active.data(); b.wnodes = s->wnodes.data(); b.tris = s->tris.data();
        s->mainB->refit(b);
        s->numNodes = s->mainB->wideCounters[0];
    }
}

extern "C" {

struct emul_scene_desc {
    const void* positions; uint64_t stride_bytes; uint32_t nverts;
    const uint32_t* indices; uint32_t ntris;
    const float* normals; const float* uvs; const uint32_t* mat_ids;
    const ohb_instance* instances; uint32_t ninstances;
    const float* mat_colors; uint32_t nmaterials;
    const uint8_t* textures; uint32_t tex_w, tex_h, tex_layers;
    const void* light_ssbo; uint64_t light_bytes;
    const float* env; uint32_t env_w, env_h;
    const float* marg; const float* cond; float env_integral;   // CDFs supplied by the caller (device kernel not emulated)
};

void* emul_scene_create(const emul_scene_desc* d) {
    EmulScene* s = new EmulScene();
    s->stride = d->stride_bytes; s->nverts = d->nverts; s->ntris = d->ntris;
    s->positions.assign((const uint8_t*)d->positions, (const uint8_t*)d->positions + size_t(d->nverts) * d->stride_bytes);
    s->indices.assign(d->indices, d->indices + size_t(d->ntris) * 3);
    s->normals.resize(d->nverts); memcpy(s->normals.data(), d->normals, size_t(d->nverts) * 16);
    s->uvs.resize(d->nverts); memcpy(s->uvs.data(), d->uvs, size_t(d->nverts) * 8);
    s->matIds.assign(d->mat_ids, d->mat_ids + d->ntris);
    s->matColors.resize(size_t(d->nmaterials) * 3); memcpy(s->matColors.data(), d->mat_colors, size_t(d->nmaterials) * 48);
    if (d->textures && d->tex_layers) { s->texW = d->tex_w; s->texH = d->tex_h; s->texLayers = d->tex_layers; s->tex.assign(d->textures, d->textures + size_t(d->tex_w) * d->tex_h * 4 * d->tex_layers); }
    if (d->light_ssbo && d->light_bytes >= 16) {
        const uint8_t* lb = (const uint8_t*)d->light_ssbo;
        memcpy(&s->lightCount, lb, 4); memcpy(&s->envMapTexIdx, lb + 4, 4); memcpy(&s->envIntensity, lb + 8, 4);
        s->lights.resize(s->lightCount + 1); memcpy(s->lights.data(), lb + 16, size_t(s->lightCount) * 80);
    }
    if (d->env && d->env_w) {
        s->envW = d->env_w; s->envH = d->env_h; s->env.resize(size_t(d->env_w) * d->env_h); memcpy(s->env.data(), d->env, s->env.size() * 16);
        s->marg.assign(d->marg, d->marg + d->env_h); s->cond.assign(d->cond, d->cond + size_t(d->env_w) * d->env_h); s->envIntegral = d->env_integral;
        if (d->env_w % 32u == 0u && d->env_h % 32u == 0u && !getenv("OHB_ENV_BLOCKED_OFF")) {      // k_env_tops
            const uint32_t wb = d->env_w >> 5, hb = d->env_h >> 5;
            s->condTop.resize(size_t(d->env_h) * wb); s->margTop.resize(hb);
            for (uint32_t y = 0; y < d->env_h; y++) for (uint32_t b = 0; b < wb; b++) s->condTop[size_t(y) * wb + b] = s->cond[size_t(y) * d->env_w + (b << 5) + 31u];
            for (uint32_t b = 0; b < hb; b++) s->margTop[b] = s->marg[(b << 5) + 31u];
        }
    } else { s->marg.assign(1, 1.0f); s->cond.assign(1, 1.0f); }
    s->instances.assign(d->instances, d->instances + d->ninstances);
    emulBuildAccel(s);
    return s;
}
void emul_scene_destroy(void* h) { EmulScene* s = (EmulScene*)h; delete s->mainB; delete s->tlasB; delete s; }
void emul_scene_set_accel_mode(void* h, int mode) { EmulScene* s = (EmulScene*)h; if ((mode != 0) != (s->accelMode != 0)) { s->accelMode = mode ? 1 : 0; emulBuildAccel(s); } }
void emul_scene_update_instances(void* h, const ohb_instance* inst, uint32_t n) { emulUpdateInstances((EmulScene*)h, inst, n); }
void emul_accel_stats(void* h, uint32_t* numNodes, float* sah) { EmulScene* s = (EmulScene*)h; *numNodes = s->numNodes; *sah = s->sah; }
uint32_t emul_accel_levels(void* h) { return ((EmulScene*)h)->levels; }
// ---- SVGF denoiser (ohb_svgf.h) driven like launchSvgf (ohb_kernels.cu) ---------------------------------------
struct emul_svgf_args {
    int32_t width, height, reset, nthreads;
    float sigma_l, sigma_normal, sigma_depth, _pad;
    uint32_t* beauty; const uint32_t* motion; const float* depth; const f4* normal;
    const h4* prev_color; const h4* prev_moments; const h4* prev_geom;
    h4* cur_color; h4* cur_moments; h4* cur_geom;
};
int emul_svgf_dispatch(emul_svgf_args* a) {
    const int W = a->width, H = a->height; const size_t n = size_t(W) * H;
    std::vector<h4> colorA(n), colorB(n); std::vector<uint16_t> varA(n), varB(n);
    SvgfTemporalArgs t{};
    t.beauty = a->beauty; t.motion = a->motion; t.depth = a->depth; t.normal = a->normal;
    t.prevColor = a->prev_color; t.prevMoments = a->prev_moments; t.prevGeom = a->prev_geom;
    t.outColor = colorA.data(); t.outMoments = a->cur_moments; t.outVariance = varA.data(); t.outGeom = a->cur_geom;
    t.W = W; t.H = H; t.reset = a->reset;
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) svgfTemporalPixel(t, x, y);
    h4* A = colorA.data(); h4* B = colorB.data(); h4* HC = a->cur_color;
    h4* inC[OHB_SVGF_ITERATIONS] = {A, HC, B, A, B}; h4* outC[OHB_SVGF_ITERATIONS] = {HC, B, A, B, A};
    uint16_t* var[2] = {varA.data(), varB.data()};
    for (int it = 0; it < OHB_SVGF_ITERATIONS; it++) {
        SvgfAtrousArgs k{};
        k.inColor = inC[it]; k.outColor16 = outC[it]; k.normal = a->normal; k.depth = a->depth; k.inVar = var[it & 1]; k.outVar = var[1 - (it & 1)]; k.outLDR = a->beauty;
        k.W = W; k.H = H; k.stepSize = 1 << it; k.isFinal = it == OHB_SVGF_ITERATIONS - 1; k.sigmaL = a->sigma_l; k.sigmaNormal = a->sigma_normal; k.sigmaDepth = a->sigma_depth;
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) svgfAtrousPixel(k, x, y);
    }
    return 0;
}
// guide AOVs as k_rt_pixel writes them: proj * view of this frame, row 2 of view
void emul_svgf_guides(const float* surf, uint32_t W, uint32_t H, const float* view, const float* proj, const float* prevViewProj, uint32_t frameIdx, uint32_t* motion, float* depth) {
    float currVP[16], row2[4] = {view[2], view[6], view[10], view[14]};
    for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) { float acc = 0.0f; for (int k = 0; k < 4; k++) acc += proj[k * 4 + row] * view[col * 4 + k]; currVP[col * 4 + row] = acc; }
    for (size_t pi = 0; pi < size_t(W) * H; pi++) {
        bool hit = surf[pi * 4 + 3] > 0.0f;
        svgfGuides(currVP, prevViewProj, row2, W, H, frameIdx, hit, hit ? mk3(surf[pi * 4], surf[pi * 4 + 1], surf[pi * 4 + 2]) : mk3(0.0f), motion[pi], depth[pi]);
    }
}
uint16_t emul_f2h(float f) { return f2h(f); }
float emul_h2f(uint16_t h) { return h2f(h); }
void emul_trav_stats(unsigned long long* nodes, unsigned long long* tris, int reset) { *nodes = g_statNodes; *tris = g_statTris; if (reset) { g_statNodes = 0; g_statTris = 0; } }
void emul_set_warp_noise(unsigned on) { g_emulWarpNoise = on; g_emulWarpState = 12345u; }

void emul_hybrid_shadow(void* h, uint32_t W, uint32_t H, const f4* gPos, const f2* gNrm, const ohb_hybrid_shadow_params* p, uint8_t* mask) {     // k_hybrid_shadow
    SceneDev sc = ((EmulScene*)h)->dev();
    HybridShadowParams pc{}; pc.lightDir = mk3(p->light_dir[0], p->light_dir[1], p->light_dir[2]); pc.lightRadius = p->light_radius;
    pc.lightPos = mk3(p->light_pos[0], p->light_pos[1], p->light_pos[2]); pc.lightRange = p->light_range; pc.W = W; pc.H = H; pc.lightType = p->light_type; pc.sampleCount = p->sample_count;
    for (uint32_t y = 0; y < H; y++) for (uint32_t x = 0; x < W; x++) mask[size_t(y) * W + x] = hybridShadowPixel(sc, pc, gPos, gNrm, x, y);
}
void emul_hybrid_gi(void* h, uint32_t W, uint32_t H, const f4* gPos, const f2* gNrm, const f4* gAlbedo, const f4* hist, const f4* instMat, const ohb_hybrid_gi_params* p, h4* out) {   // k_hybrid_gi
    SceneDev sc = ((EmulScene*)h)->dev();
    HybridGiParams pc{}; pc.lightPos = mk3(p->light_pos[0], p->light_pos[1], p->light_pos[2]); pc.lightIntensity = p->light_intensity; pc.W = W; pc.H = H; pc.sampleCount = p->sample_count; pc.frameIndex = p->frame_index;
    for (uint32_t y = 0; y < H; y++) for (uint32_t x = 0; x < W; x++) out[size_t(y) * W + x] = hybridGiPixel(sc, pc, gPos, gNrm, gAlbedo, hist, instMat, x, y);
}
void emul_nrd_pack_batch(const float* in6, const float* nr4, uint32_t n, f4* packedRad, f4* packedNormal, float* unpackedRgb) {     // k_nrd_pack
    for (uint32_t i = 0; i < n; i++) {
        const float* a = in6 + size_t(i) * 6u;
        f4 pr = nrdPackRadianceHitDist(mk3(a[0], a[1], a[2]), a[3], a[4], a[5]);
        packedRad[i] = pr; packedNormal[i] = nrdPackNormalRoughness(mk3(nr4[4 * i], nr4[4 * i + 1], nr4[4 * i + 2]), nr4[4 * i + 3]);
        f3 back = nrdYCoCgToLinear(xyz(pr));
        unpackedRgb[3 * i] = back.x; unpackedRgb[3 * i + 1] = back.y; unpackedRgb[3 * i + 2] = back.z;
    }
}
void emul_trace_batch(void* h, const ohb_ray* rays, uint32_t n, ohb_hit* hits) {
    SceneDev sc = ((EmulScene*)h)->dev();
    for (uint32_t i = 0; i < n; i++)
        hits[i] = traceClosest(sc, mk3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), mk3(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]), rays[i].tmin, rays[i].tmax);
}
void emul_occluded_batch(void* h, const ohb_ray* rays, uint32_t n, uint8_t* occ) {
    SceneDev sc = ((EmulScene*)h)->dev();
    for (uint32_t i = 0; i < n; i++)
        occ[i] = traceAny(sc, mk3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), mk3(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]), rays[i].tmin, rays[i].tmax);
}
void emul_env_sample_batch(void* h, const float* u12, uint32_t n, float* dirPdf, float* pdfOfDir) {
    SceneDev sc = ((EmulScene*)h)->dev();
    for (uint32_t i = 0; i < n; i++) { f3 d; float p; sampleEnvMap(sc, u12[2 * i], u12[2 * i + 1], d, p); dirPdf[4 * i] = d.x; dirPdf[4 * i + 1] = d.y; dirPdf[4 * i + 2] = d.z; dirPdf[4 * i + 3] = p; pdfOfDir[i] = pdfEnvMap(sc, d); }
}

struct emul_render_args {
    float view[16], proj[16]; uint32_t width, height, first_sample_index, history_count, nsamples;
    uint32_t tile_x, tile_y, tile_w, tile_h; ohb_settings settings;
    float* accum; uint8_t* ldr; float* albedo; float* normal; float* sample_dump; uint64_t counters[4];
};
static void inv4(const float* m, float* out) {
    double a[4][8];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { a[r][c] = m[c * 4 + r]; a[r][c + 4] = r == c; }
    for (int i = 0; i < 4; i++) {
        int p = i; for (int r = i + 1; r < 4; r++) if (fabs(a[r][i]) > fabs(a[p][i])) p = r;
        for (int c = 0; c < 8; c++) std::swap(a[i][c], a[p][c]);
        double d = a[i][i]; for (int c = 0; c < 8; c++) a[i][c] /= d;
        for (int r = 0; r < 4; r++) if (r != i) { double f = a[r][i]; for (int c = 0; c < 8; c++) a[r][c] -= f * a[i][c]; }
    }
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out[c * 4 + r] = float(a[r][c + 4]);
}
// Mimics ohb_render + launchOfflineBatch with one batch per call.
int emul_render_offline(void* h, emul_render_args* a) {
    EmulScene* es = (EmulScene*)h; SceneDev sc = es->dev();
    float iv[16], ip[16]; inv4(a->view, iv); inv4(a->proj, ip);
    FrameParams fr{};
    fr.camPos = mk3(iv[12], iv[13], iv[14]); fr.fwd = mk3(-iv[8], -iv[9], -iv[10]); fr.right = mk3(iv[0], iv[1], iv[2]); fr.up = mk3(iv[4], iv[5], iv[6]);
    fr.tanY = fabsf(ip[5]); fr.tanX = fr.tanY * (float(a->width) / float(a->height));
    fr.W = a->width; fr.H = a->height; fr.maxBounces = a->settings.max_bounces & 0xFFFFu; fr.flags = a->settings.flags;
    bool envOn = es->envW && es->envMapTexIdx != 0xFFFFFFFFu;
    fr.envW = envOn ? es->envW : 0u; fr.envH = envOn ? float(es->envH) : 0.0f;
    fr.fireflyClamp = a->settings.firefly_clamp_lum; fr.sss = a->settings.subsurface_strength; fr.aniso = a->settings.anisotropy_strength; fr.anisoRot = a->settings.anisotropy_rotation;
    fr.samplerType = a->settings.sampler_type;
    fr.tileX = a->tile_x; fr.tileY = a->tile_y; fr.tileW = a->tile_w ? a->tile_w : a->width; fr.tileH = a->tile_h ? a->tile_h : a->height;
    uint32_t tilesX = (fr.tileW + 7u) / 8u, tilesY = (fr.tileH + 3u) / 4u, numPixels = tilesX * tilesY * 32u;
    uint32_t total = numPixels * a->nsamples;
    std::vector<f4> rayO(total), rayD(total), thr(total), rad(total), pendA(total), pendB(total), fh0(total), fh1(total), fh2(total), pay0(total), pay1(total), pay2(total), pay3(total), shO(size_t(total) * 2), shD(size_t(total) * 2);
    std::vector<ohb_hit> hit(total); std::vector<u4> meta(total); std::vector<uint32_t> qa(total), qb(total);
    uint32_t small[4] = {0, 0, 0, 0}; unsigned long long counters[8] = {0};
    size_t npx = size_t(a->width) * a->height;
    std::vector<f4> albedo(npx), normal(npx), accum(npx); std::vector<uint32_t> ldr(npx);
    memcpy(accum.data(), a->accum, npx * 16);
    std::vector<u4> sobolTab(a->nsamples);
    for (uint32_t i = 0; i < a->nsamples; i++) sobolTab[i] = sobolQuad(a->first_sample_index + i);
    std::vector<uint32_t> qs(total); uint32_t sortCount[2] = {0, 0};
    PathArrays P{};
    P.sobolTab = sobolTab.data(); P.queueSorted = qs.data(); P.sortCount = sortCount;
    P.rayO = rayO.data(); P.rayD = rayD.data(); P.hit = hit.data(); P.thr = thr.data(); P.rad = rad.data(); P.pendA = pendA.data(); P.pendB = pendB.data(); P.meta = meta.data();
    P.fh0 = fh0.data(); P.fh1 = fh1.data(); P.fh2 = fh2.data(); P.pay0 = pay0.data(); P.pay1 = pay1.data(); P.pay2 = pay2.data(); P.pay3 = pay3.data(); P.shO = shO.data(); P.shD = shD.data();
    P.queueIn = qa.data(); P.queueOut = qb.data(); P.countIn = &small[0]; P.countOut = &small[1]; P.shCount = &small[2]; P.counters = counters;
    P.albedoAOV = albedo.data(); P.normalAOV = normal.data(); P.numPixels = numPixels; P.samplesInBatch = a->nsamples; P.firstSampleIndex = a->first_sample_index;
    for (uint32_t p = 0; p < total; p++) { raygenPath(fr, P, p); if (OHB_ST_STAGE(P.meta[p].w) != ST_DONE) P.queueIn[(*P.countIn)++] = p | OHB_Q_PRIMARY; }
    counters[0] += *P.countIn;
    uint32_t iters = 1u + 2u * fr.maxBounces;
    for (uint32_t it = 0; it < iters; it++) {
        uint32_t n = *P.countIn;
        for (uint32_t i = 0; i < n; i++) { uint32_t p = OHB_Q_PATH(P.queueIn[i]); P.hit[p] = traceClosest(sc, xyz(P.rayO[p]), xyz(P.rayD[p]), 0.001f, 10000.0f); counters[3] += P.hit[p].prim != OHB_MISS; }
        sortCount[0] = sortCount[1] = 0;                                                                                                 // k_sort_hits
        for (uint32_t i = 0; i < n; i++) { uint32_t e = P.queueIn[i]; if (P.hit[OHB_Q_PATH(e)].prim != OHB_MISS) P.queueSorted[sortCount[0]++] = e; else P.queueSorted[n - 1u - sortCount[1]++] = e | OHB_Q_MISS; }
        for (uint32_t i = 0; i < n; i++) { uint32_t e = shadePath(sc, fr, P, P.queueSorted[i]); if (e != OHB_Q_NONE) P.queueOut[(*P.countOut)++] = e; }   // k_shade
        uint32_t ns = *P.shCount;
        for (uint32_t i = 0; i < ns; i++) {
            f4 o = P.shO[i], d = P.shD[i];
            if (traceAny(sc, xyz(o), xyz(d), 0.001f, o.w)) { uint32_t tag = f2u(d.w); ((tag & 1u) ? P.pendB : P.pendA)[tag >> 1] = mk4(0, 0, 0, 0); }
        }
        counters[1] += n; counters[2] += ns; *P.countIn = 0; *P.shCount = 0;
        std::swap(P.queueIn, P.queueOut); std::swap(P.countIn, P.countOut);
    }
    FilmArrays F{}; F.accum = accum.data(); F.ldr = ldr.data(); F.sampleDump = a->sample_dump; F.historyCount = a->history_count; F.sumMode = 0;
    for (uint32_t pix = 0; pix < numPixels; pix++) filmPixel(fr, P, F, pix);
    memcpy(a->accum, accum.data(), npx * 16);
    if (a->ldr) memcpy(a->ldr, ldr.data(), npx * 4);
    if (a->albedo) memcpy(a->albedo, albedo.data(), npx * 16);
    if (a->normal) memcpy(a->normal, normal.data(), npx * 16);
    for (int i = 0; i < 4; i++) a->counters[i] = counters[i];
    return 0;
}

// Mimics renderRealtime + launchRealtimeFrame for ONE frame (the caller owns the ping-ponged history images).
struct emul_rt_args {
    float view[16], proj[16], prev_view_proj[16];
    uint32_t width, height, frame_index, history_count, view_changed;
    ohb_settings settings;
    const float* accum_prev; float* accum_curr; const float* surf_prev; float* surf_curr; const float* shad_prev; float* shad_curr;
    const float* res_prev[3]; float* res_curr[3];
    float* albedo; float* normal; float* radiance_dump; float* gi_dump; float* denoised; uint8_t* ldr;
    uint64_t counters[4];
};
int emul_render_realtime(void* h, emul_rt_args* a) {
    EmulScene* es = (EmulScene*)h; SceneDev sc = es->dev();
    float iv[16], ip[16]; inv4(a->view, iv); inv4(a->proj, ip);
    FrameParams fr{};
    fr.camPos = mk3(iv[12], iv[13], iv[14]); fr.fwd = mk3(-iv[8], -iv[9], -iv[10]); fr.right = mk3(iv[0], iv[1], iv[2]); fr.up = mk3(iv[4], iv[5], iv[6]);
    fr.tanY = fabsf(ip[5]); fr.tanX = fr.tanY * (float(a->width) / float(a->height));
    fr.W = a->width; fr.H = a->height; fr.maxBounces = a->settings.max_bounces & 0xFFFFu; fr.flags = a->settings.flags;
    bool envOn = es->envW && es->envMapTexIdx != 0xFFFFFFFFu;
    fr.envW = envOn ? es->envW : 0u; fr.envH = envOn ? float(es->envH) : 0.0f;
    fr.fireflyClamp = a->settings.firefly_clamp_lum; fr.sss = a->settings.subsurface_strength; fr.aniso = a->settings.anisotropy_strength; fr.anisoRot = a->settings.anisotropy_rotation;
    fr.samplerType = a->settings.sampler_type;
    fr.tileX = 0; fr.tileY = 0; fr.tileW = a->width; fr.tileH = a->height;
    memcpy(fr.prevViewProj, a->prev_view_proj, 64);
    uint32_t spf = a->settings.samples_per_frame; spf = spf < 1u ? 1u : (spf > 64u ? 64u : spf);
    fr.frameIdx = a->frame_index; fr.historyCount = a->history_count; fr.viewChanged = a->view_changed; fr.spf = spf; fr.jitterSobol = sobolQuad(fr.frameIdx);
    uint32_t tilesX = (fr.tileW + 7u) / 8u, tilesY = (fr.tileH + 3u) / 4u, numPixels = tilesX * tilesY * 32u, total = numPixels * spf;
    std::vector<f4> rayO(total), rayD(total), thr(total), rad(total), pendA(total), pendB(total), fh0(total), fh1(total), fh2(total),
        pay0(total), pay1(total), pay2(total), pay3(total), shO(size_t(total) * 2), shD(size_t(total) * 2);
    std::vector<ohb_hit> hit(total); std::vector<u4> meta(total); std::vector<uint32_t> qa(total), qb(total), qs(total);
    uint32_t small[4] = {0, 0, 0, 0}, sortCount[2] = {0, 0}; unsigned long long counters[8] = {0};
    std::vector<u4> sobolTab(spf); for (uint32_t i = 0; i < spf; i++) sobolTab[i] = sobolQuad(fr.frameIdx * spf + i);
    size_t npx = size_t(a->width) * a->height;
    std::vector<uint32_t> ldr(npx);
    PathArrays P{};
    P.rayO = rayO.data(); P.rayD = rayD.data(); P.hit = hit.data(); P.thr = thr.data(); P.rad = rad.data(); P.pendA = pendA.data(); P.pendB = pendB.data(); P.meta = meta.data();
    P.fh0 = fh0.data(); P.fh1 = fh1.data(); P.fh2 = fh2.data(); P.pay0 = pay0.data(); P.pay1 = pay1.data(); P.pay2 = pay2.data(); P.pay3 = pay3.data();
    P.shO = shO.data(); P.shD = shD.data(); P.queueIn = qa.data(); P.queueOut = qb.data(); P.queueSorted = qs.data(); P.sortCount = sortCount;
    P.countIn = &small[0]; P.countOut = &small[1]; P.shCount = &small[2]; P.counters = counters;
    P.albedoAOV = reinterpret_cast<f4*>(a->albedo); P.normalAOV = reinterpret_cast<f4*>(a->normal);
    P.numPixels = numPixels; P.samplesInBatch = spf; P.firstSampleIndex = fr.frameIdx * spf; P.sobolTab = sobolTab.data();
    for (uint32_t p = 0; p < total; p++) { raygenPathRT(fr, P, p); if (OHB_ST_STAGE(P.meta[p].w) != ST_DONE) P.queueIn[(*P.countIn)++] = p; }
    counters[0] += *P.countIn;
    for (uint32_t it = 0; it < 2u + fr.maxBounces; it++) {
        uint32_t n = *P.countIn;
        for (uint32_t i = 0; i < n; i++) { uint32_t p = P.queueIn[i]; P.hit[p] = traceClosest(sc, xyz(P.rayO[p]), xyz(P.rayD[p]), 0.001f, 10000.0f); counters[3] += P.hit[p].prim != OHB_MISS; }
        sortCount[0] = sortCount[1] = 0;
        for (uint32_t i = 0; i < n; i++) { uint32_t p = P.queueIn[i]; if (surfacePath(sc, fr, P, p)) P.queueSorted[sortCount[0]++] = p; else P.queueSorted[n - 1u - sortCount[1]++] = p; }
        for (uint32_t i = 0; i < n; i++) { uint32_t p = P.queueSorted[i]; if (bouncePathRT(sc, fr, P, p)) P.queueOut[(*P.countOut)++] = p; }
        uint32_t ns = *P.shCount;
        for (uint32_t i = 0; i < ns; i++) {
            f4 o = P.shO[i], d = P.shD[i];
            if (traceAny(sc, xyz(o), xyz(d), 0.001f, o.w)) { uint32_t tag = f2u(d.w); ((tag & 1u) ? P.pendB : P.pendA)[tag >> 1] = mk4(0, 0, 0, 0); }
        }
        counters[1] += n; counters[2] += ns; *P.countIn = 0; *P.shCount = 0;
        std::swap(P.queueIn, P.queueOut); std::swap(P.countIn, P.countOut);
    }
    RTImagesDev im{};
    im.accumPrev = reinterpret_cast<const f4*>(a->accum_prev); im.accumCurr = reinterpret_cast<f4*>(a->accum_curr);
    im.surfPrev = reinterpret_cast<const f4*>(a->surf_prev); im.surfCurr = reinterpret_cast<f4*>(a->surf_curr);
    im.shadPrev = reinterpret_cast<const f4*>(a->shad_prev); im.shadCurr = reinterpret_cast<f4*>(a->shad_curr);
    im.res0Prev = reinterpret_cast<const f4*>(a->res_prev[0]); im.res1Prev = reinterpret_cast<const f4*>(a->res_prev[1]); im.res2Prev = reinterpret_cast<const f4*>(a->res_prev[2]);
    im.res0Curr = reinterpret_cast<f4*>(a->res_curr[0]); im.res1Curr = reinterpret_cast<f4*>(a->res_curr[1]); im.res2Curr = reinterpret_cast<f4*>(a->res_curr[2]);
    im.radianceDump = a->radiance_dump; im.giDump = a->gi_dump; im.counters = counters;
    for (uint32_t pix = 0; pix < numPixels; pix++) pixelRT(sc, fr, P, im, pix);
    for (uint32_t i = 0; i < uint32_t(npx); i++) denoiseRT(fr, im.accumCurr, P.normalAOV, ldr.data(), a->denoised, i);
    if (a->ldr) memcpy(a->ldr, ldr.data(), npx * 4);
    for (int i = 0; i < 4; i++) a->counters[i] = counters[i];
    return 0;
}

}  // extern "C"

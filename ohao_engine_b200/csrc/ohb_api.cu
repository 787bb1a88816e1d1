// ohb_api.cu — the C ABI of include/ohao_b200.h: context, device memory, scene upload,
// acceleration-structure build, the render loop and the readbacks.
//
// Host-side mirror of PathTracer (ohao/render/rt/path_tracer.{hpp,cpp}, path_tracer_render.cpp):
// the context owns its output/accumulation/AOV images and *copies* the scene arrays it is handed
// (the Vulkan profile borrowed VkBuffers; a C ABI cannot borrow host memory safely), keeps
// m_sampleIndex / m_historyFrameCount / m_renderSeed with the reference's reset rules
// (path_tracer.cpp:189-211) and turns one ohb_render() into `nsamples` reference frames.
// There is no CPU path: every entry point needs a live CUDA context.
#include "ohb_device.h"
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

using namespace ohb;

namespace {

struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    cudaError_t reserve(size_t n) {
        if (n <= bytes && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// scratch of one builder instance (launchBuild): the two-level build runs several BLAS builds concurrently, one set per stream
struct BuilderSet {
    DevBuf wtri, primLo, primHi, boundsBits, keys, vals, keysTmp, valsTmp, sortTemp, left, right, parentInner, parentLeaf, nodeLo, nodeHi, visit, wideCounters, itemsA, itemsB, sah;
    cudaStream_t stream = nullptr; cudaEvent_t done = nullptr;
    void release() {
        DevBuf* all[] = {&wtri, &primLo, &primHi, &boundsBits, &keys, &vals, &keysTmp, &valsTmp, &sortTemp, &left, &right, &parentInner, &parentLeaf, &nodeLo, &nodeHi, &visit, &wideCounters, &itemsA, &itemsB, &sah};
        for (DevBuf* b : all) b->release();
        if (stream) cudaStreamDestroy(stream);
        if (done) cudaEventDestroy(done);
        stream = nullptr; done = nullptr;
    }
};
struct M4h { float m[16]; };
static M4h inverse4(const float* m) {   // cofactor inverse, fp32 (what glm::inverse computes)
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    float id = 1.0f / det;
    M4h o; for (int i = 0; i < 16; i++) o.m[i] = inv[i] * id;
    return o;
}

static std::string g_createError;

}  // namespace

struct ohb_ctx {
    int device = 0; cudaStream_t stream = nullptr; int numSMs = 148;
    uint32_t W = 0, H = 0; int profile = 0;
    ohb_settings settings{};
    uint32_t seed = 0, sampleIndex = 0, historyCount = 0; bool viewChanged = false;
    uint32_t tileX = 0, tileY = 0, tileW = 0, tileH = 0;
    int sumMode = 0;
    std::string err;
    uint64_t launches = 0;
    // scene (host metadata)
    uint32_t nverts = 0, ntris = 0, nmat = 0, maxMatId = 0; uint64_t posStride = 0;
    std::vector<ohb_instance> instances;
    uint32_t lightCount = 0, envMapTexIdx = 0xFFFFFFFFu; float envIntensity = 1.0f;
    uint32_t texW = 0, texH = 0, texLayers = 0, envW = 0, envH = 0; float envIntegral = 0.0f;
    bool accelValid = false; uint32_t numActive = 0;
    int accelMode = OHB_ACCEL_FLATTEN; uint32_t numTlasPrims = 0;
    std::vector<BuilderSet> blasSets;        // extra builder scratch + streams for concurrent BLAS builds (set 0 = the context's own buffers and stream)
    ohb_accel_stats stats{};
    // scene (device)
    DevBuf positions, indices, normals, uvs, matIds, triInst, instXform, instNormalMat, instInv, matColors, tex, lights, env, marg, cond, rowTotal, integral, margTop, condTop;
    bool envTops = false;
    // accel (device)
    DevBuf activeTris, wtri, primLo, primHi, boundsBits, keys, vals, keysTmp, valsTmp, sortTemp, left, right, parentInner, parentLeaf,
           nodeLo, nodeHi, visit, wideCounters, wideItemsA, wideItemsB, sah, wnodes, tris;
    // two-level structure: per-instance BLAS tables + the TLAS with its own (persistent) builder state for MODE_UPDATE refits
    DevBuf blasLo, blasHi, blasInfo, instOfPrim, maxLevels, tlasNodes, tlasLeaves;
    DevBuf tWtri, tPrimLo, tPrimHi, tBounds, tKeys, tVals, tKeysTmp, tValsTmp, tSortTemp, tLeft, tRight, tParentInner, tParentLeaf, tNodeLo, tNodeHi, tVisit, tWideCounters, tItemsA, tItemsB, tSah;
    // paths (device)
    DevBuf rayO, rayD, hit, thr, rad, pendA, pendB, meta, fh0, fh1, fh2, pay0, pay1, pay2, pay3, shO, shD, queueA, queueB, queueS, hitFlag, sobolTab, smallCounters, devCounters, octPerm, octHist, octScanTmp;
    uint32_t pathCapacity = 0;
    // film
    DevBuf accum, ldr, albedoAOV, normalAOV, sampleDump;
    // realtime profile: ping-ponged history (index rtCur = the images written by the last frame)
    DevBuf accumPrev, surf[2], shad[2], res[2][3], rtDump[3];
    // SVGF denoiser (DenoiseMode::Atrous): guide AOVs, ping-ponged history, scratch; svgfCur = history written by the last frame
    DevBuf motionAOV, depthAOV, svgfHistColor[2], svgfHistMoments[2], svgfHistGeom[2], svgfColor[2], svgfVar[2];
    int svgfCur = 0;
    int rtCur = 0; float prevViewProj[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float* rtDumpHost[3] = {nullptr, nullptr, nullptr};
    float* dumpHost = nullptr; size_t dumpCapacity = 0;
    // hooks
    DevBuf hookRays, hookOut, hookAux;
    TimingHooks timing; bool timingOn = false;
    cudaEvent_t evA = nullptr, evB = nullptr, evT0 = nullptr, evT1 = nullptr;
    // second lane of the offline wavefront (ohb_render): its own stream, fork / film-order / join events
    cudaStream_t stream2 = nullptr; cudaEvent_t evFork = nullptr, evFilm = nullptr, evJoin = nullptr;
};

#define OHB_FAIL(ctx, msg) do { (ctx)->err = (msg); return 1; } while (0)
#define CU(ctx, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e); return 1; } } while (0)

static int ensureFilm(ohb_ctx* c) {
    size_t n = size_t(c->W) * c->H;
    CU(c, c->accum.reserve(n * 16)); CU(c, c->ldr.reserve(n * 4)); CU(c, c->albedoAOV.reserve(n * 16)); CU(c, c->normalAOV.reserve(n * 16));
    if (c->profile == OHB_PROFILE_REALTIME) {
        CU(c, c->accumPrev.reserve(n * 16));
        for (int i = 0; i < 2; i++) { CU(c, c->surf[i].reserve(n * 16)); CU(c, c->shad[i].reserve(n * 16)); for (int k = 0; k < 3; k++) CU(c, c->res[i][k].reserve(n * 16)); }
    }
    return 0;
}
static int clearFilm(ohb_ctx* c) {
    size_t n = size_t(c->W) * c->H;
    CU(c, cudaMemsetAsync(c->accum.p, 0, n * 16, c->stream)); CU(c, cudaMemsetAsync(c->ldr.p, 0, n * 4, c->stream));
    CU(c, cudaMemsetAsync(c->albedoAOV.p, 0, n * 16, c->stream)); CU(c, cudaMemsetAsync(c->normalAOV.p, 0, n * 16, c->stream));
    if (c->profile == OHB_PROFILE_REALTIME) {
        CU(c, cudaMemsetAsync(c->accumPrev.p, 0, n * 16, c->stream));
        for (int i = 0; i < 2; i++) {
            CU(c, cudaMemsetAsync(c->surf[i].p, 0, n * 16, c->stream)); CU(c, cudaMemsetAsync(c->shad[i].p, 0, n * 16, c->stream));
            for (int k = 0; k < 3; k++) CU(c, cudaMemsetAsync(c->res[i][k].p, 0, n * 16, c->stream));
        }
    }
    return 0;
}
static void defaultSettings(ohb_ctx* c) {
    ohb_settings s{};
    if (c->profile == OHB_PROFILE_REALTIME) {   // kRealtimeRTSettings (rt_settings.hpp:37-48); sampler: Sobol (quirk Q3)
        s.profile = OHB_PROFILE_REALTIME; s.max_bounces = 2; s.flags = OHB_FLAG_ENABLE_AOVS | OHB_FLAG_ENABLE_INTERNAL_DENOISE | OHB_FLAG_ENABLE_FIREFLY_CLAMP;
        s.firefly_clamp_lum = 10.0f;
    } else {                                    // kOfflineRTSettings (rt_settings.hpp:49-61)
        s.profile = OHB_PROFILE_OFFLINE; s.max_bounces = 4; s.flags = OHB_FLAG_ENABLE_AOVS; s.firefly_clamp_lum = 0.0f;
    }
    s.sampler_type = OHB_SAMPLER_SOBOL; s.samples_per_frame = 1;
    c->settings = s;
}

extern "C" {

uint32_t ohb_abi_version(void) { return OHB_ABI_VERSION; }

const char* ohb_last_error(const ohb_ctx* c) { return c ? c->err.c_str() : g_createError.c_str(); }

ohb_ctx* ohb_create(int device_ordinal, uint32_t width, uint32_t height, int profile) {
    g_createError.clear();
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { g_createError = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)"; return nullptr; }
    if (device_ordinal < 0 || device_ordinal >= ndev) { g_createError = "device ordinal out of range"; return nullptr; }
    if (width == 0 || height == 0 || width > 65535u || height > 65535u) { g_createError = "bad resolution"; return nullptr; }
    if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) { g_createError = cudaGetErrorString(e); return nullptr; }
    ohb_ctx* c = new ohb_ctx();
    c->device = device_ordinal; c->W = width; c->H = height; c->profile = profile;
    c->tileW = width; c->tileH = height;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, device_ordinal); c->numSMs = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess || (e = uploadConstants()) != cudaSuccess) {
        g_createError = cudaGetErrorString(e); delete c; return nullptr;
    }
    cudaEventCreate(&c->evA); cudaEventCreate(&c->evB); cudaEventCreate(&c->evT0); cudaEventCreate(&c->evT1);
    defaultSettings(c);
    if (ensureFilm(c) || clearFilm(c) || c->smallCounters.reserve(64 * 4) != cudaSuccess || c->devCounters.reserve(16 * 8) != cudaSuccess) {
        g_createError = c->err.empty() ? "allocation failed" : c->err; ohb_destroy(c); return nullptr;
    }
    cudaMemsetAsync(c->smallCounters.p, 0, 64 * 4, c->stream); cudaMemsetAsync(c->devCounters.p, 0, 128, c->stream);
    // dummy 1-float CDFs like the reference when no env is loaded (light_upload.cpp:505-535)
    c->marg.reserve(4); c->cond.reserve(4); c->integral.reserve(4);
    cudaStreamSynchronize(c->stream);
    return c;
}

void ohb_destroy(ohb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf* all[] = {&c->positions, &c->indices, &c->normals, &c->uvs, &c->matIds, &c->triInst, &c->instXform, &c->instNormalMat, &c->instInv, &c->matColors,
                     &c->tex, &c->lights, &c->env, &c->marg, &c->cond, &c->rowTotal, &c->integral, &c->margTop, &c->condTop, &c->activeTris, &c->wtri, &c->primLo, &c->primHi, &c->boundsBits,
                     &c->keys, &c->vals, &c->keysTmp, &c->valsTmp, &c->sortTemp, &c->left, &c->right, &c->parentInner, &c->parentLeaf,
                     &c->nodeLo, &c->nodeHi, &c->visit, &c->wideCounters, &c->wideItemsA, &c->wideItemsB, &c->sah, &c->wnodes, &c->tris, &c->rayO, &c->rayD, &c->hit, &c->thr, &c->rad,
                     &c->pendA, &c->pendB, &c->meta, &c->fh0, &c->fh1, &c->fh2, &c->pay0, &c->pay1, &c->pay2, &c->pay3, &c->shO, &c->shD, &c->queueA, &c->queueB, &c->queueS, &c->hitFlag, &c->sobolTab, &c->smallCounters, &c->devCounters, &c->octPerm, &c->octHist, &c->octScanTmp,
                     &c->accum, &c->ldr, &c->albedoAOV, &c->normalAOV, &c->sampleDump, &c->hookRays, &c->hookOut, &c->hookAux,
                     &c->accumPrev, &c->surf[0], &c->surf[1], &c->shad[0], &c->shad[1], &c->res[0][0], &c->res[0][1], &c->res[0][2],
                     &c->res[1][0], &c->res[1][1], &c->res[1][2], &c->rtDump[0], &c->rtDump[1], &c->rtDump[2],
                     &c->blasLo, &c->blasHi, &c->blasInfo, &c->instOfPrim, &c->maxLevels, &c->tlasNodes, &c->tlasLeaves,
                     &c->tWtri, &c->tPrimLo, &c->tPrimHi, &c->tBounds, &c->tKeys, &c->tVals, &c->tKeysTmp, &c->tValsTmp, &c->tSortTemp, &c->tLeft, &c->tRight, &c->tParentInner, &c->tParentLeaf,
                     &c->tNodeLo, &c->tNodeHi, &c->tVisit, &c->tWideCounters, &c->tItemsA, &c->tItemsB, &c->tSah,
                     &c->motionAOV, &c->depthAOV, &c->svgfHistColor[0], &c->svgfHistColor[1], &c->svgfHistMoments[0], &c->svgfHistMoments[1],
                     &c->svgfHistGeom[0], &c->svgfHistGeom[1], &c->svgfColor[0], &c->svgfColor[1], &c->svgfVar[0], &c->svgfVar[1]};
    for (DevBuf* b : all) b->release();
    for (BuilderSet& bs : c->blasSets) bs.release();
    if (c->evA) cudaEventDestroy(c->evA);
    if (c->evB) cudaEventDestroy(c->evB);
    if (c->evT0) cudaEventDestroy(c->evT0);
    if (c->evT1) cudaEventDestroy(c->evT1);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evFilm) cudaEventDestroy(c->evFilm);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int ohb_resize(ohb_ctx* c, uint32_t width, uint32_t height) {
    if (!c) return 1;
    if (width == 0 || height == 0 || width > 65535u || height > 65535u) OHB_FAIL(c, "bad resolution");
    cudaSetDevice(c->device);
    c->W = width; c->H = height; c->tileX = c->tileY = 0; c->tileW = width; c->tileH = height;
    if (ensureFilm(c) || clearFilm(c)) return 1;
    ohb_reset_accumulation(c);
    return 0;
}

int ohb_set_geometry(ohb_ctx* c, const void* positions, size_t stride_bytes, uint32_t nverts, const uint32_t* idx, uint32_t ntris,
                     const float* normals, const float* uvs, const uint32_t* mat_ids) {
    if (!c) return 1;
    if (!positions || !idx || !normals || !uvs || !mat_ids) OHB_FAIL(c, "ohb_set_geometry: null array");
    if (stride_bytes < 12 || (stride_bytes & 3)) OHB_FAIL(c, "ohb_set_geometry: stride must be >= 12 and a multiple of 4");
    if (ntris >= (1u << 29)) OHB_FAIL(c, "ohb_set_geometry: too many triangles (leaf refs hold 29 bits)");
    cudaSetDevice(c->device);
    for (size_t i = 0; i < size_t(ntris) * 3; i++) if (idx[i] >= nverts) OHB_FAIL(c, "ohb_set_geometry: index out of range");
    uint32_t maxMat = 0u;
    for (size_t i = 0; i < ntris; i++) maxMat = std::max(maxMat, mat_ids[i]);
    c->maxMatId = maxMat;             // checked against the material count by ohb_set_materials / ohb_render (closest-hit reads matColors[matID*3..] unchecked)
    c->nverts = nverts; c->ntris = ntris; c->posStride = stride_bytes; c->accelValid = false;
    CU(c, c->positions.reserve(size_t(nverts) * stride_bytes)); CU(c, c->indices.reserve(size_t(ntris) * 12));
    CU(c, c->normals.reserve(size_t(nverts) * 16)); CU(c, c->uvs.reserve(size_t(nverts) * 8)); CU(c, c->matIds.reserve(size_t(ntris) * 4));
    CU(c, cudaMemcpyAsync(c->positions.p, positions, size_t(nverts) * stride_bytes, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->indices.p, idx, size_t(ntris) * 12, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->normals.p, normals, size_t(nverts) * 16, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->uvs.p, uvs, size_t(nverts) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->matIds.p, mat_ids, size_t(ntris) * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int ohb_set_instances(ohb_ctx* c, const ohb_instance* inst, uint32_t n) {
    if (!c) return 1;
    if (n && !inst) OHB_FAIL(c, "ohb_set_instances: null array");
    c->instances.assign(inst, inst + n); c->accelValid = false;
    return 0;
}

int ohb_set_materials(ohb_ctx* c, const float* mc, uint32_t nmat) {
    if (!c) return 1;
    if (!mc || !nmat) OHB_FAIL(c, "ohb_set_materials: empty");
    if (c->ntris && c->maxMatId >= nmat) OHB_FAIL(c, "ohb_set_materials: fewer material records than the geometry's material ids address");
    cudaSetDevice(c->device);
    c->nmat = nmat;
    CU(c, c->matColors.reserve(size_t(nmat) * 48));
    CU(c, cudaMemcpyAsync(c->matColors.p, mc, size_t(nmat) * 48, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int ohb_set_textures(ohb_ctx* c, const uint8_t* layers, uint32_t w, uint32_t h, uint32_t n) {
    if (!c) return 1;
    cudaSetDevice(c->device);
    if (!layers || !n) { c->texW = c->texH = c->texLayers = 0; return 0; }
    size_t bytes = size_t(w) * h * 4u * n;
    CU(c, c->tex.reserve(bytes));
    CU(c, cudaMemcpyAsync(c->tex.p, layers, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->texW = w; c->texH = h; c->texLayers = n;
    return 0;
}

int ohb_set_lights(ohb_ctx* c, const void* ssbo, size_t bytes) {
    if (!c) return 1;
    cudaSetDevice(c->device);
    if (!ssbo || bytes < 16) { c->lightCount = 0; c->envMapTexIdx = 0xFFFFFFFFu; c->envIntensity = 1.0f; return 0; }   // no light SSBO (quirk Q13)
    const uint8_t* b = static_cast<const uint8_t*>(ssbo);
    uint32_t cnt; memcpy(&cnt, b, 4); memcpy(&c->envMapTexIdx, b + 4, 4); memcpy(&c->envIntensity, b + 8, 4);
    uint32_t avail = uint32_t((bytes - 16) / 80);
    if (cnt > avail) OHB_FAIL(c, "ohb_set_lights: lightCount exceeds buffer");
    c->lightCount = cnt;
    if (cnt) {
        CU(c, c->lights.reserve(size_t(cnt) * 80));
        CU(c, cudaMemcpyAsync(c->lights.p, b + 16, size_t(cnt) * 80, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int ohb_set_env(ohb_ctx* c, const float* rgba, uint32_t w, uint32_t h) {
    if (!c) return 1;
    cudaSetDevice(c->device);
    if (!rgba || !w || !h) { c->envW = c->envH = 0; c->envIntegral = 0.0f; return 0; }
    size_t n = size_t(w) * h;
    CU(c, c->env.reserve(n * 16)); CU(c, c->cond.reserve(n * 4)); CU(c, c->marg.reserve(size_t(h) * 4)); CU(c, c->rowTotal.reserve(size_t(h) * 4)); CU(c, c->integral.reserve(4));
    CU(c, cudaMemcpyAsync(c->env.p, rgba, n * 16, cudaMemcpyHostToDevice, c->stream));
    static const bool blockedOn = []() { const char* e = getenv("OHB_ENV_BLOCKED"); return e ? atoi(e) != 0 : true; }();
    c->envTops = blockedOn && (w % 32u) == 0u && (h % 32u) == 0u;      // the blocked CDF search needs whole 32-entry blocks
    if (c->envTops) { CU(c, c->condTop.reserve(size_t(h) * (w / 32) * 4)); CU(c, c->margTop.reserve(size_t(h / 32) * 4)); }
    launchEnvCdf(c->env.as<f4>(), w, h, c->cond.as<float>(), c->marg.as<float>(), c->rowTotal.as<float>(), c->integral.as<float>(),
                 c->envTops ? c->condTop.as<float>() : nullptr, c->envTops ? c->margTop.as<float>() : nullptr, c->stream, &c->launches);
    CU(c, cudaMemcpyAsync(&c->envIntegral, c->integral.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->envW = w; c->envH = h;
    return 0;
}

int ohb_get_env_cdf(ohb_ctx* c, float* marg, float* cond, float* integral) {
    if (!c) return 1;
    if (!c->envW) OHB_FAIL(c, "ohb_get_env_cdf: no environment map");
    cudaSetDevice(c->device);
    if (marg) CU(c, cudaMemcpyAsync(marg, c->marg.p, size_t(c->envH) * 4, cudaMemcpyDeviceToHost, c->stream));
    if (cond) CU(c, cudaMemcpyAsync(cond, c->cond.p, size_t(c->envW) * c->envH * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (integral) *integral = c->envIntegral;
    return 0;
}

static void fillScene(ohb_ctx* c, SceneDev& s) {
    memset(&s, 0, sizeof(s));
    s.wnodes = c->wnodes.as<u4>(); s.tris = c->tris.as<f4>(); s.numTris = c->accelValid ? c->numActive : 0u; s.numWideNodes = c->accelValid ? c->stats.num_nodes : 0u;
    s.twoLevel = (c->accelValid && c->accelMode == OHB_ACCEL_TWO_LEVEL) ? 1u : 0u;
    s.tlasNodes = c->tlasNodes.as<u4>(); s.tlasLeaves = c->tlasLeaves.as<f4>(); s.blasInfo = c->blasInfo.as<u4>();
    if (s.twoLevel) s.numWideNodes = 0u;
    s.indices = c->indices.as<uint32_t>(); s.normals = c->normals.as<f4>(); s.uvs = c->uvs.as<f2>(); s.matIds = c->matIds.as<uint32_t>();
    s.triInst = c->triInst.as<uint32_t>(); s.instNormalMat = c->instNormalMat.as<f4>(); s.instInv = c->instInv.as<f4>();
    s.matColors = c->matColors.as<f4>();
    s.tex = c->tex.as<uint8_t>(); s.texW = c->texW; s.texH = c->texH; s.texLayers = c->texLayers;
    s.lights = c->lights.as<GPULight>(); s.lightCount = c->lightCount; s.envMapTexIdx = c->envMapTexIdx; s.envIntensity = c->envIntensity;
    s.env = c->envW ? c->env.as<f4>() : nullptr; s.envW = c->envW; s.envH = c->envH;
    s.marg = c->marg.as<float>(); s.cond = c->cond.as<float>(); s.envIntegral = c->envIntegral;
    s.margTop = (c->envW && c->envTops) ? c->margTop.as<float>() : nullptr; s.condTop = (c->envW && c->envTops) ? c->condTop.as<float>() : nullptr;
    if (!c->envW) s.envMapTexIdx = 0xFFFFFFFFu;
}

int ohb_env_sample_batch(ohb_ctx* c, const float* u12, uint32_t n, float* dir_pdf, float* pdf_of_dir) {
    if (!c) return 1;
    if (!c->envW) OHB_FAIL(c, "ohb_env_sample_batch: no environment map");
    if (!u12 || !dir_pdf) OHB_FAIL(c, "ohb_env_sample_batch: null array");
    if (n == 0) return 0;
    cudaSetDevice(c->device);
    CU(c, c->hookRays.reserve(size_t(n) * 8)); CU(c, c->hookOut.reserve(size_t(n) * 16)); CU(c, c->hookAux.reserve(size_t(n) * 4));
    CU(c, cudaMemcpyAsync(c->hookRays.p, u12, size_t(n) * 8, cudaMemcpyHostToDevice, c->stream));
    SceneDev s; fillScene(c, s);
    launchEnvSample(s, c->hookRays.as<float>(), n, c->hookOut.as<f4>(), c->hookAux.as<float>(), c->stream, &c->launches);
    CU(c, cudaMemcpyAsync(dir_pdf, c->hookOut.p, size_t(n) * 16, cudaMemcpyDeviceToHost, c->stream));
    if (pdf_of_dir) CU(c, cudaMemcpyAsync(pdf_of_dir, c->hookAux.p, size_t(n) * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int ohb_env_pdf_batch(ohb_ctx* c, const float* dirs3, uint32_t n, float* pdf) {
    if (!c) return 1;
    if (!c->envW) OHB_FAIL(c, "ohb_env_pdf_batch: no environment map");
    if (!dirs3 || !pdf) OHB_FAIL(c, "ohb_env_pdf_batch: null array");
    if (n == 0) return 0;
    cudaSetDevice(c->device);
    CU(c, c->hookRays.reserve(size_t(n) * 12)); CU(c, c->hookAux.reserve(size_t(n) * 4));
    CU(c, cudaMemcpyAsync(c->hookRays.p, dirs3, size_t(n) * 12, cudaMemcpyHostToDevice, c->stream));
    SceneDev s; fillScene(c, s);
    launchEnvPdf(s, c->hookRays.as<float>(), n, c->hookAux.as<float>(), c->stream, &c->launches);
    CU(c, cudaMemcpyAsync(pdf, c->hookAux.p, size_t(n) * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

static int hybridUpload(ohb_ctx* c, const float* pos, const float* nrm, const float* albedo, const float* hist, const float* instMat, uint32_t ninst, float** dev) {
    const size_t n = size_t(c->W) * c->H;
    CU(c, c->hookRays.reserve(n * (16 + 8 + 16 + 16) + size_t(ninst) * 16 + 64));
    float* base = c->hookRays.as<float>();
    dev[0] = base; dev[1] = base + n * 4; dev[2] = dev[1] + n * 2; dev[3] = dev[2] + n * 4; dev[4] = dev[3] + n * 4;
    CU(c, cudaMemcpyAsync(dev[0], pos, n * 16, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(dev[1], nrm, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (albedo) CU(c, cudaMemcpyAsync(dev[2], albedo, n * 16, cudaMemcpyHostToDevice, c->stream));
    if (hist) CU(c, cudaMemcpyAsync(dev[3], hist, n * 16, cudaMemcpyHostToDevice, c->stream));
    if (instMat) CU(c, cudaMemcpyAsync(dev[4], instMat, size_t(ninst) * 16, cudaMemcpyHostToDevice, c->stream));
    return 0;
}
int ohb_hybrid_shadow(ohb_ctx* c, const float* pos, const float* nrm, const ohb_hybrid_shadow_params* p, uint8_t* mask) {
    if (!c) return 1;
    if (!pos || !nrm || !p || !mask) OHB_FAIL(c, "ohb_hybrid_shadow: null argument");
    if (!c->accelValid) OHB_FAIL(c, "ohb_hybrid_shadow: no acceleration structure");
    cudaSetDevice(c->device);
    float* dev[5]; if (hybridUpload(c, pos, nrm, nullptr, nullptr, nullptr, 0, dev)) return 1;
    const size_t n = size_t(c->W) * c->H;
    CU(c, c->hookOut.reserve(n));
    HybridShadowParams pc{}; pc.lightDir = mk3(p->light_dir[0], p->light_dir[1], p->light_dir[2]); pc.lightRadius = p->light_radius;
    pc.lightPos = mk3(p->light_pos[0], p->light_pos[1], p->light_pos[2]); pc.lightRange = p->light_range; pc.W = c->W; pc.H = c->H; pc.lightType = p->light_type; pc.sampleCount = p->sample_count;
    SceneDev s; fillScene(c, s);
    launchHybridShadow(s, pc, reinterpret_cast<const f4*>(dev[0]), reinterpret_cast<const f2*>(dev[1]), c->hookOut.as<uint8_t>(), c->stream, &c->launches);
    CU(c, cudaMemcpyAsync(mask, c->hookOut.p, n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream)); CU(c, cudaGetLastError());
    return 0;
}
int ohb_hybrid_gi(ohb_ctx* c, const float* pos, const float* nrm, const float* albedo, const float* hist, const float* instMat, uint32_t ninst, const ohb_hybrid_gi_params* p, uint16_t* out) {
    if (!c) return 1;
    if (!pos || !nrm || !albedo || !hist || !instMat || !p || !out) OHB_FAIL(c, "ohb_hybrid_gi: null argument");
    if (!c->accelValid) OHB_FAIL(c, "ohb_hybrid_gi: no acceleration structure");
    if (ninst < c->instances.size()) OHB_FAIL(c, "ohb_hybrid_gi: fewer instance materials than instances");
    cudaSetDevice(c->device);
    float* dev[5]; if (hybridUpload(c, pos, nrm, albedo, hist, instMat, ninst, dev)) return 1;
    const size_t n = size_t(c->W) * c->H;
    CU(c, c->hookOut.reserve(n * 8));
    HybridGiParams pc{}; pc.lightPos = mk3(p->light_pos[0], p->light_pos[1], p->light_pos[2]); pc.lightIntensity = p->light_intensity; pc.W = c->W; pc.H = c->H;
    pc.sampleCount = p->sample_count; pc.frameIndex = p->frame_index;
    SceneDev s; fillScene(c, s);
    launchHybridGi(s, pc, reinterpret_cast<const f4*>(dev[0]), reinterpret_cast<const f2*>(dev[1]), reinterpret_cast<const f4*>(dev[2]), reinterpret_cast<const f4*>(dev[3]),
                   reinterpret_cast<const f4*>(dev[4]), c->hookOut.as<h4>(), c->stream, &c->launches);
    CU(c, cudaMemcpyAsync(out, c->hookOut.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream)); CU(c, cudaGetLastError());
    return 0;
}

int ohb_nrd_pack_batch(ohb_ctx* c, const float* rad_hd_vz_rough, const float* normal_rough, uint32_t n, float* packed_radiance, float* packed_normal, float* unpacked_rgb) {
    if (!c) return 1;
    if (!rad_hd_vz_rough || !normal_rough || !packed_radiance || !packed_normal || !unpacked_rgb) OHB_FAIL(c, "ohb_nrd_pack_batch: null array");
    if (n == 0) return 0;
    cudaSetDevice(c->device);
    CU(c, c->hookRays.reserve(size_t(n) * 40)); CU(c, c->hookOut.reserve(size_t(n) * 32)); CU(c, c->hookAux.reserve(size_t(n) * 12));
    float* in6 = c->hookRays.as<float>(); float* nr4 = in6 + size_t(n) * 6;
    CU(c, cudaMemcpyAsync(in6, rad_hd_vz_rough, size_t(n) * 24, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(nr4, normal_rough, size_t(n) * 16, cudaMemcpyHostToDevice, c->stream));
    f4* pr = c->hookOut.as<f4>(); f4* pn = pr + n;
    launchNrdPack(in6, nr4, n, pr, pn, c->hookAux.as<float>(), c->stream, &c->launches);
    CU(c, cudaMemcpyAsync(packed_radiance, pr, size_t(n) * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(packed_normal, pn, size_t(n) * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(unpacked_rgb, c->hookAux.p, size_t(n) * 12, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaGetLastError());
    return 0;
}

// instance tables: object->world rows, normal matrix = transpose(inverse(mat3)) (pt_closesthit.rchit:66), world->object rows
static void instanceTables(const std::vector<ohb_instance>& insts, std::vector<float>& xf, std::vector<float>& nm, std::vector<float>& iv) {
    const size_t ni = insts.size();
    xf.assign(ni * 12 + 12, 0.0f); nm.assign(ni * 12 + 12, 0.0f); iv.assign(ni * 12 + 12, 0.0f);
    for (size_t i = 0; i < ni; i++) {
        const float* a = insts[i].xform;
        memcpy(&xf[i * 12], a, 48);
        float co[3][3];
        co[0][0] = a[5] * a[10] - a[6] * a[9];  co[0][1] = a[6] * a[8] - a[4] * a[10]; co[0][2] = a[4] * a[9] - a[5] * a[8];
        co[1][0] = a[2] * a[9] - a[1] * a[10];  co[1][1] = a[0] * a[10] - a[2] * a[8]; co[1][2] = a[1] * a[8] - a[0] * a[9];
        co[2][0] = a[1] * a[6] - a[2] * a[5];   co[2][1] = a[2] * a[4] - a[0] * a[6];  co[2][2] = a[0] * a[5] - a[1] * a[4];
        float det = a[0] * co[0][0] + a[1] * co[0][1] + a[2] * co[0][2];
        float id = 1.0f / det;
        for (int r = 0; r < 3; r++) {
            for (int k = 0; k < 3; k++) nm[i * 12 + r * 4 + k] = co[r][k] * id;
            nm[i * 12 + r * 4 + 3] = 0.0f;
        }
        for (int r = 0; r < 3; r++) {
            float* row = &iv[i * 12 + r * 4];
            for (int k = 0; k < 3; k++) row[k] = nm[i * 12 + k * 4 + r];
            row[3] = -(row[0] * a[3] + row[1] * a[7] + row[2] * a[11]);
        }
    }
}
static int uploadInstanceTables(ohb_ctx* c) {
    std::vector<float> xf, nm, iv; instanceTables(c->instances, xf, nm, iv);
    CU(c, c->instXform.reserve(xf.size() * 4)); CU(c, c->instNormalMat.reserve(nm.size() * 4)); CU(c, c->instInv.reserve(iv.size() * 4));
    CU(c, cudaMemcpyAsync(c->instXform.p, xf.data(), xf.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->instNormalMat.p, nm.data(), nm.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->instInv.p, iv.data(), iv.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));        // the staging vectors die here
    return 0;
}
// BuildArrays over the context's main builder buffers (sized for `cap` primitives) — the flattened tree, or one BLAS at a time
static int mainBuildArrays(ohb_ctx* c, size_t cap, BuildArrays& b) {
    const size_t nn = cap;
    CU(c, c->wtri.reserve(nn * 48)); CU(c, c->primLo.reserve(nn * 16)); CU(c, c->primHi.reserve(nn * 16)); CU(c, c->boundsBits.reserve(32));
    CU(c, c->keys.reserve(nn * 8)); CU(c, c->vals.reserve(nn * 4)); CU(c, c->keysTmp.reserve(nn * 8)); CU(c, c->valsTmp.reserve(nn * 4));
    CU(c, c->sortTemp.reserve(size_t(radixSortTempWords(uint32_t(cap))) * 4));
    CU(c, c->left.reserve(nn * 4)); CU(c, c->right.reserve(nn * 4)); CU(c, c->parentInner.reserve(nn * 4)); CU(c, c->parentLeaf.reserve(nn * 4));
    CU(c, c->nodeLo.reserve(nn * 16)); CU(c, c->nodeHi.reserve(nn * 16));
    CU(c, c->visit.reserve(nn * 4)); CU(c, c->wideCounters.reserve(16)); CU(c, c->sah.reserve(8));
    CU(c, c->wideItemsA.reserve((nn / 4 + 2) * sizeof(WideItem))); CU(c, c->wideItemsB.reserve((nn / 4 + 2) * sizeof(WideItem)));
    memset(&b, 0, sizeof(b));
    b.positions = c->positions.as<uint8_t>(); b.posStride = c->posStride; b.indices = c->indices.as<uint32_t>(); b.triInst = c->triInst.as<uint32_t>();
    b.instXform = c->instXform.as<f4>(); b.activeTris = c->activeTris.as<uint32_t>(); b.n = uint32_t(cap);
    b.wtri = c->wtri.as<f4>(); b.primLo = c->primLo.as<f4>(); b.primHi = c->primHi.as<f4>(); b.boundsBits = c->boundsBits.as<uint32_t>();
    b.keys = c->keys.as<uint64_t>(); b.vals = c->vals.as<uint32_t>();
    b.left = c->left.as<int32_t>(); b.right = c->right.as<int32_t>(); b.parentInner = c->parentInner.as<int32_t>(); b.parentLeaf = c->parentLeaf.as<int32_t>();
    b.nodeLo = c->nodeLo.as<f4>(); b.nodeHi = c->nodeHi.as<f4>();
    b.visit = c->visit.as<uint32_t>(); b.wideCounters = c->wideCounters.as<uint32_t>(); b.sah = c->sah.as<float>();
    b.wnodes = c->wnodes.as<u4>(); b.tris = c->tris.as<f4>();
    return 0;
}
static int tlasBuildArrays(ohb_ctx* c, uint32_t np, BuildArrays& b) {
    const size_t nn = np;
    CU(c, c->tWtri.reserve(nn * 48)); CU(c, c->tPrimLo.reserve(nn * 16)); CU(c, c->tPrimHi.reserve(nn * 16)); CU(c, c->tBounds.reserve(32));
    CU(c, c->tKeys.reserve(nn * 8)); CU(c, c->tVals.reserve(nn * 4)); CU(c, c->tKeysTmp.reserve(nn * 8)); CU(c, c->tValsTmp.reserve(nn * 4));
    CU(c, c->tSortTemp.reserve(size_t(radixSortTempWords(np)) * 4));
    CU(c, c->tLeft.reserve(nn * 4)); CU(c, c->tRight.reserve(nn * 4)); CU(c, c->tParentInner.reserve(nn * 4)); CU(c, c->tParentLeaf.reserve(nn * 4));
    CU(c, c->tNodeLo.reserve(nn * 16)); CU(c, c->tNodeHi.reserve(nn * 16));
    CU(c, c->tVisit.reserve(nn * 4)); CU(c, c->tWideCounters.reserve(16)); CU(c, c->tSah.reserve(8));
    CU(c, c->tItemsA.reserve((nn / 4 + 2) * sizeof(WideItem))); CU(c, c->tItemsB.reserve((nn / 4 + 2) * sizeof(WideItem)));
    CU(c, c->tlasNodes.reserve((nn + 1) * 16 * OHB_WNODE_VECS)); CU(c, c->tlasLeaves.reserve((nn + 1) * 48));
    memset(&b, 0, sizeof(b));
    b.instXform = c->instXform.as<f4>(); b.n = np;
    b.wtri = c->tWtri.as<f4>(); b.primLo = c->tPrimLo.as<f4>(); b.primHi = c->tPrimHi.as<f4>(); b.boundsBits = c->tBounds.as<uint32_t>();
    b.keys = c->tKeys.as<uint64_t>(); b.vals = c->tVals.as<uint32_t>();
    b.left = c->tLeft.as<int32_t>(); b.right = c->tRight.as<int32_t>(); b.parentInner = c->tParentInner.as<int32_t>(); b.parentLeaf = c->tParentLeaf.as<int32_t>();
    b.nodeLo = c->tNodeLo.as<f4>(); b.nodeHi = c->tNodeHi.as<f4>();
    b.visit = c->tVisit.as<uint32_t>(); b.wideCounters = c->tWideCounters.as<uint32_t>(); b.sah = c->tSah.as<float>();
    b.wnodes = c->tlasNodes.as<u4>(); b.tris = c->tlasLeaves.as<f4>();
    return 0;
}
static int setBuildArrays(ohb_ctx* c, BuilderSet& q, size_t cap, BuildArrays& b) {
    const size_t nn = cap;
    CU(c, q.wtri.reserve(nn * 48)); CU(c, q.primLo.reserve(nn * 16)); CU(c, q.primHi.reserve(nn * 16)); CU(c, q.boundsBits.reserve(32));
    CU(c, q.keys.reserve(nn * 8)); CU(c, q.vals.reserve(nn * 4)); CU(c, q.keysTmp.reserve(nn * 8)); CU(c, q.valsTmp.reserve(nn * 4));
    CU(c, q.sortTemp.reserve(size_t(radixSortTempWords(uint32_t(cap))) * 4));
    CU(c, q.left.reserve(nn * 4)); CU(c, q.right.reserve(nn * 4)); CU(c, q.parentInner.reserve(nn * 4)); CU(c, q.parentLeaf.reserve(nn * 4));
    CU(c, q.nodeLo.reserve(nn * 16)); CU(c, q.nodeHi.reserve(nn * 16));
    CU(c, q.visit.reserve(nn * 4)); CU(c, q.wideCounters.reserve(16)); CU(c, q.sah.reserve(8));
    CU(c, q.itemsA.reserve((nn / 4 + 2) * sizeof(WideItem))); CU(c, q.itemsB.reserve((nn / 4 + 2) * sizeof(WideItem)));
    if (!q.stream) { CU(c, cudaStreamCreateWithFlags(&q.stream, cudaStreamNonBlocking)); CU(c, cudaEventCreateWithFlags(&q.done, cudaEventDisableTiming)); }
    memset(&b, 0, sizeof(b));
    b.positions = c->positions.as<uint8_t>(); b.posStride = c->posStride; b.indices = c->indices.as<uint32_t>(); b.triInst = c->triInst.as<uint32_t>();
    b.instXform = c->instXform.as<f4>(); b.activeTris = c->activeTris.as<uint32_t>(); b.n = uint32_t(cap);
    b.wtri = q.wtri.as<f4>(); b.primLo = q.primLo.as<f4>(); b.primHi = q.primHi.as<f4>(); b.boundsBits = q.boundsBits.as<uint32_t>();
    b.keys = q.keys.as<uint64_t>(); b.vals = q.vals.as<uint32_t>();
    b.left = q.left.as<int32_t>(); b.right = q.right.as<int32_t>(); b.parentInner = q.parentInner.as<int32_t>(); b.parentLeaf = q.parentLeaf.as<int32_t>();
    b.nodeLo = q.nodeLo.as<f4>(); b.nodeHi = q.nodeHi.as<f4>();
    b.visit = q.visit.as<uint32_t>(); b.wideCounters = q.wideCounters.as<uint32_t>(); b.sah = q.sah.as<float>();
    b.wnodes = c->wnodes.as<u4>(); b.tris = c->tris.as<f4>();
    return 0;
}
static uint32_t treeletPassesKnob() { static const uint32_t v = []() { const char* e = getenv("OHB_TREELET_PASSES"); return e ? uint32_t(strtoul(e, nullptr, 10)) : 3u; }(); return v; }

int ohb_set_accel_mode(ohb_ctx* c, int mode) {
    if (!c) return 1;
    if (mode != OHB_ACCEL_FLATTEN && mode != OHB_ACCEL_TWO_LEVEL) OHB_FAIL(c, "ohb_set_accel_mode: unknown mode");
    if (mode != c->accelMode) { c->accelMode = mode; c->accelValid = false; }
    return 0;
}

int ohb_build_accel(ohb_ctx* c) {
    if (!c) return 1;
    cudaSetDevice(c->device);
    if (!c->ntris) OHB_FAIL(c, "ohb_build_accel: no geometry");
    const bool twoLevel = c->accelMode == OHB_ACCEL_TWO_LEVEL;
    uint32_t ni = uint32_t(c->instances.size());
    std::vector<uint32_t> triInst(c->ntris, 0xFFFFFFFFu), active;
    struct Range { uint32_t inst, first, count; }; std::vector<Range> ranges;     // per visible, non-empty instance: its slice of `active`
    for (uint32_t i = 0; i < ni; i++) {
        const ohb_instance& in = c->instances[i];
        if (uint64_t(in.first_tri) + in.tri_count > c->ntris) OHB_FAIL(c, "ohb_build_accel: instance triangle range out of bounds");
        if ((in.mask & 0xFFu) == 0u) continue;   // invisible to cullMask 0xFF
        const uint32_t first = uint32_t(active.size());
        for (uint32_t t = in.first_tri; t < in.first_tri + in.tri_count; t++) {
            if (triInst[t] == 0xFFFFFFFFu) active.push_back(t);
            else if (twoLevel) OHB_FAIL(c, "ohb_build_accel: two-level mode needs disjoint triangle ranges per instance (one BLAS per actor, rt_build.cpp:851-897)");
            triInst[t] = i;
        }
        if (active.size() > first) ranges.push_back({i, first, uint32_t(active.size()) - first});
    }
    uint32_t n = uint32_t(active.size());
    c->numActive = n; c->numTlasPrims = 0;
    if (uploadInstanceTables(c)) return 1;
    CU(c, c->triInst.reserve(size_t(c->ntris) * 4));
    CU(c, cudaMemcpyAsync(c->triInst.p, triInst.data(), size_t(c->ntris) * 4, cudaMemcpyHostToDevice, c->stream));
    memset(&c->stats, 0, sizeof(c->stats));
    c->stats.num_tris = n;
    if (n == 0) { c->accelValid = true; CU(c, cudaStreamSynchronize(c->stream)); return 0; }
    size_t nn = n;
    CU(c, c->activeTris.reserve(nn * 4)); CU(c, cudaMemcpyAsync(c->activeTris.p, active.data(), nn * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, c->wnodes.reserve(nn * 16 * OHB_WNODE_VECS)); CU(c, c->tris.reserve(nn * 48));     // a wide node is a binary node with > 3 triangles: fewer than n
    const uint32_t treeletPasses = treeletPassesKnob();
    c->stats.treelet_passes = treeletPasses;
    if (!twoLevel) {
        BuildArrays b; if (mainBuildArrays(c, nn, b)) return 1;
        CU(c, cudaEventRecord(c->evA, c->stream));
        launchBuild(b, c->keysTmp.as<uint64_t>(), c->valsTmp.as<uint32_t>(), c->sortTemp.as<uint32_t>(), c->wideItemsA.as<WideItem>(), c->wideItemsB.as<WideItem>(),
                    treeletPasses, c->stream, &c->launches);
        CU(c, cudaEventRecord(c->evB, c->stream));
        uint32_t wc[4] = {1, 0, 0, 0}; float sah[2] = {0, 0}; f4 rootLo{}, rootHi{};
        CU(c, cudaMemcpyAsync(wc, c->wideCounters.p, 16, cudaMemcpyDeviceToHost, c->stream));
        if (n >= 2) {
            CU(c, cudaMemcpyAsync(&rootLo, c->nodeLo.p, 16, cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaMemcpyAsync(&rootHi, c->nodeHi.p, 16, cudaMemcpyDeviceToHost, c->stream));
        }
        CU(c, cudaMemcpyAsync(sah, c->sah.p, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        CU(c, cudaGetLastError());
        if (wc[3] > OHB_MAX_LEVELS) OHB_FAIL(c, "ohb_build_accel: the 8-wide BVH is deeper than OHB_MAX_LEVELS (degenerate geometry: too many coincident triangles)");
        c->accelValid = true;
        cudaEventElapsedTime(&c->stats.build_ms, c->evA, c->evB);
        c->stats.num_nodes = wc[0];
        c->stats.max_leaf_tris = OHB_MAX_LEAF;
        c->stats.levels = wc[3] + 1u;
        float ra = boxArea(xyz(rootLo), xyz(rootHi));
        c->stats.sah_cost = (n >= 2 && ra > 0.0f) ? (sah[0] + sah[1]) / ra : float(n);
        return 0;
    }
    // ---- two-level: one object-space BLAS per instance (createBLAS, rt_acceleration_structure.cpp:205-405) + the TLAS (buildTLAS, :419-535) ----
    uint32_t maxCount = 0; for (const Range& r : ranges) maxCount = std::max(maxCount, r.count);
    const uint32_t np = uint32_t(ranges.size());
    CU(c, c->blasLo.reserve(size_t(ni) * 16 + 16)); CU(c, c->blasHi.reserve(size_t(ni) * 16 + 16)); CU(c, c->blasInfo.reserve(size_t(ni) * 16 + 16));
    CU(c, c->instOfPrim.reserve(size_t(np) * 4 + 4)); CU(c, c->maxLevels.reserve(4));
    std::vector<uint32_t> info(size_t(ni) * 4, 0u), iop(np);
    for (uint32_t k = 0; k < np; k++) { const Range& r = ranges[k]; info[size_t(r.inst) * 4 + 0] = r.first; info[size_t(r.inst) * 4 + 1] = r.first; info[size_t(r.inst) * 4 + 2] = r.count; iop[k] = r.inst; }
    CU(c, cudaMemcpyAsync(c->blasInfo.p, info.data(), info.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->instOfPrim.p, iop.data(), iop.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemsetAsync(c->maxLevels.p, 0, 4, c->stream));
    // BLAS builds are chains of small, latency-bound kernels: run them on OHB_BLAS_STREAMS streams (default 8), each with its own
    // builder scratch, so that several are in flight; everything joins the context's stream before the TLAS is built
    static const uint32_t kStreams = []() { const char* e = getenv("OHB_BLAS_STREAMS"); uint32_t v = e ? uint32_t(strtoul(e, nullptr, 10)) : 8u; return std::min(std::max(v, 1u), 32u); }();
    const uint32_t ns = std::min<uint32_t>(kStreams, np);
    if (c->blasSets.size() < ns) c->blasSets.resize(ns);
    std::vector<BuildArrays> sets(ns);
    for (uint32_t k = 0; k < ns; k++) if (setBuildArrays(c, c->blasSets[k], maxCount, sets[k])) return 1;
    CU(c, cudaEventRecord(c->evA, c->stream));
    for (uint32_t k = 0; k < ns; k++) CU(c, cudaStreamWaitEvent(c->blasSets[k].stream, c->evA, 0));       // the uploads above are on the context's stream
    for (uint32_t j = 0; j < np; j++) {
        const Range& r = ranges[j]; BuilderSet& q = c->blasSets[j % ns];
        BuildArrays bi = sets[j % ns];
        bi.objectSpace = 1u; bi.activeTris = c->activeTris.as<uint32_t>() + r.first; bi.n = r.count;
        bi.wnodes = c->wnodes.as<u4>() + size_t(r.first) * OHB_WNODE_VECS; bi.tris = c->tris.as<f4>() + size_t(r.first) * 3u;   // node / triangle capacity of a BLAS = its triangle count
        launchBuildBlas(bi, q.keysTmp.as<uint64_t>(), q.valsTmp.as<uint32_t>(), q.sortTemp.as<uint32_t>(), q.itemsA.as<WideItem>(), q.itemsB.as<WideItem>(), treeletPasses,
                        c->blasLo.as<f4>(), c->blasHi.as<f4>(), r.inst, c->maxLevels.as<uint32_t>(), q.stream, &c->launches);
    }
    for (uint32_t k = 0; k < ns; k++) { CU(c, cudaEventRecord(c->blasSets[k].done, c->blasSets[k].stream)); CU(c, cudaStreamWaitEvent(c->stream, c->blasSets[k].done, 0)); }
    BuildArrays tb; if (tlasBuildArrays(c, np, tb)) return 1;
    launchBuildTlas(tb, c->blasLo.as<f4>(), c->blasHi.as<f4>(), c->instOfPrim.as<uint32_t>(), false, c->tKeysTmp.as<uint64_t>(), c->tValsTmp.as<uint32_t>(), c->tSortTemp.as<uint32_t>(),
                    c->tItemsA.as<WideItem>(), c->tItemsB.as<WideItem>(), c->stream, &c->launches);
    CU(c, cudaEventRecord(c->evB, c->stream));
    uint32_t twc[4] = {1, 0, 0, 0}, blasLevels = 0;
    CU(c, cudaMemcpyAsync(twc, c->tWideCounters.p, 16, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(&blasLevels, c->maxLevels.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaGetLastError());
    // one stack: TLAS groups + sentinel + BLAS groups + postponed triangle groups
    if (blasLevels > OHB_MAX_LEVELS || twc[3] > OHB_MAX_LEVELS || blasLevels + twc[3] + 4u + OHB_POSTPONE_SLOTS > OHB_STACK_SIZE)
        OHB_FAIL(c, "ohb_build_accel: TLAS + BLAS depth exceeds the traversal stack");
    c->numTlasPrims = np;
    c->accelValid = true;
    cudaEventElapsedTime(&c->stats.build_ms, c->evA, c->evB);
    c->stats.num_nodes = twc[0];                 // TLAS nodes; the BLAS node counts stay on the device
    c->stats.max_leaf_tris = OHB_MAX_LEAF;
    c->stats.levels = twc[3] + 1u + blasLevels + 1u;
    c->stats.sah_cost = 0.0f;
    return 0;
}

// MODE_UPDATE (rt_acceleration_structure.cpp:503-512): same instances, new object->world transforms.  Flattened structure:
// triangles re-transformed, boxes refit bottom-up, wide nodes re-emitted (no sort / hierarchy / treelets).  Two-level: only
// the TLAS is refit — the object-space BLASes do not change.
int ohb_update_instances(ohb_ctx* c, const ohb_instance* inst, uint32_t n) {
    if (!c) return 1;
    if (!c->accelValid) OHB_FAIL(c, "ohb_update_instances: no acceleration structure to update (call ohb_build_accel)");
    if (!inst || n != c->instances.size()) OHB_FAIL(c, "ohb_update_instances: the instance count must not change (rebuild instead)");
    for (uint32_t i = 0; i < n; i++)
        if (inst[i].first_tri != c->instances[i].first_tri || inst[i].tri_count != c->instances[i].tri_count || (inst[i].mask & 0xFFu) != (c->instances[i].mask & 0xFFu))
            OHB_FAIL(c, "ohb_update_instances: only transforms may change (rebuild instead)");
    cudaSetDevice(c->device);
    for (uint32_t i = 0; i < n; i++) memcpy(c->instances[i].xform, inst[i].xform, 48);
    if (uploadInstanceTables(c)) return 1;
    if (c->numActive == 0) return 0;
    CU(c, cudaEventRecord(c->evA, c->stream));
    if (c->accelMode == OHB_ACCEL_TWO_LEVEL) {
        BuildArrays tb; if (tlasBuildArrays(c, c->numTlasPrims, tb)) return 1;
        launchBuildTlas(tb, c->blasLo.as<f4>(), c->blasHi.as<f4>(), c->instOfPrim.as<uint32_t>(), true, c->tKeysTmp.as<uint64_t>(), c->tValsTmp.as<uint32_t>(), c->tSortTemp.as<uint32_t>(),
                        c->tItemsA.as<WideItem>(), c->tItemsB.as<WideItem>(), c->stream, &c->launches);
    } else {
        BuildArrays b; if (mainBuildArrays(c, c->numActive, b)) return 1;
        launchRefit(b, c->wideItemsA.as<WideItem>(), c->wideItemsB.as<WideItem>(), c->stream, &c->launches);
    }
    CU(c, cudaEventRecord(c->evB, c->stream));
    uint32_t wc[4] = {1, 0, 0, 0};
    CU(c, cudaMemcpyAsync(wc, c->accelMode == OHB_ACCEL_TWO_LEVEL ? c->tWideCounters.p : c->wideCounters.p, 16, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaGetLastError());
    if (wc[3] > OHB_MAX_LEVELS) { c->accelValid = false; OHB_FAIL(c, "ohb_update_instances: refit tree too deep"); }
    cudaEventElapsedTime(&c->stats.update_ms, c->evA, c->evB);
    c->stats.num_nodes = wc[0];
    return 0;
}

int ohb_get_accel_stats(ohb_ctx* c, ohb_accel_stats* s) { if (!c || !s) return 1; *s = c->stats; return 0; }

int ohb_set_settings(ohb_ctx* c, const ohb_settings* s) {
    if (!c || !s) return 1;
    if ((s->max_bounces & 0xFFFFu) > 15u) OHB_FAIL(c, "ohb_set_settings: max_bounces > 15");
    if (s->profile != OHB_PROFILE_OFFLINE && s->profile != OHB_PROFILE_REALTIME) OHB_FAIL(c, "ohb_set_settings: unknown profile");
    if (s->sampler_type != OHB_SAMPLER_PCG && s->sampler_type != OHB_SAMPLER_SOBOL) OHB_FAIL(c, "ohb_set_settings: unknown sampler_type");
    // the offline raygen's in-shader a-trous (pt_raygen_offline.rgen:1264) is not built: refuse rather than ignore the bit
    if (s->profile == OHB_PROFILE_OFFLINE && (s->flags & OHB_FLAG_ENABLE_INTERNAL_DENOISE)) OHB_FAIL(c, "ohb_set_settings: OHB_FLAG_ENABLE_INTERNAL_DENOISE is implemented for the realtime profile only");
    if (s->profile == OHB_PROFILE_REALTIME && c->profile != OHB_PROFILE_REALTIME) OHB_FAIL(c, "ohb_set_settings: realtime settings need a context created with OHB_PROFILE_REALTIME");
    if (s->denoise_mode != OHB_DENOISE_NONE && s->denoise_mode != OHB_DENOISE_ATROUS) OHB_FAIL(c, "ohb_set_settings: denoise_mode must be OHB_DENOISE_NONE or OHB_DENOISE_ATROUS (OIDN / NRD / DLSS-RR are not part of this library)");
    if (s->denoise_mode == OHB_DENOISE_ATROUS && s->profile != OHB_PROFILE_REALTIME) OHB_FAIL(c, "ohb_set_settings: OHB_DENOISE_ATROUS (SVGF) runs in the realtime profile only");
    c->settings = *s;
    uint32_t spf = s->samples_per_frame; c->settings.samples_per_frame = spf < 1u ? 1u : (spf > 64u ? 64u : spf);   // clampSamplesPerFrame
    return 0;
}
int ohb_get_settings(ohb_ctx* c, ohb_settings* s) { if (!c || !s) return 1; *s = c->settings; return 0; }
void ohb_reset_accumulation(ohb_ctx* c) { if (!c) return; c->sampleIndex = c->seed; c->historyCount = 0; c->viewChanged = false; }
void ohb_set_seed(ohb_ctx* c, uint32_t seed) { if (!c) return; c->seed = seed; ohb_reset_accumulation(c); }
void ohb_notify_view_changed(ohb_ctx* c) {
    if (!c) return;
    if (c->settings.profile == OHB_PROFILE_OFFLINE) ohb_reset_accumulation(c);   // resetsAccumulationOnViewChange()
    else c->viewChanged = true;
}
uint32_t ohb_frame_index(const ohb_ctx* c) { return c ? c->sampleIndex : 0u; }

int ohb_set_tile(ohb_ctx* c, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h) {
    if (!c) return 1;
    if (w == 0 || h == 0 || uint64_t(x0) + w > c->W || uint64_t(y0) + h > c->H) OHB_FAIL(c, "ohb_set_tile: rectangle outside the frame");
    c->tileX = x0; c->tileY = y0; c->tileW = w; c->tileH = h;
    return 0;
}

#ifndef OHB_LANES_DEFAULT
#define OHB_LANES_DEFAULT 0     // 0 = by acceleration-structure size, 1 = one lane, 2 = two lanes
#endif
int ensurePaths(ohb_ctx* c, uint32_t cap) {
    if (cap <= c->pathCapacity) return 0;
    size_t n = cap;
    CU(c, c->rayO.reserve(n * 16)); CU(c, c->rayD.reserve(n * 16)); CU(c, c->hit.reserve(n * 16)); CU(c, c->thr.reserve(n * 16)); CU(c, c->rad.reserve(n * 16));
    CU(c, c->pendA.reserve(n * 16)); CU(c, c->pendB.reserve(n * 16)); CU(c, c->meta.reserve(n * 16));
    CU(c, c->fh0.reserve(n * 16)); CU(c, c->fh1.reserve(n * 16)); CU(c, c->fh2.reserve(n * 16));
    if (c->profile == OHB_PROFILE_REALTIME) {   // the payload goes through memory only between k_surface and k_bounce_rt; the offline k_shade keeps it in registers
        CU(c, c->pay0.reserve(n * 16)); CU(c, c->pay1.reserve(n * 16)); CU(c, c->pay2.reserve(n * 16)); CU(c, c->pay3.reserve(n * 16));
    }
    CU(c, c->shO.reserve(n * 32)); CU(c, c->shD.reserve(n * 32)); CU(c, c->queueA.reserve(n * 4)); CU(c, c->queueB.reserve(n * 4)); CU(c, c->queueS.reserve(n * 4)); CU(c, c->hitFlag.reserve(n));
    static const bool octBin = []() { const char* e = getenv("OHB_OCT_BIN"); return e && atoi(e) != 0; }();
    if (octBin) {   // octant binning, an A/B variant that is off by default (268 MB at 32 Mi paths): 4096-entry blocks, 8 bins per block; the visibility-ray queue holds up to 2 rays per path
        const size_t blocks = (2 * n + 4095) / 4096;
        CU(c, c->octPerm.reserve(2 * n * 4)); CU(c, c->octHist.reserve(blocks * 8 * 4 + 64)); CU(c, c->octScanTmp.reserve((blocks * 8 / 1024 + 4096) * 4));
    }
    c->pathCapacity = cap;
    return 0;
}

static void fillFrameCommon(ohb_ctx* c, const float* view, const float* proj, FrameParams& fr) {
    M4h iv = inverse4(view), ip = inverse4(proj);
    fr.camPos = mk3(iv.m[12], iv.m[13], iv.m[14]); fr.fwd = mk3(-iv.m[8], -iv.m[9], -iv.m[10]);
    fr.right = mk3(iv.m[0], iv.m[1], iv.m[2]); fr.up = mk3(iv.m[4], iv.m[5], iv.m[6]);
    float aspect = float(c->W) / float(c->H);
    fr.tanY = fabsf(ip.m[5]); fr.tanX = fr.tanY * aspect;
    fr.W = c->W; fr.H = c->H; fr.maxBounces = c->settings.max_bounces & 0xFFFFu; fr.flags = c->settings.flags;
    bool envOn = c->envW && c->envMapTexIdx != 0xFFFFFFFFu;
    fr.envW = envOn ? c->envW : 0u; fr.envH = envOn ? float(c->envH) : 0.0f;
    fr.fireflyClamp = c->settings.firefly_clamp_lum; fr.sss = c->settings.subsurface_strength;
    fr.aniso = c->settings.anisotropy_strength; fr.anisoRot = c->settings.anisotropy_rotation; fr.jitX = fr.jitY = 0.0f;   // no Halton jitter without NRD/DLSS-RR (Q15)
    fr.samplerType = c->settings.sampler_type;
    fr.tileX = c->tileX; fr.tileY = c->tileY; fr.tileW = c->tileW; fr.tileH = c->tileH;
}
static void fillPaths(ohb_ctx* c, PathArrays& P) {
    uint32_t* small = c->smallCounters.as<uint32_t>();
    P.rayO = c->rayO.as<f4>(); P.rayD = c->rayD.as<f4>(); P.hit = c->hit.as<ohb_hit>(); P.thr = c->thr.as<f4>(); P.rad = c->rad.as<f4>();
    P.pendA = c->pendA.as<f4>(); P.pendB = c->pendB.as<f4>(); P.meta = c->meta.as<u4>();
    P.fh0 = c->fh0.as<f4>(); P.fh1 = c->fh1.as<f4>(); P.fh2 = c->fh2.as<f4>();
    P.pay0 = c->pay0.as<f4>(); P.pay1 = c->pay1.as<f4>(); P.pay2 = c->pay2.as<f4>(); P.pay3 = c->pay3.as<f4>();
    P.shO = c->shO.as<f4>(); P.shD = c->shD.as<f4>();
    P.queueIn = c->queueA.as<uint32_t>(); P.queueOut = c->queueB.as<uint32_t>();
    P.countIn = small + 0; P.countOut = small + 1; P.shCount = small + 2;
    P.counters = c->devCounters.as<unsigned long long>();
    P.albedoAOV = c->albedoAOV.as<f4>(); P.normalAOV = c->normalAOV.as<f4>();
    P.sobolTab = c->sobolTab.as<u4>(); P.queueSorted = c->queueS.as<uint32_t>(); P.sortCount = small + 12; P.hitFlag = c->hitFlag.as<uint8_t>();
    P.octPerm = c->octPerm.as<uint32_t>(); P.octHist = c->octHist.as<uint32_t>(); P.octScanTmp = c->octScanTmp.as<uint32_t>(); P.octBlocks = uint32_t((2 * size_t(c->pathCapacity) + 4095) / 4096);
}
int ensurePaths(ohb_ctx* c, uint32_t cap);
static int ensureSvgf(ohb_ctx* c);
static void runSvgf(ohb_ctx* c, bool reset);

// One PathTracer::render() of the realtime profile per frame (path_tracer_render.cpp:686-724, :1274-1280):
// frame index = m_sampleIndex, history count, view-changed flag, prevViewProj = last frame's proj*view, ping-pong flip.
static int renderRealtime(ohb_ctx* c, const float* view, const float* proj, uint32_t nframes) {
    if (c->settings.flags & OHB_FLAG_RESTIRGI_LEGACY) OHB_FAIL(c, "ohb_render: OHB_FLAG_RESTIRGI_LEGACY (multi-bounce Stage C) is not implemented");
    if (c->tileX || c->tileY || c->tileW != c->W || c->tileH != c->H) OHB_FAIL(c, "ohb_render: the realtime profile renders full frames only (ReSTIR reuse reads neighbours)");
    uint32_t spf = c->settings.samples_per_frame; spf = spf < 1u ? 1u : (spf > 64u ? 64u : spf);
    uint32_t tilesX = (c->W + 7u) / 8u, tilesY = (c->H + 3u) / 4u, numPixels = tilesX * tilesY * 32u;
    if (uint64_t(numPixels) * spf > (64ull << 20)) OHB_FAIL(c, "ohb_render: width*height*samples_per_frame too large");
    if (ensurePaths(c, numPixels * spf)) return 1;
    CU(c, c->sobolTab.reserve(size_t(spf) * 16));
    size_t npx = size_t(c->W) * c->H;
    for (int k = 0; k < 3; k++) if (c->rtDumpHost[k]) CU(c, c->rtDump[k].reserve(npx * 16));
    SceneDev s; fillScene(c, s);
    uint32_t* small = c->smallCounters.as<uint32_t>();
    const bool svgf = c->settings.denoise_mode == OHB_DENOISE_ATROUS;
    if (svgf && ensureSvgf(c)) return 1;
    float currVP[16];   // proj * view (column-major)
    for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) {
        float acc = 0.0f; for (int k = 0; k < 4; k++) acc += proj[k * 4 + row] * view[col * 4 + k];
        currVP[col * 4 + row] = acc;
    }
    for (uint32_t f = 0; f < nframes; f++) {
        FrameParams fr{}; fillFrameCommon(c, view, proj, fr);
        memcpy(fr.prevViewProj, c->prevViewProj, 64);
        memcpy(fr.currViewProj, currVP, 64);
        fr.viewRow2[0] = view[2]; fr.viewRow2[1] = view[6]; fr.viewRow2[2] = view[10]; fr.viewRow2[3] = view[14];
        // denoiseWantsFreshSample (path_tracer_render.cpp:707-712): with the SVGF denoiser the raygen sees historyFrameCount = 0
        // every frame (a fresh N-spp beauty, no EMA, no ReSTIR reuse); the denoiser owns the temporal accumulation
        fr.frameIdx = c->sampleIndex; fr.historyCount = svgf ? 0u : c->historyCount; fr.viewChanged = c->viewChanged ? 1u : 0u; fr.spf = spf;
        fr.jitterSobol = sobolQuad(fr.frameIdx);
        PathArrays P{}; fillPaths(c, P);
        P.numPixels = numPixels; P.samplesInBatch = spf; P.firstSampleIndex = fr.frameIdx * spf;
        int prev = c->rtCur, cur = 1 - c->rtCur;
        std::swap(c->accum, c->accumPrev);            // c->accum = image written by this frame
        RTImagesDev im{};
        im.accumPrev = c->accumPrev.as<f4>(); im.accumCurr = c->accum.as<f4>();
        im.surfPrev = c->surf[prev].as<f4>(); im.surfCurr = c->surf[cur].as<f4>(); im.shadPrev = c->shad[prev].as<f4>(); im.shadCurr = c->shad[cur].as<f4>();
        im.res0Prev = c->res[prev][0].as<f4>(); im.res1Prev = c->res[prev][1].as<f4>(); im.res2Prev = c->res[prev][2].as<f4>();
        im.res0Curr = c->res[cur][0].as<f4>(); im.res1Curr = c->res[cur][1].as<f4>(); im.res2Curr = c->res[cur][2].as<f4>();
        im.radianceDump = c->rtDumpHost[0] ? c->rtDump[0].as<float>() : nullptr; im.giDump = c->rtDumpHost[1] ? c->rtDump[1].as<float>() : nullptr;
        im.counters = c->devCounters.as<unsigned long long>();
        im.motionAOV = svgf ? c->motionAOV.as<uint32_t>() : nullptr; im.depthAOV = svgf ? c->depthAOV.as<float>() : nullptr;
        launchRealtimeFrame(s, fr, P, im, c->ldr.as<uint32_t>(), c->rtDumpHost[2] ? c->rtDump[2].as<float>() : nullptr, small + 4, c->numSMs, c->stream,
                            &c->launches, c->timingOn ? &c->timing : nullptr);
        c->rtCur = cur;
        if (svgf) runSvgf(c, c->historyCount == 0u);   // AtrousDenoiser::dispatch on the finished beauty; history resets on the first frame only (path_tracer_render.cpp:756-762)
        // m_prevViewProj = proj * view (column-major)
        for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) {
            float acc = 0.0f; for (int k = 0; k < 4; k++) acc += proj[k * 4 + row] * view[col * 4 + k];
            c->prevViewProj[col * 4 + row] = acc;
        }
        c->sampleIndex += 1; c->historyCount += 1; c->viewChanged = false;
    }
    CU(c, cudaGetLastError());
    bool any = false;
    for (int k = 0; k < 3; k++) if (c->rtDumpHost[k]) { CU(c, cudaMemcpyAsync(c->rtDumpHost[k], c->rtDump[k].p, npx * 16, cudaMemcpyDeviceToHost, c->stream)); any = true; }
    if (any) CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int ohb_render(ohb_ctx* c, const float view[16], const float proj[16], uint32_t nsamples) {
    if (!c) return 1;
    if (!view || !proj) OHB_FAIL(c, "ohb_render: null matrix");
    if (!c->accelValid) { c->err = "ohb_render: no acceleration structure (call ohb_build_accel)"; return 1; }   // reference: silent return (path_tracer_render.cpp:42)
    if (!c->nmat) OHB_FAIL(c, "ohb_render: no materials");
    if (c->maxMatId >= c->nmat) OHB_FAIL(c, "ohb_render: a triangle's material id is out of range of the material buffer");
    if (nsamples == 0) return 0;
    cudaSetDevice(c->device);
    if (c->settings.profile == OHB_PROFILE_REALTIME) return renderRealtime(c, view, proj, nsamples);
    FrameParams fr{}; fillFrameCommon(c, view, proj, fr);
    uint32_t tilesX = (c->tileW + 7u) / 8u, tilesY = (c->tileH + 3u) / 4u;
    if (uint64_t(tilesX) * tilesY * 32u > (1ull << 28)) OHB_FAIL(c, "ohb_render: more than 2^28 pixels in one tile (queue entries hold a 28-bit path index): render in tiles with ohb_set_tile");
    uint32_t numPixels = tilesX * tilesY * 32u;
    static const uint32_t maxPaths = []() { const char* e = getenv("OHB_MAX_PATHS"); uint32_t v = e ? uint32_t(strtoul(e, nullptr, 10)) : 0u; return v ? std::min(v, 1u << 28) : (32u << 20); }();   // queue entries hold a 28-bit path index (OHB_Q_PATH)
    uint32_t spb = std::max(1u, std::min(nsamples, maxPaths / std::max(numPixels, 1u)));
    if (ensurePaths(c, numPixels * spb)) return 1;
    CU(c, c->sobolTab.reserve(size_t(spb) * 16));
    SceneDev s; fillScene(c, s);
    float* dumpDev = nullptr;
    if (c->dumpHost) {
        size_t need = size_t(nsamples) * c->W * c->H * 4u;
        if (need > c->dumpCapacity) OHB_FAIL(c, "ohb_render: sample dump buffer too small");
        CU(c, c->sampleDump.reserve(need * 4)); CU(c, cudaMemsetAsync(c->sampleDump.p, 0, need * 4, c->stream));
        dumpDev = c->sampleDump.as<float>();
    }
    uint32_t* small = c->smallCounters.as<uint32_t>();
    uint32_t done = 0;
    // Two lanes (OHB_LANES): a batch's samples are split in two halves that run the wavefront on two streams, each in
    // its own slice of the path arrays, so the drain of every launch of one lane (persistent CTAs leave as the queue empties) is
    // filled by the other lane's kernels.  k_film folds sample by sample in fp32 through the stored accumulation image, and the two
    // film launches are ordered lane 0 -> lane 1 by an event: the image is bit-identical to the one-lane run.  Per-kernel timing
    // (ohb_enable_timing) needs serial launches and runs one lane.
    // Measured (profiles/r2ag_sweep_lanes.txt): +1.5 % on the 50 K-triangle scene, +2.2 % on Cornell, -1.1 % on the 2 M-triangle scene,
    // whose 130 MB of nodes + triangles already overflow the L2 with ONE kernel's working set — so the default (OHB_LANES unset or 0)
    // runs two lanes only while the acceleration structure is at most a quarter of the 126 MB L2.
    static const int lanesEnv = []() { const char* e = getenv("OHB_LANES"); return e ? atoi(e) : OHB_LANES_DEFAULT; }();
    const size_t accelBytes = size_t(c->numActive) * 48u + size_t(c->stats.num_nodes) * 128u;
    const bool twoLanes = (lanesEnv >= 2 || (lanesEnv == 0 && accelBytes <= (32u << 20))) && !c->timingOn;
    if (twoLanes && !c->stream2) {
        CU(c, cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        CU(c, cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming)); CU(c, cudaEventCreateWithFlags(&c->evFilm, cudaEventDisableTiming)); CU(c, cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    }
    while (done < nsamples) {
        uint32_t k = std::min(spb, nsamples - done);
        const uint32_t kA = (twoLanes && k >= 2u) ? (k + 1u) / 2u : k, kB = k - kA;
        PathArrays P{}; fillPaths(c, P);
        P.numPixels = numPixels; P.samplesInBatch = kA; P.firstSampleIndex = c->sampleIndex;
        FilmArrays F{};
        F.accum = c->accum.as<f4>(); F.ldr = c->ldr.as<uint32_t>(); F.historyCount = c->historyCount; F.sumMode = c->sumMode;
        F.sampleDump = dumpDev ? dumpDev + size_t(done) * c->W * c->H * 4u : nullptr;
        if (!kB) launchOfflineBatch(s, fr, P, F, small + 4, c->numSMs, c->stream, &c->launches, c->timingOn ? &c->timing : nullptr, 3);
        else {
            PathArrays Q = P; FilmArrays G = F;
            const size_t base = size_t(numPixels) * kA;      // path index = sample * numPixels + pixel: lane 1 owns the upper samples
            Q.rayO += base; Q.rayD += base; Q.hit += base; Q.thr += base; Q.rad += base; Q.pendA += base; Q.pendB += base; Q.meta += base;
            Q.fh0 += base; Q.fh1 += base; Q.fh2 += base; Q.shO += 2 * base; Q.shD += 2 * base;
            Q.queueIn += base; Q.queueOut += base; Q.queueSorted += base; Q.hitFlag += base; Q.octPerm = nullptr;
            Q.countIn += 32; Q.countOut += 32; Q.shCount += 32; Q.sortCount += 32; Q.counters += 8; Q.sobolTab += kA;
            Q.samplesInBatch = kB; Q.firstSampleIndex = c->sampleIndex + kA;
            P.albedoAOV = nullptr;                           // the AOVs are the first hit of the batch's LAST sample: lane 1 writes them
            G.historyCount = F.historyCount + kA; G.sampleDump = F.sampleDump ? F.sampleDump + size_t(kA) * c->W * c->H * 4u : nullptr;
            CU(c, cudaEventRecord(c->evFork, c->stream)); CU(c, cudaStreamWaitEvent(c->stream2, c->evFork, 0));
            launchOfflineBatch(s, fr, P, F, small + 4, c->numSMs, c->stream, &c->launches, nullptr, 1);
            launchOfflineBatch(s, fr, Q, G, small + 32 + 4, c->numSMs, c->stream2, &c->launches, nullptr, 1);
            launchOfflineBatch(s, fr, P, F, small + 4, c->numSMs, c->stream, &c->launches, nullptr, 2);
            CU(c, cudaEventRecord(c->evFilm, c->stream)); CU(c, cudaStreamWaitEvent(c->stream2, c->evFilm, 0));
            launchOfflineBatch(s, fr, Q, G, small + 32 + 4, c->numSMs, c->stream2, &c->launches, nullptr, 2);
            CU(c, cudaEventRecord(c->evJoin, c->stream2)); CU(c, cudaStreamWaitEvent(c->stream, c->evJoin, 0));
        }
        c->sampleIndex += k; c->historyCount += k; done += k;   // path_tracer_render.cpp:1274-1275
    }
    c->viewChanged = false;
    CU(c, cudaGetLastError());
    if (c->dumpHost) {
        CU(c, cudaMemcpyAsync(c->dumpHost, c->sampleDump.p, size_t(nsamples) * c->W * c->H * 16u, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int ohb_synchronize(ohb_ctx* c) {
    if (!c) return 1;
    cudaSetDevice(c->device);
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaGetLastError());
    return 0;
}
int ohb_read_ldr(ohb_ctx* c, uint8_t* rgba8) {
    if (!c || !rgba8) return 1;
    cudaSetDevice(c->device);
    CU(c, cudaMemcpyAsync(rgba8, c->ldr.p, size_t(c->W) * c->H * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int ohb_read_hdr(ohb_ctx* c, float* accum, float* albedo, float* normal) {
    if (!c) return 1;
    cudaSetDevice(c->device);
    size_t bytes = size_t(c->W) * c->H * 16;
    if (accum) CU(c, cudaMemcpyAsync(accum, c->accum.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (albedo) CU(c, cudaMemcpyAsync(albedo, c->albedoAOV.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (normal) CU(c, cudaMemcpyAsync(normal, c->normalAOV.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}
void* ohb_accum_dev_ptr(ohb_ctx* c, size_t* bytes) { if (!c) return nullptr; if (bytes) *bytes = size_t(c->W) * c->H * 16; return c->accum.p; }
int ohb_set_accum_mode(ohb_ctx* c, int sum_mode) { if (!c) return 1; c->sumMode = sum_mode ? 1 : 0; return 0; }
int ohb_clear_accum(ohb_ctx* c) {
    if (!c) return 1;
    cudaSetDevice(c->device);
    CU(c, cudaMemsetAsync(c->accum.p, 0, size_t(c->W) * c->H * 16, c->stream));
    return 0;
}
int ohb_resolve(ohb_ctx* c) {
    if (!c) return 1;
    cudaSetDevice(c->device);
    launchResolve(c->accum.as<f4>(), c->ldr.as<uint32_t>(), c->W * c->H, c->sumMode, c->stream, &c->launches);
    CU(c, cudaGetLastError());
    return 0;
}

static int traceHook(ohb_ctx* c, const ohb_ray* rays, uint32_t n, ohb_hit* hits, uint8_t* occ) {
    if (!c) return 1;
    if (!c->accelValid) OHB_FAIL(c, "trace: no acceleration structure");
    if (n == 0) return 0;
    cudaSetDevice(c->device);
    CU(c, c->hookRays.reserve(size_t(n) * sizeof(ohb_ray))); CU(c, c->hookOut.reserve(size_t(n) * 16));
    CU(c, cudaMemcpyAsync(c->hookRays.p, rays, size_t(n) * sizeof(ohb_ray), cudaMemcpyHostToDevice, c->stream));
    SceneDev s; fillScene(c, s);
    launchTraceBatch(s, c->hookRays.as<ohb_ray>(), n, hits ? c->hookOut.as<ohb_hit>() : nullptr, c->hookOut.as<uint8_t>(), c->smallCounters.as<uint32_t>() + 8, c->numSMs, c->stream, &c->launches);
    if (hits) CU(c, cudaMemcpyAsync(hits, c->hookOut.p, size_t(n) * 16, cudaMemcpyDeviceToHost, c->stream));
    else      CU(c, cudaMemcpyAsync(occ, c->hookOut.p, n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaGetLastError());
    return 0;
}
int ohb_trace_batch(ohb_ctx* c, const ohb_ray* rays, uint32_t n, ohb_hit* hits) { if (!rays || !hits) return 1; return traceHook(c, rays, n, hits, nullptr); }
int ohb_occluded_batch(ohb_ctx* c, const ohb_ray* rays, uint32_t n, uint8_t* occ) { if (!rays || !occ) return 1; return traceHook(c, rays, n, nullptr, occ); }

static int ensureSvgf(ohb_ctx* c) {
    size_t npx = size_t(c->W) * c->H;
    CU(c, c->motionAOV.reserve(npx * 4)); CU(c, c->depthAOV.reserve(npx * 4));
    for (int i = 0; i < 2; i++) {
        CU(c, c->svgfHistColor[i].reserve(npx * 8)); CU(c, c->svgfHistMoments[i].reserve(npx * 8)); CU(c, c->svgfHistGeom[i].reserve(npx * 8));
        CU(c, c->svgfColor[i].reserve(npx * 8)); CU(c, c->svgfVar[i].reserve(npx * 2));
    }
    return 0;
}
static void runSvgf(ohb_ctx* c, bool reset) {
    SvgfBuffers b{};
    b.beauty = c->ldr.as<uint32_t>(); b.motion = c->motionAOV.as<uint32_t>(); b.depth = c->depthAOV.as<float>(); b.normal = c->normalAOV.as<f4>();
    for (int i = 0; i < 2; i++) {
        b.histColor[i] = c->svgfHistColor[i].as<h4>(); b.histMoments[i] = c->svgfHistMoments[i].as<h4>(); b.histGeom[i] = c->svgfHistGeom[i].as<h4>();
        b.color[i] = c->svgfColor[i].as<h4>(); b.var[i] = c->svgfVar[i].as<uint16_t>();
    }
    b.sigmaL = OHB_SVGF_SIGMA_L; b.sigmaNormal = OHB_SVGF_SIGMA_NORMAL; b.sigmaDepth = OHB_SVGF_SIGMA_DEPTH;
    const int scur = 1 - c->svgfCur;
    launchSvgf(b, c->W, c->H, scur, reset, c->stream, &c->launches, c->timingOn ? &c->timing : nullptr);
    c->svgfCur = scur;
}
int ohb_svgf_dispatch(ohb_ctx* c, uint8_t* beauty, const float* normal, const float* depth, const uint32_t* motion, int reset) {
    if (!c) return 1;
    if (!beauty || !normal || !depth || !motion) OHB_FAIL(c, "ohb_svgf_dispatch: null image");
    if (c->profile != OHB_PROFILE_REALTIME) OHB_FAIL(c, "ohb_svgf_dispatch: needs a context created with OHB_PROFILE_REALTIME");
    cudaSetDevice(c->device);
    if (ensureSvgf(c)) return 1;
    size_t npx = size_t(c->W) * c->H;
    CU(c, cudaMemcpyAsync(c->ldr.p, beauty, npx * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->normalAOV.p, normal, npx * 16, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->depthAOV.p, depth, npx * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->motionAOV.p, motion, npx * 4, cudaMemcpyHostToDevice, c->stream));
    runSvgf(c, reset != 0);
    CU(c, cudaMemcpyAsync(beauty, c->ldr.p, npx * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaGetLastError());
    return 0;
}
int ohb_read_denoise_state(ohb_ctx* c, uint16_t* color, uint16_t* moments, uint16_t* geom, uint32_t* motion, float* depth) {
    if (!c) return 1;
    if (!c->motionAOV.p) OHB_FAIL(c, "ohb_read_denoise_state: no frame has been rendered with OHB_DENOISE_ATROUS");
    cudaSetDevice(c->device);
    size_t n = size_t(c->W) * c->H; int k = c->svgfCur;
    if (color) CU(c, cudaMemcpyAsync(color, c->svgfHistColor[k].p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (moments) CU(c, cudaMemcpyAsync(moments, c->svgfHistMoments[k].p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (geom) CU(c, cudaMemcpyAsync(geom, c->svgfHistGeom[k].p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (motion) CU(c, cudaMemcpyAsync(motion, c->motionAOV.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (depth) CU(c, cudaMemcpyAsync(depth, c->depthAOV.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int ohb_set_sample_dump(ohb_ctx* c, float* host, size_t cap) { if (!c) return 1; c->dumpHost = host; c->dumpCapacity = host ? cap : 0; return 0; }

int ohb_get_counters(ohb_ctx* c, ohb_counters* out) {
    if (!c || !out) return 1;
    cudaSetDevice(c->device);
    unsigned long long d[16];
    CU(c, cudaMemcpyAsync(d, c->devCounters.p, 128, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 8; i++) d[i] += d[8 + i];          // the second lane of ohb_render counts in its own block
    memset(out, 0, sizeof(*out));
    out->samples = d[0]; out->closest_rays = d[1]; out->shadow_rays = d[2]; out->closest_hits = d[3]; out->kernel_launches = c->launches;
    return 0;
}
void ohb_reset_counters(ohb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaMemsetAsync(c->devCounters.p, 0, 128, c->stream);
    c->launches = 0; c->timing.reset();
}
int ohb_enable_timing(ohb_ctx* c, int enable) {
    if (!c) return 1;
    if (!enable && c->timingOn) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); c->timing.collect(); }
    c->timingOn = enable != 0; return 0;
}
int ohb_get_timing(ohb_ctx* c, float* trace_ms, float* shade_ms, float* total_ms) {
    if (!c) return 1;
    cudaSetDevice(c->device);
    CU(c, cudaStreamSynchronize(c->stream));
    c->timing.collect();
    if (trace_ms) *trace_ms = float(c->timing.ms[0] + c->timing.ms[2]);
    if (shade_ms) *shade_ms = float(c->timing.ms[1] + c->timing.ms[4]);
    if (total_ms) { double t = 0; for (int i = 0; i < 8; i++) t += c->timing.ms[i]; *total_ms = float(t); }
    return 0;
}

int ohb_get_timing_detail(ohb_ctx* c, float ms[8], uint64_t launches[8]) {
    if (!c || !ms || !launches) return 1;
    cudaSetDevice(c->device);
    CU(c, cudaStreamSynchronize(c->stream));
    c->timing.collect();
    for (int i = 0; i < 8; i++) { ms[i] = float(c->timing.ms[i]); launches[i] = c->timing.count[i]; }
    return 0;
}

int ohb_set_realtime_dump(ohb_ctx* c, float* radiance, float* gi, float* denoised) {
    if (!c) return 1;
    c->rtDumpHost[0] = radiance; c->rtDumpHost[1] = gi; c->rtDumpHost[2] = denoised;
    return 0;
}
int ohb_read_realtime_state(ohb_ctx* c, float* res0, float* res1, float* res2, float* surf, float* shad) {
    if (!c) return 1;
    if (c->profile != OHB_PROFILE_REALTIME) OHB_FAIL(c, "ohb_read_realtime_state: context was not created with OHB_PROFILE_REALTIME");
    cudaSetDevice(c->device);
    size_t bytes = size_t(c->W) * c->H * 16; int k = c->rtCur;
    if (res0) CU(c, cudaMemcpyAsync(res0, c->res[k][0].p, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (res1) CU(c, cudaMemcpyAsync(res1, c->res[k][1].p, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (res2) CU(c, cudaMemcpyAsync(res2, c->res[k][2].p, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (surf) CU(c, cudaMemcpyAsync(surf, c->surf[k].p, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (shad) CU(c, cudaMemcpyAsync(shad, c->shad[k].p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int ohb_timer_start(ohb_ctx* c) {
    if (!c) return 1;
    cudaSetDevice(c->device);
    CU(c, cudaEventRecord(c->evT0, c->stream));
    return 0;
}
int ohb_timer_stop(ohb_ctx* c, float* ms) {
    if (!c || !ms) return 1;
    cudaSetDevice(c->device);
    CU(c, cudaEventRecord(c->evT1, c->stream));
    CU(c, cudaEventSynchronize(c->evT1));
    CU(c, cudaEventElapsedTime(ms, c->evT0, c->evT1));
    return 0;
}

}  // extern "C"

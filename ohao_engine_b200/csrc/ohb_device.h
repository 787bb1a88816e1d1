// ohb_device.h — launch wrappers exported by ohb_kernels.cu to the C-ABI layer (ohb_api.cu).
#pragma once
#include "ohb_bvh.h"
#include "ohb_integrator.h"
#include "ohb_realtime.h"
#include "ohb_hybrid.h"
#include <cuda_runtime.h>
#include <vector>

namespace ohb {

// CUDA-event timing of kernel categories on the launching stream (bench.py roofline line):
// 0 = closest-hit traversal, 1 = bounce (raygen body), 2 = any-hit traversal, 3 = film, 4 = surface (hit/miss shaders),
// 5 = realtime per-pixel pass, 6 = SVGF denoiser, 7 = hit/miss queue sort (k_sort_hits).
struct TimingHooks {
    struct Span { cudaEvent_t a, b; int cat; };
    std::vector<Span> spans; size_t used = 0; 
    double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0}; uint64_t count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    static constexpr size_t kMaxSpans = 8192;     // a caller that never reads the timing must not grow the event pool without bound
    void begin(int cat, cudaStream_t st) {
        if (used == kMaxSpans) { cudaStreamSynchronize(st); collect(); }
        if (used == spans.size()) { Span s; cudaEventCreate(&s.a); cudaEventCreate(&s.b); s.cat = cat; spans.push_back(s); }
        spans[used].cat = cat; cudaEventRecord(spans[used].a, st);
    }
    void end(int, cudaStream_t st) { cudaEventRecord(spans[used].b, st); used++; }
    void collect() {   // call after the stream is synchronised
        for (size_t i = 0; i < used; i++) { float t = 0; cudaEventElapsedTime(&t, spans[i].a, spans[i].b); ms[spans[i].cat] += t; count[spans[i].cat]++; }
        used = 0;
    }
    void reset() { collect(); for (int i = 0; i < 8; i++) { ms[i] = 0; count[i] = 0; } }
    ~TimingHooks() { for (auto& s : spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); } }
};

cudaError_t uploadConstants();
uint32_t radixSortTempWords(uint32_t n);
void launchBuild(const BuildArrays& b, uint64_t* keysTmp, uint32_t* valsTmp, uint32_t* sortTemp, WideItem* itemsA, WideItem* itemsB,
                 uint32_t treeletPasses, cudaStream_t st, uint64_t* launches);
void launchRefit(const BuildArrays& b, WideItem* itemsA, WideItem* itemsB, cudaStream_t st, uint64_t* launches);
void launchBuildBlas(const BuildArrays& b, uint64_t* keysTmp, uint32_t* valsTmp, uint32_t* sortTemp, WideItem* itemsA, WideItem* itemsB, uint32_t treeletPasses,
                     f4* blasLo, f4* blasHi, uint32_t inst, uint32_t* maxLevels, cudaStream_t st, uint64_t* launches);
void launchBuildTlas(const BuildArrays& b, const f4* blasLo, const f4* blasHi, const uint32_t* instOfPrim, bool refit, uint64_t* keysTmp, uint32_t* valsTmp, uint32_t* sortTemp,
                     WideItem* itemsA, WideItem* itemsB, cudaStream_t st, uint64_t* launches);
void launchEnvCdf(const f4* env, uint32_t W, uint32_t H, float* cond, float* marg, float* rowTotal, float* integral, float* condTop, float* margTop, cudaStream_t st, uint64_t* launches);
void launchEnvSample(const SceneDev& sc, const float* u12, uint32_t n, f4* dirPdf, float* pdfOfDir, cudaStream_t st, uint64_t* launches);
void launchHybridShadow(const SceneDev& sc, const HybridShadowParams& pc, const f4* gPos, const f2* gNrm, uint8_t* mask, cudaStream_t st, uint64_t* launches);
void launchHybridGi(const SceneDev& sc, const HybridGiParams& pc, const f4* gPos, const f2* gNrm, const f4* gAlbedo, const f4* history, const f4* instMat, h4* out, cudaStream_t st, uint64_t* launches);
void launchNrdPack(const float* in6, const float* nr4, uint32_t n, f4* packedRad, f4* packedNormal, float* unpackedRgb, cudaStream_t st, uint64_t* launches);
void launchEnvPdf(const SceneDev& sc, const float* dirs3, uint32_t n, float* pdf, cudaStream_t st, uint64_t* launches);
void launchOfflineBatch(const SceneDev& sc, const FrameParams& fr, PathArrays P, const FilmArrays& F,
                        uint32_t* work, int numSMs, cudaStream_t st, uint64_t* launches, TimingHooks* th, int phases = 3);
void launchRealtimeFrame(const SceneDev& sc, const FrameParams& fr, PathArrays P, const RTImagesDev& im, uint32_t* ldr, float* denoisedDump,
                         uint32_t* work, int numSMs, cudaStream_t st, uint64_t* launches, TimingHooks* th);
// images of the SVGF denoiser: persistent ping-ponged history (colour, moments, geometry) + per-frame scratch
struct SvgfBuffers {
    uint32_t* beauty; const uint32_t* motion; const float* depth; const f4* normal;
    h4* histColor[2]; h4* histMoments[2]; h4* histGeom[2]; h4* color[2]; uint16_t* var[2];
    float sigmaL, sigmaNormal, sigmaDepth;
};
void launchSvgf(const SvgfBuffers& b, uint32_t W, uint32_t H, int cur, bool reset, cudaStream_t st, uint64_t* launches, TimingHooks* th);
void launchResolve(f4* accum, uint32_t* ldr, uint32_t n, int sumMode, cudaStream_t st, uint64_t* launches);
void launchTraceBatch(const SceneDev& sc, const ohb_ray* rays, uint32_t n, ohb_hit* hits, uint8_t* occ, uint32_t* work, int numSMs, cudaStream_t st, uint64_t* launches);

}  // namespace ohb

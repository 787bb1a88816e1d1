// ohb_kernels.cu — sm_100a kernels of the B200 path tracer and their launch wrappers.
//
// Every kernel here is a thin scheduling shell around the per-thread code of ohb_bvh.h /
// ohb_traverse.h / ohb_integrator.h.  Scheduling choices (DESIGN.md "Kernels"):
//   * traversal kernels are persistent: gridDim = SMs x resident blocks, each warp pulls 32 rays at
//     a time from a global work counter, so the tail of an incoherent wavefront does not idle SMs;
//   * path queues are emitted stable at tile granularity (TileEmit: 1024 entries staged in shared memory, one atomic
//     per tile; shadow-ray pushes are warp-aggregated), no host round trip between the iterations of a wavefront:
//     queue sizes live in device memory and kernels read them;
//   * the radix sort ranks keys with __match_any_sync per warp-private digit counters (stable,
//     no shared-memory atomics in the scatter loop).
#include "ohb_device.h"
#include <cuda_runtime.h>
#include <cstdlib>
#include <type_traits>

namespace ohb {

cudaError_t uploadConstants() { return cudaSuccess; }   // constants are statically initialised (ohb_scene.h)

static inline unsigned gridFor(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

// =============================================================================================
// BVH build
// =============================================================================================
__global__ void k_world_tris(BuildArrays b) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < b.n) buildWorldTri(b, i); }
__global__ void k_morton(BuildArrays b) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < b.n) buildMorton(b, i); }
__global__ void k_hierarchy(BuildArrays b) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i + 1 < b.n) buildHierarchyNode(b, int(i)); }
__global__ void k_sweep(BuildArrays b, uint32_t gamma) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < b.n) sweepFromLeaf(b, i, gamma); }
__global__ void k_clear_visit(BuildArrays b) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i + 1 < b.n) b.visit[i] = 0u; }
// One level of the 8-wide collapse: thread i expands item i of `in` (count read from the device) and appends the
// items of its inner children to `out`.  The grid is sized for the worst case of the level; surplus threads exit.
__global__ void __launch_bounds__(64) k_wide_level(BuildArrays b, const WideItem* in, const uint32_t* inCount, WideItem* out, uint32_t* outCount) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *inCount) emitWideNode(b, in[i], out, outCount);
}
__global__ void k_wide_begin(BuildArrays b, WideItem* items) {
    if (blockIdx.x || threadIdx.x) return;
    WideItem r; r.bvh2 = 0; r.wide = 0u; r.triStart = 0u; r.pad = 0u; items[0] = r;
    b.wideCounters[0] = 1u; b.wideCounters[1] = 1u; b.wideCounters[2] = 0u; b.wideCounters[3] = 0u;
}
__global__ void k_wide_next(uint32_t* consumed, uint32_t* levels, const uint32_t* produced) {
    if (blockIdx.x || threadIdx.x) return;
    *consumed = 0u; if (*produced) (*levels)++;
}
__global__ void k_init_build(BuildArrays b) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < b.n) b.visit[i] = 0u;
    if (i == 0) {
        b.boundsBits[0] = b.boundsBits[1] = b.boundsBits[2] = 0xFFFFFFFFu;
        b.boundsBits[3] = b.boundsBits[4] = b.boundsBits[5] = 0u;
        b.sah[0] = 0.0f; b.sah[1] = 0.0f;
    }
}

// ---- exclusive scan of u32 (block = 256 threads x 8 items) -----------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
__global__ void k_scan_block(uint32_t* data, uint32_t n, uint32_t* blockSums) {
    __shared__ uint32_t warpSums[SCAN_THREADS / 32];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS]; uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? data[base + k] : 0u; sum += v[k]; }
    uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += t; }
    if (lane == 31u) warpSums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? warpSums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= (uint32_t)o) w += t; }
        if (lane < SCAN_THREADS / 32) warpSums[lane] = w;
    }
    __syncthreads();
    uint32_t excl = inc - sum + (warp ? warpSums[warp - 1] : 0u);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) data[base + k] = excl; excl += v[k]; }
    if (threadIdx.x == SCAN_THREADS - 1 && blockSums) blockSums[blockIdx.x] = excl;
}
__global__ void k_scan_add(uint32_t* data, uint32_t n, const uint32_t* blockSums) {
    uint32_t i = blockIdx.x * SCAN_TILE + threadIdx.x;
    uint32_t add = blockSums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { uint32_t j = i + k * SCAN_THREADS; if (j < n) data[j] += add; }
}
// temp must hold ceil(n/2048) + ceil(that/2048) + ... + 8 entries
void exclusiveScanU32(uint32_t* data, uint32_t n, uint32_t* temp, cudaStream_t st, uint64_t* launches) {
    uint32_t nb = gridFor(n, SCAN_TILE);
    k_scan_block<<<nb, SCAN_THREADS, 0, st>>>(data, n, nb > 1 ? temp : nullptr); (*launches)++;
    if (nb > 1) {
        exclusiveScanU32(temp, nb, temp + nb, st, launches);
        k_scan_add<<<nb, SCAN_THREADS, 0, st>>>(data, n, temp); (*launches)++;
    }
}

// ---- LSD radix sort, 8 bits per pass, key u64 / value u32 --------------------------------------
#define RS_THREADS 256
#define RS_WARPS 8
#define RS_ROUNDS 16
#define RS_TILE (RS_THREADS * RS_ROUNDS)
__global__ void k_radix_hist(const uint64_t* keys, uint32_t n, int shift, uint32_t* blockHist, uint32_t numBlocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t base = blockIdx.x * RS_TILE;
    for (uint32_t i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        uint32_t idx = base + i;
        if (idx < n) atomicAdd(&h[uint32_t(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    blockHist[threadIdx.x * numBlocks + blockIdx.x] = h[threadIdx.x];
}
__global__ void k_radix_scatter(const uint64_t* keysIn, const uint32_t* valsIn, uint64_t* keysOut, uint32_t* valsOut,
                                uint32_t n, int shift, const uint32_t* scannedHist, uint32_t numBlocks) {
    __shared__ uint32_t wh[RS_WARPS][256];
    uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (uint32_t i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0u;
    __syncthreads();
    uint32_t chunk = blockIdx.x * RS_TILE + warp * (32u * RS_ROUNDS);
    for (int r = 0; r < RS_ROUNDS; r++) {
        uint32_t idx = chunk + r * 32u + lane;
        bool valid = idx < n;
        uint32_t digit = valid ? (uint32_t(keysIn[idx] >> shift) & 255u) : 256u;
        unsigned peers = __match_any_sync(0xffffffffu, digit);
        if (valid && lane == uint32_t(__ffs(int(peers)) - 1)) wh[warp][digit] += uint32_t(__popc(peers));
        __syncwarp();
    }
    __syncthreads();
    {
        uint32_t d = threadIdx.x;
        uint32_t run = scannedHist[d * numBlocks + blockIdx.x];
        for (int w = 0; w < RS_WARPS; w++) { uint32_t t = wh[w][d]; wh[w][d] = run; run += t; }
    }
    __syncthreads();
    for (int r = 0; r < RS_ROUNDS; r++) {
        uint32_t idx = chunk + r * 32u + lane;
        bool valid = idx < n;
        uint64_t key = valid ? keysIn[idx] : 0ull;
        uint32_t digit = valid ? (uint32_t(key >> shift) & 255u) : 256u;
        unsigned peers = __match_any_sync(0xffffffffu, digit);
        uint32_t rank = uint32_t(__popc(peers & ((1u << lane) - 1u)));
        if (valid) {
            uint32_t dst = wh[warp][digit] + rank;
            keysOut[dst] = key; valsOut[dst] = valsIn[idx];
        }
        __syncwarp();
        if (valid && lane == uint32_t(__ffs(int(peers)) - 1)) wh[warp][digit] += uint32_t(__popc(peers));
        __syncwarp();
    }
}
uint32_t radixSortTempWords(uint32_t n) { uint32_t nb = gridFor(n, RS_TILE); uint32_t h = 256u * nb; return h + h / 1024u + 4096u; }
// Sorts (keys, vals) ascending; result ends in keys/vals (8 passes = even number of ping-pongs).
void radixSort64(uint64_t* keys, uint32_t* vals, uint64_t* keysTmp, uint32_t* valsTmp, uint32_t n, uint32_t* temp, cudaStream_t st, uint64_t* launches) {
    uint32_t nb = gridFor(n, RS_TILE);
    uint32_t* hist = temp; uint32_t* scanTmp = temp + 256u * nb;
    uint64_t* kin = keys; uint32_t* vin = vals; uint64_t* kout = keysTmp; uint32_t* vout = valsTmp;
    for (int pass = 0; pass < 8; pass++) {
        int shift = pass * 8;
        k_radix_hist<<<nb, RS_THREADS, 0, st>>>(kin, n, shift, hist, nb); (*launches)++;
        exclusiveScanU32(hist, 256u * nb, scanTmp, st, launches);
        k_radix_scatter<<<nb, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, hist, nb); (*launches)++;
        uint64_t* tk = kin; kin = kout; kout = tk; uint32_t* tv = vin; vin = vout; vout = tv;
    }
}

__global__ void k_tlas_prims(BuildArrays b, const f4* blasLo, const f4* blasHi, const uint32_t* instOfPrim) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < b.n) buildTlasPrim(b, blasLo, blasHi, instOfPrim, i);
}
__global__ void k_store_root_box(BuildArrays b, f4* blasLo, f4* blasHi, uint32_t inst, uint32_t* maxLevels) {
    if (blockIdx.x || threadIdx.x) return;
    storeBlasRootBox(b, blasLo, blasHi, inst, maxLevels);
}
// the 8-wide collapse, level by level.  Level l holds at most min(8^l, n / 4) nodes (every inner child owns > 3
// triangles); the loop runs OHB_MAX_LEVELS + 1 times so that wideCounters[3] (levels that produced
// children) exceeds OHB_MAX_LEVELS exactly when the tree is too deep for the traversal stack.
static void launchCollapse(const BuildArrays& b, WideItem* itemsA, WideItem* itemsB, cudaStream_t st, uint64_t* launches) {
    k_wide_begin<<<1, 32, 0, st>>>(b, itemsA); (*launches)++;
    WideItem* in = itemsA; WideItem* out = itemsB; uint32_t* cin = b.wideCounters + 1; uint32_t* cout = b.wideCounters + 2;
    uint64_t cap = 1;
    for (int level = 0; level <= OHB_MAX_LEVELS; level++) {
        uint64_t bound = cap < uint64_t(b.n / 4u + 1u) ? cap : uint64_t(b.n / 4u + 1u);
        k_wide_level<<<gridFor(bound, 64), 64, 0, st>>>(b, in, cin, out, cout);
        k_wide_next<<<1, 32, 0, st>>>(cin, b.wideCounters + 3, cout);
        *launches += 2;
        WideItem* ti = in; in = out; out = ti; uint32_t* tc = cin; cin = cout; cout = tc;
        if (cap < (uint64_t(1) << 40)) cap *= 8;
    }
}
// primitives (wtri / primLo / primHi / boundsBits) -> sorted Morton codes -> Karras hierarchy -> refit -> treelets -> 8-wide collapse
static void launchBuildFromPrims(const BuildArrays& b, uint64_t* keysTmp, uint32_t* valsTmp, uint32_t* sortTemp, WideItem* itemsA, WideItem* itemsB,
                                 uint32_t treeletPasses, cudaStream_t st, uint64_t* launches) {
    const unsigned T = 256; unsigned g = gridFor(b.n, T);
    k_morton<<<g, T, 0, st>>>(b); (*launches)++;
    radixSort64(b.keys, b.vals, keysTmp, valsTmp, b.n, sortTemp, st, launches);
    if (b.n >= 2) {
        k_hierarchy<<<g, T, 0, st>>>(b);
        k_sweep<<<g, T, 0, st>>>(b, 0u);                        // refit + counts + SAH cost
        *launches += 2;
        for (uint32_t pass = 0, gamma = OHB_TREELET_LEAVES; pass < treeletPasses; pass++, gamma *= 2u) {
            k_clear_visit<<<g, T, 0, st>>>(b);
            k_sweep<<<gridFor(b.n, 64), 64, 0, st>>>(b, gamma);  // treelet restructuring (1.2 KB of DP tables per thread)
            *launches += 2;
        }
    }
    launchCollapse(b, itemsA, itemsB, st, launches);
}
void launchBuild(const BuildArrays& b, uint64_t* keysTmp, uint32_t* valsTmp, uint32_t* sortTemp, WideItem* itemsA, WideItem* itemsB,
                 uint32_t treeletPasses, cudaStream_t st, uint64_t* launches) {
    const unsigned T = 256; unsigned g = gridFor(b.n, T);
    k_init_build<<<g, T, 0, st>>>(b);
    k_world_tris<<<g, T, 0, st>>>(b);
    *launches += 2;
    launchBuildFromPrims(b, keysTmp, valsTmp, sortTemp, itemsA, itemsB, treeletPasses, st, launches);
}
// MODE_UPDATE of rt_acceleration_structure.cpp:508 for the flattened structure: the instance transforms changed, the topology
// of the binary tree is kept — re-transform the triangles, refit the boxes bottom-up, re-emit the wide nodes.  No Morton codes,
// no sort, no hierarchy, no treelets (the sorted order b.vals and the links b.left / b.right / parents stay as built).
void launchRefit(const BuildArrays& b, WideItem* itemsA, WideItem* itemsB, cudaStream_t st, uint64_t* launches) {
    const unsigned T = 256; unsigned g = gridFor(b.n, T);
    k_init_build<<<g, T, 0, st>>>(b);
    k_world_tris<<<g, T, 0, st>>>(b);
    *launches += 2;
    if (b.n >= 2) { k_sweep<<<g, T, 0, st>>>(b, 0u); (*launches)++; }
    launchCollapse(b, itemsA, itemsB, st, launches);
}
// one BLAS of the two-level structure: launchBuild in object space, then its root box -> blasLo/Hi[inst]
void launchBuildBlas(const BuildArrays& b, uint64_t* keysTmp, uint32_t* valsTmp, uint32_t* sortTemp, WideItem* itemsA, WideItem* itemsB, uint32_t treeletPasses,
                     f4* blasLo, f4* blasHi, uint32_t inst, uint32_t* maxLevels, cudaStream_t st, uint64_t* launches) {
    launchBuild(b, keysTmp, valsTmp, sortTemp, itemsA, itemsB, treeletPasses, st, launches);
    k_store_root_box<<<1, 32, 0, st>>>(b, blasLo, blasHi, inst, maxLevels); (*launches)++;
}
// TLAS build (refit == false) or MODE_UPDATE (refit == true: same instances, new transforms -> boxes refit, topology kept)
void launchBuildTlas(const BuildArrays& b, const f4* blasLo, const f4* blasHi, const uint32_t* instOfPrim, bool refit, uint64_t* keysTmp, uint32_t* valsTmp, uint32_t* sortTemp,
                     WideItem* itemsA, WideItem* itemsB, cudaStream_t st, uint64_t* launches) {
    const unsigned T = 256; unsigned g = gridFor(b.n, T);
    k_init_build<<<g, T, 0, st>>>(b);
    k_tlas_prims<<<g, T, 0, st>>>(b, blasLo, blasHi, instOfPrim);
    *launches += 2;
    if (!refit) { launchBuildFromPrims(b, keysTmp, valsTmp, sortTemp, itemsA, itemsB, 3u, st, launches); return; }
    if (b.n >= 2) { k_sweep<<<g, T, 0, st>>>(b, 0u); (*launches)++; }
    launchCollapse(b, itemsA, itemsB, st, launches);
}

// =============================================================================================
// Env CDF build — EnvCDF::build (env_cdf.cpp:13-61) in the reference's summation order:
// one thread per row runs the sequential fp32 prefix sum; one thread runs the marginal.
// =============================================================================================
__global__ void k_env_rows(const f4* env, uint32_t W, uint32_t H, float* cond, float* rowTotal) {
    uint32_t y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= H) return;
    const float pi = 3.14159265358979323846f;
    // fp32 argument like the reference; the sine itself is evaluated in fp64 and rounded once so it equals a
    // correctly rounded sinf (glibc's), which CUDA's 1-2 ulp sinf does not guarantee
    float sinTheta = float(sin(double(__fdiv_rn(__fmul_rn(pi, float(y) + 0.5f), float(H)))));
    float* row = cond + size_t(y) * W;
    float run = 0.0f;
    for (uint32_t x = 0; x < W; x++) {
        f4 p = env[size_t(y) * W + x];
        float lum = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, p.x), __fmul_rn(0.7152f, p.y)), __fmul_rn(0.0722f, p.z));
        run = __fadd_rn(run, __fmul_rn(lum, sinTheta));
        row[x] = run;
    }
    if (run > 0.0f) for (uint32_t x = 0; x < W; x++) row[x] = __fdiv_rn(row[x], run);
    else            for (uint32_t x = 0; x < W; x++) row[x] = float(x + 1) / float(W);
    rowTotal[y] = run;
}
__global__ void k_env_marginal(const float* rowTotal, uint32_t H, float* marg, float* integral) {
    if (blockIdx.x || threadIdx.x) return;
    float total = 0.0f;
    for (uint32_t y = 0; y < H; y++) { total = __fadd_rn(total, rowTotal[y]); marg[y] = total; }
    *integral = total;
    if (total > 0.0f) for (uint32_t y = 0; y < H; y++) marg[y] = __fdiv_rn(marg[y], total);
    else              for (uint32_t y = 0; y < H; y++) marg[y] = float(y + 1) / float(H);
}
// compact tables for cdfLowerBoundBlocked: every 32nd entry of the marginal and of each conditional row
__global__ void k_env_tops(const float* cond, const float* marg, uint32_t W, uint32_t H, float* condTop, float* margTop) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, wb = W >> 5, hb = H >> 5;
    if (i < H * wb) { const uint32_t y = i / wb, b = i % wb; condTop[i] = cond[size_t(y) * W + (b << 5) + 31u]; }
    if (i < hb) margTop[i] = marg[(i << 5) + 31u];
}
void launchEnvCdf(const f4* env, uint32_t W, uint32_t H, float* cond, float* marg, float* rowTotal, float* integral, float* condTop, float* margTop, cudaStream_t st, uint64_t* launches) {
    k_env_rows<<<gridFor(H, 64), 64, 0, st>>>(env, W, H, cond, rowTotal);
    k_env_marginal<<<1, 32, 0, st>>>(rowTotal, H, marg, integral);
    *launches += 2;
    if (condTop) { k_env_tops<<<gridFor(uint64_t(H) * (W >> 5), 128), 128, 0, st>>>(cond, marg, W, H, condTop, margTop); (*launches)++; }
}
__global__ void k_env_sample(SceneDev sc, const float* u12, uint32_t n, f4* dirPdf, float* pdfOfDir) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 d; float p; sampleEnvMap(sc, u12[2 * i], u12[2 * i + 1], d, p);
    dirPdf[i] = mk4(d, p);
    if (pdfOfDir) pdfOfDir[i] = pdfEnvMap(sc, d);
}
void launchEnvSample(const SceneDev& sc, const float* u12, uint32_t n, f4* dirPdf, float* pdfOfDir, cudaStream_t st, uint64_t* launches) {
    k_env_sample<<<gridFor(n, 128), 128, 0, st>>>(sc, u12, n, dirPdf, pdfOfDir); (*launches)++;
}

__global__ void k_env_pdf(SceneDev sc, const float* dirs3, uint32_t n, float* pdf) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pdf[i] = pdfEnvMap(sc, mk3(dirs3[3 * i], dirs3[3 * i + 1], dirs3[3 * i + 2]));
}
void launchEnvPdf(const SceneDev& sc, const float* dirs3, uint32_t n, float* pdf, cudaStream_t st, uint64_t* launches) {
    k_env_pdf<<<gridFor(n, 128), 128, 0, st>>>(sc, dirs3, n, pdf); (*launches)++;
}

// Hybrid-RT techniques (ohb_hybrid.h): one thread per pixel, 8x8-pixel blocks so that a warp covers an 8x4 patch of neighbouring rays
__global__ void __launch_bounds__(64) k_hybrid_shadow(SceneDev sc, HybridShadowParams pc, const f4* gPos, const f2* gNrm, uint8_t* mask) {
    uint32_t x = blockIdx.x * 8u + (threadIdx.x & 7u), y = blockIdx.y * 8u + (threadIdx.x >> 3);
    if (x < pc.W && y < pc.H) mask[size_t(y) * pc.W + x] = hybridShadowPixel(sc, pc, gPos, gNrm, x, y);
}
__global__ void __launch_bounds__(64) k_hybrid_gi(SceneDev sc, HybridGiParams pc, const f4* gPos, const f2* gNrm, const f4* gAlbedo, const f4* history, const f4* instMat, h4* out) {
    uint32_t x = blockIdx.x * 8u + (threadIdx.x & 7u), y = blockIdx.y * 8u + (threadIdx.x >> 3);
    if (x < pc.W && y < pc.H) out[size_t(y) * pc.W + x] = hybridGiPixel(sc, pc, gPos, gNrm, gAlbedo, history, instMat, x, y);
}
void launchHybridShadow(const SceneDev& sc, const HybridShadowParams& pc, const f4* gPos, const f2* gNrm, uint8_t* mask, cudaStream_t st, uint64_t* launches) {
    k_hybrid_shadow<<<dim3((pc.W + 7u) / 8u, (pc.H + 7u) / 8u), 64, 0, st>>>(sc, pc, gPos, gNrm, mask); (*launches)++;
}
void launchHybridGi(const SceneDev& sc, const HybridGiParams& pc, const f4* gPos, const f2* gNrm, const f4* gAlbedo, const f4* history, const f4* instMat, h4* out, cudaStream_t st, uint64_t* launches) {
    k_hybrid_gi<<<dim3((pc.W + 7u) / 8u, (pc.H + 7u) / 8u), 64, 0, st>>>(sc, pc, gPos, gNrm, gAlbedo, history, instMat, out); (*launches)++;
}

// NRD front-end packing hook (nrd_frontend.glsl:11-41): in6 = (radiance.rgb, hitDist, viewZ, roughness), nr4 = (normal, roughness)
__global__ void k_nrd_pack(const float* in6, const float* nr4, uint32_t n, f4* packedRad, f4* packedNormal, float* unpackedRgb) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = in6 + size_t(i) * 6u;
    f4 pr = nrdPackRadianceHitDist(mk3(a[0], a[1], a[2]), a[3], a[4], a[5]);
    packedRad[i] = pr;
    packedNormal[i] = nrdPackNormalRoughness(mk3(nr4[4 * i], nr4[4 * i + 1], nr4[4 * i + 2]), nr4[4 * i + 3]);
    f3 back = nrdYCoCgToLinear(xyz(pr));
    unpackedRgb[3 * i] = back.x; unpackedRgb[3 * i + 1] = back.y; unpackedRgb[3 * i + 2] = back.z;
}
void launchNrdPack(const float* in6, const float* nr4, uint32_t n, f4* packedRad, f4* packedNormal, float* unpackedRgb, cudaStream_t st, uint64_t* launches) {
    k_nrd_pack<<<gridFor(n, 128), 128, 0, st>>>(in6, nr4, n, packedRad, packedNormal, unpackedRgb); (*launches)++;
}

// =============================================================================================
// Traversal kernels (persistent, warp-granular dynamic fetch)
// =============================================================================================
#define TRACE_THREADS 128
// Scheduling knobs of the traversal kernels (the env overrides exist for A/B measurements on the GPU box):
//   minActive   : pause traversal and refill idle lanes when fewer lanes than this still traverse   (OHB_TRACE_MIN_ACTIVE)
//   postponeDen : travRun's triangle-postponing threshold                                            (OHB_POSTPONE_DEN)
#define OHB_TRACE_MIN_ACTIVE_DEFAULT 20
struct TraceKnobs { int minActive, postponeDen; };
static TraceKnobs traceKnobs() {
    static const TraceKnobs k = []() {
        auto geti = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
        TraceKnobs r; r.minActive = geti("OHB_TRACE_MIN_ACTIVE", OHB_TRACE_MIN_ACTIVE_DEFAULT); r.postponeDen = geti("OHB_POSTPONE_DEN", OHB_POSTPONE_DEN_DEFAULT);
        if (r.postponeDen == 1) r.postponeDen = 2;      // 1 would postpone every triangle forever; 0 = never postpone
        r.postponeDen = (r.postponeDen & 0xFF) | (geti("OHB_TRACE_PREFETCH", 0) << 8);      // prefetch mode rides in bits 8.. (travRun)
        return r;
    }();
    return k;
}
// Persistent warp loop shared by every traversal kernel.  Lanes pull rays one by one from a global counter
// (warp-aggregated atomic); a lane that finishes its ray waits at the reconvergence point of the `have`
// block until the traversing lanes either finish or drop below TRACE_MIN_ACTIVE, then all idle lanes are
// refilled together.  IO = { load(i, o, d, tmin, tmax), store(i, Trav&) }.
// VAR (compile-time A/B variants, selected by OHB_TRACE_VAR): bit 0 = the first OHB_SMEM_STACK stack entries of every thread
// in shared memory, bit 1 = the top OHB_SMEM_TOP wide nodes (levels 0-2 of the tree) staged in shared memory per CTA.
#define OHB_SMEM_STACK 8
#define OHB_SMEM_TOP 73u
template <bool ANY, int VAR, class IO, bool TL = false>
__device__ __forceinline__ void persistentTrace(const SceneDev& sc, uint32_t n, uint32_t* work, IO& io, int minActive, int postponeDen, const uint32_t* perm = nullptr) {
    // VAR bit 2: the default scheduling knobs as compile-time constants — the per-iteration loads and tests of the two kernel
    // parameters go away (launchTraceClosest / launchTraceShadow pick this variant when the knobs are at their defaults)
    if (VAR & 4) { minActive = OHB_TRACE_MIN_ACTIVE_DEFAULT; postponeDen = OHB_POSTPONE_DEN_DEFAULT; }
    const uint32_t lane = threadIdx.x & 31u;
    __shared__ TravStackEntry sStack[(VAR & 1) ? OHB_SMEM_STACK * TRACE_THREADS : 1];
    __shared__ u4 sTop[(VAR & 2) ? OHB_SMEM_TOP * OHB_WNODE_VECS : 1];
    TopNodes top{nullptr, 0u};
    if (VAR & 2) {
        const uint32_t cnt = sc.numWideNodes < OHB_SMEM_TOP ? sc.numWideNodes : OHB_SMEM_TOP;
        for (uint32_t i = threadIdx.x; i < cnt * OHB_WNODE_VECS; i += TRACE_THREADS) sTop[i] = sc.wnodes[i];
        __syncthreads();
        top.top = sTop; top.count = cnt;
    }
    typename std::conditional<(VAR & 1) != 0, SharedStack<OHB_SMEM_STACK, TRACE_THREADS>, ArrayStack>::type stack;
    TravStackEntry localMem[(VAR & 1) ? 1 : OHB_STACK_SIZE];
    if constexpr ((VAR & 1) != 0) stack.sh = sStack + threadIdx.x; else stack.p = localMem;
    Trav t; uint32_t idx = 0; bool have = false, exhausted = false;
    for (;;) {
        unsigned need = __ballot_sync(0xffffffffu, !have && !exhausted);
        if (need) {
            uint32_t base = 0;
            if (lane == uint32_t(__ffs(int(need)) - 1)) base = atomicAdd(work, uint32_t(__popc(need)));
            base = __shfl_sync(0xffffffffu, base, __ffs(int(need)) - 1);
            if (!have && !exhausted) {
                idx = base + uint32_t(__popc(need & ((1u << lane) - 1u)));
                if (idx < n) {
                    if (perm) idx = perm[idx];                   // queue slot of the idx-th ray in (octant, queue) order
                    f3 o, d; float tmin, tmax; io.load(idx, o, d, tmin, tmax); travInit(t, sc, o, d, tmin, tmax); have = true;
                }
                else exhausted = true;
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        bool done = false;
        if (have) done = travRun<ANY, decltype(stack), TL>(t, stack, sc, top, minActive, postponeDen);
        __syncwarp();
        if (done) { io.store(idx, t); have = false; }
    }
}

struct PathClosestIO {
    PathArrays P; uint32_t hits;
    __device__ __forceinline__ void load(uint32_t i, f3& o, f3& d, float& tmin, float& tmax) {
        uint32_t p = OHB_Q_PATH(P.queueIn[i]); o = xyz(P.rayO[p]); d = xyz(P.rayD[p]); tmin = 0.001f; tmax = 10000.0f;
    }
    __device__ __forceinline__ void store(uint32_t i, Trav& t) {
        uint32_t p = OHB_Q_PATH(P.queueIn[i]);
        const ohb_hit h = travResult(t);
        const bool hit = h.prim != OHB_MISS;
        reinterpret_cast<float4*>(P.hit)[p] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
        P.hitFlag[i] = hit ? 1u : 0u;
        hits += hit;
    }
};
template <int MINB, int VAR>
__global__ void __launch_bounds__(TRACE_THREADS, MINB) k_trace_closest(SceneDev sc, PathArrays P, uint32_t* work, int minActive, int postponeDen) {
    PathClosestIO io{P, 0u};
    persistentTrace<false, VAR>(sc, *P.countIn, work, io, minActive, postponeDen, P.octPerm);
    uint32_t hits = __reduce_add_sync(0xffffffffu, io.hits);
    if ((threadIdx.x & 31u) == 0 && hits) atomicAdd(P.counters + 3, (unsigned long long)hits);
}
struct PathShadowIO {
    PathArrays P;
    __device__ __forceinline__ void load(uint32_t i, f3& o, f3& d, float& tmin, float& tmax) {
        f4 a = P.shO[i], b = P.shD[i]; o = xyz(a); d = xyz(b); tmin = 0.001f; tmax = a.w;
    }
    __device__ __forceinline__ void store(uint32_t i, Trav& t) {
        if (!t.anyHit) return;
        uint32_t tag = __float_as_uint(P.shD[i].w);
        f4* pend = (tag & 1u) ? P.pendB : P.pendA;
        pend[tag >> 1] = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    }
};
template <int MINB, int VAR>
__global__ void __launch_bounds__(TRACE_THREADS, MINB) k_trace_shadow(SceneDev sc, PathArrays P, uint32_t* work, int minActive, int postponeDen) {
    PathShadowIO io{P};
    persistentTrace<true, VAR>(sc, *P.shCount, work, io, minActive, postponeDen, P.octPerm);
}
// ---- stable octant binning of a ray queue (PathArrays::octPerm): one pass of the LSD radix sort above with the ray octant as the
// digit (8 bins) and the queue slot as the value; the queue length is read on the device, the grid covers the capacity -------------
#ifndef OHB_OCT_BIN_DEFAULT
#define OHB_OCT_BIN_DEFAULT 0
#endif
__device__ __forceinline__ uint32_t dirOctant(f4 d) {   // == octantOf(prepRay(...)): bit set = the ray travels toward + on that axis
    return ((__float_as_uint(d.x) >> 31) ? 0u : 4u) | ((__float_as_uint(d.y) >> 31) ? 0u : 2u) | ((__float_as_uint(d.z) >> 31) ? 0u : 1u);
}
template <bool SHADOW>
__device__ __forceinline__ uint32_t queueOctant(const PathArrays& P, uint32_t i) { return dirOctant(SHADOW ? P.shD[i] : P.rayD[OHB_Q_PATH(P.queueIn[i])]); }
template <bool SHADOW>
__global__ void __launch_bounds__(RS_THREADS) k_oct_hist(PathArrays P) {
    __shared__ uint32_t h[8];
    const uint32_t n = SHADOW ? *P.shCount : *P.countIn;
    if (threadIdx.x < 8u) h[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    if (base < n) {
        for (uint32_t i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
            const uint32_t idx = base + i;
            const uint32_t oct = idx < n ? queueOctant<SHADOW>(P, idx) : 8u;
            const unsigned peers = __match_any_sync(0xffffffffu, oct);
            if (oct < 8u && (threadIdx.x & 31u) == uint32_t(__ffs(int(peers)) - 1)) atomicAdd(&h[oct], uint32_t(__popc(peers)));
        }
    }
    __syncthreads();
    if (threadIdx.x < 8u) P.octHist[threadIdx.x * P.octBlocks + blockIdx.x] = h[threadIdx.x];
}
template <bool SHADOW>
__global__ void __launch_bounds__(RS_THREADS) k_oct_scatter(PathArrays P) {
    __shared__ uint32_t wh[RS_WARPS][8];
    const uint32_t n = SHADOW ? *P.shCount : *P.countIn;
    if (blockIdx.x * RS_TILE >= n) return;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (threadIdx.x < RS_WARPS * 8) (&wh[0][0])[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t chunk = blockIdx.x * RS_TILE + warp * (32u * RS_ROUNDS);
    uint32_t octs = 0u;                                           // this lane's 16 octants, 4 bits each (8 = past the end)... two words
    uint32_t octs2 = 0u;
    for (int r = 0; r < RS_ROUNDS; r++) {
        const uint32_t idx = chunk + uint32_t(r) * 32u + lane;
        const uint32_t oct = idx < n ? queueOctant<SHADOW>(P, idx) : 8u;
        if (r < 8) octs |= oct << (4 * r); else octs2 |= oct << (4 * (r - 8));
        const unsigned peers = __match_any_sync(0xffffffffu, oct);
        if (oct < 8u && lane == uint32_t(__ffs(int(peers)) - 1)) wh[warp][oct] += uint32_t(__popc(peers));
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 8u) {
        const uint32_t d = threadIdx.x;
        uint32_t run = P.octHist[d * P.octBlocks + blockIdx.x];    // exclusive scan of the (octant-major, block-minor) histogram
        for (int w = 0; w < RS_WARPS; w++) { uint32_t t = wh[w][d]; wh[w][d] = run; run += t; }
    }
    __syncthreads();
    for (int r = 0; r < RS_ROUNDS; r++) {
        const uint32_t idx = chunk + uint32_t(r) * 32u + lane;
        const uint32_t oct = r < 8 ? (octs >> (4 * r)) & 15u : (octs2 >> (4 * (r - 8))) & 15u;
        const unsigned peers = __match_any_sync(0xffffffffu, oct);
        if (oct < 8u) P.octPerm[wh[warp][oct] + uint32_t(__popc(peers & ((1u << lane) - 1u)))] = idx;
        __syncwarp();
        if (oct < 8u && lane == uint32_t(__ffs(int(peers)) - 1)) wh[warp][oct] += uint32_t(__popc(peers));
        __syncwarp();
    }
}
static bool octBinOn() { static const bool v = []() { const char* e = getenv("OHB_OCT_BIN"); return e ? atoi(e) != 0 : OHB_OCT_BIN_DEFAULT != 0; }(); return v; }
template <bool SHADOW>
static void launchOctBin(const PathArrays& P, cudaStream_t st, uint64_t* launches) {
    k_oct_hist<SHADOW><<<P.octBlocks, RS_THREADS, 0, st>>>(P); (*launches)++;
    exclusiveScanU32(P.octHist, 8u * P.octBlocks, P.octScanTmp, st, launches);
    k_oct_scatter<SHADOW><<<P.octBlocks, RS_THREADS, 0, st>>>(P); (*launches)++;
}
// two-level variants (ohb_set_accel_mode(OHB_ACCEL_TWO_LEVEL)): the same persistent loop, travRun with the instance descent
__global__ void __launch_bounds__(TRACE_THREADS, 6) k_trace_closest_tl(SceneDev sc, PathArrays P, uint32_t* work, int minActive, int postponeDen) {
    PathClosestIO io{P, 0u};
    persistentTrace<false, 0, PathClosestIO, true>(sc, *P.countIn, work, io, minActive, postponeDen);
    uint32_t hits = __reduce_add_sync(0xffffffffu, io.hits);
    if ((threadIdx.x & 31u) == 0 && hits) atomicAdd(P.counters + 3, (unsigned long long)hits);
}
__global__ void __launch_bounds__(TRACE_THREADS, 6) k_trace_shadow_tl(SceneDev sc, PathArrays P, uint32_t* work, int minActive, int postponeDen) {
    PathShadowIO io{P};
    persistentTrace<true, 0, PathShadowIO, true>(sc, *P.shCount, work, io, minActive, postponeDen);
}
// resident CTAs per SM the traversal kernels are compiled for: 5 -> 96 registers, 6 -> 80, 7 -> 72, 8 -> 64, 9 -> 56, 10 -> 48.
#ifndef OHB_TRACE_OCC_DEFAULT
#define OHB_TRACE_OCC_DEFAULT 8
#endif
#ifndef OHB_TRACE_VAR_DEFAULT
#define OHB_TRACE_VAR_DEFAULT 0
#endif
static int traceOcc() {
    static const int v = []() { const char* e = getenv("OHB_TRACE_OCC"); int o = e ? atoi(e) : OHB_TRACE_OCC_DEFAULT; return o >= 10 ? 10 : (o <= 5 ? 5 : o); }();
    return v;
}
static int traceVar() {
    static const int v = []() { const char* e = getenv("OHB_TRACE_VAR"); int o = e ? atoi(e) : OHB_TRACE_VAR_DEFAULT; return o & 3; }();
    return v;
}
// occupancy variants 5..10 exist for VAR 0; the shared-memory variants (VAR 1..3) are built for 7 and 8 CTAs per SM
#define OHB_TRACE_DISPATCH(KERNEL, ...) do { \
    const int o = traceOcc(), v = traceVar(); const unsigned grid = smGrid8 / 8u * unsigned(v ? (o >= 8 ? 8 : 7) : o); \
    if (v == 0) { \
        if (o == 10) KERNEL<10, 0><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); else if (o == 9) KERNEL<9, 0><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); \
        else if (o == 8) KERNEL<8, 0><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); else if (o == 7) KERNEL<7, 0><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); \
        else if (o == 6) KERNEL<6, 0><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); else KERNEL<5, 0><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); \
    } else if (o >= 8) { \
        if (v == 1) KERNEL<8, 1><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); else if (v == 2) KERNEL<8, 2><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); else KERNEL<8, 3><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); \
    } else { \
        if (v == 1) KERNEL<7, 1><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); else if (v == 2) KERNEL<7, 2><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); else KERNEL<7, 3><<<grid, TRACE_THREADS, 0, st>>>(__VA_ARGS__); \
    } } while (0)
static void launchTraceClosest(unsigned smGrid8, const SceneDev& sc, const PathArrays& P, uint32_t* work, cudaStream_t st) {
    const TraceKnobs k = traceKnobs();
    if (sc.twoLevel) { k_trace_closest_tl<<<smGrid8 / 8u * 6u, TRACE_THREADS, 0, st>>>(sc, P, work, k.minActive, k.postponeDen & 0xFF); return; }
    if (traceOcc() == 8 && traceVar() == 0 && k.minActive == OHB_TRACE_MIN_ACTIVE_DEFAULT && k.postponeDen == OHB_POSTPONE_DEN_DEFAULT) {
        k_trace_closest<8, 4><<<smGrid8, TRACE_THREADS, 0, st>>>(sc, P, work, k.minActive, k.postponeDen); return;
    }
    OHB_TRACE_DISPATCH(k_trace_closest, sc, P, work, k.minActive, k.postponeDen);
}
static void launchTraceShadow(unsigned smGrid8, const SceneDev& sc, const PathArrays& P, uint32_t* work, cudaStream_t st) {
    const TraceKnobs k = traceKnobs();
    if (sc.twoLevel) { k_trace_shadow_tl<<<smGrid8 / 8u * 6u, TRACE_THREADS, 0, st>>>(sc, P, work, k.minActive, k.postponeDen & 0xFF); return; }
    if (traceOcc() == 8 && traceVar() == 0 && k.minActive == OHB_TRACE_MIN_ACTIVE_DEFAULT && k.postponeDen == OHB_POSTPONE_DEN_DEFAULT) {
        k_trace_shadow<8, 4><<<smGrid8, TRACE_THREADS, 0, st>>>(sc, P, work, k.minActive, k.postponeDen); return;
    }
    OHB_TRACE_DISPATCH(k_trace_shadow, sc, P, work, k.minActive, k.postponeDen);
}
#define SHADE_THREADS 128
#ifndef OHB_SHADE_PREFETCH_DEFAULT
#define OHB_SHADE_PREFETCH_DEFAULT 0
#endif
// ---- queue emission, STABLE AT TILE GRANULARITY ------------------------------------------------------------------
// A CTA walks a contiguous tile of QTILE queue entries (QTILE_ROUNDS rounds of SHADE_THREADS), stages what it emits in
// shared memory in queue order and claims its output range with ONE atomic per tile.  With per-warp atomics every pass
// cuts the queue into 32-entry chunks that land in arrival order; a few iterations later the 32 paths a warp reads come
// from many places, the 16-B path-record accesses stop sharing sectors and the shading kernels — which live on that
// locality — slow down.  Measured (profiles/r1z_sweep.txt): helmet 1034 -> 1258, Cornell 674 -> 894, 2 M 215 -> 230
// Msamples/s against the same kernels with warp-granular emission.
#define QTILE_ROUNDS 8
#define QTILE (SHADE_THREADS * QTILE_ROUNDS)
struct TileEmit {
    uint32_t* sOut; uint32_t* sW; uint32_t* sBase; uint32_t kept;
    __device__ __forceinline__ void begin() { kept = 0u; }
    // every thread of the CTA calls this once per round (keep = this thread emits e)
    __device__ __forceinline__ void round(bool keep, uint32_t e) {
        const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0u) sW[warp] = uint32_t(__popc(m));
        __syncthreads();
        uint32_t off = kept, tot = 0u;
#pragma unroll
        for (uint32_t w = 0; w < SHADE_THREADS / 32; w++) { if (w < warp) off += sW[w]; tot += sW[w]; }
        if (keep) sOut[off + uint32_t(__popc(m & ((1u << lane) - 1u)))] = e;
        kept += tot;
        __syncthreads();
    }
    // forward = true: queue[base + j]; false: queue[last - (base + j)] (the back-to-front half of a two-ended queue)
    __device__ __forceinline__ void flush(uint32_t* counter, uint32_t* queue, bool forward, uint32_t last) {
        if (threadIdx.x == 0u) *sBase = kept ? atomicAdd(counter, kept) : 0u;
        __syncthreads();
        const uint32_t base = *sBase;
        for (uint32_t j = threadIdx.x; j < kept; j += SHADE_THREADS) queue[forward ? base + j : last - (base + j)] = sOut[j];
        __syncthreads();
    }
};
// one round of TWO emitters behind one pair of barriers (k_sort_hits: hits and misses of the same entries)
__device__ __forceinline__ void tileEmitRound2(TileEmit& a, bool keepA, uint32_t eA, TileEmit& b, bool keepB, uint32_t eB) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned ma = __ballot_sync(0xffffffffu, keepA), mb = __ballot_sync(0xffffffffu, keepB);
    if (lane == 0u) { a.sW[warp] = uint32_t(__popc(ma)); b.sW[warp] = uint32_t(__popc(mb)); }
    __syncthreads();
    uint32_t offA = a.kept, totA = 0u, offB = b.kept, totB = 0u;
#pragma unroll
    for (uint32_t w = 0; w < SHADE_THREADS / 32; w++) { const uint32_t ca = a.sW[w], cb = b.sW[w]; if (w < warp) { offA += ca; offB += cb; } totA += ca; totB += cb; }
    if (keepA) a.sOut[offA + uint32_t(__popc(ma & ((1u << lane) - 1u)))] = eA;
    if (keepB) b.sOut[offB + uint32_t(__popc(mb & ((1u << lane) - 1u)))] = eB;
    a.kept += totA; b.kept += totB;
    __syncthreads();
}
#define OHB_TILE_EMIT(name) __shared__ uint32_t name##_out[QTILE], name##_w[SHADE_THREADS / 32], name##_base; TileEmit name{name##_out, name##_w, &name##_base, 0u}

template <bool RT>
__global__ void __launch_bounds__(SHADE_THREADS) k_raygen(FrameParams fr, PathArrays P, uint32_t total) {
    OHB_TILE_EMIT(q);
    const uint32_t tile = blockIdx.x * QTILE;
    q.begin();
    for (int r = 0; r < QTILE_ROUNDS; r++) {
        const uint32_t p = tile + uint32_t(r) * SHADE_THREADS + threadIdx.x;
        bool keep = false;
        if (p < total) { if (RT) raygenPathRT(fr, P, p); else raygenPath(fr, P, p); keep = OHB_ST_STAGE(P.meta[p].w) != ST_DONE; }
        q.round(keep, RT ? p : (p | OHB_Q_PRIMARY));
    }
    q.flush(P.countIn, P.queueIn, true, 0u);
}
// hits to the front of queueSorted, misses (flagged) to the back, in queue order: the shading kernels then run hit-only
// and miss-only warps (35 % of the helmet scene's bounce rays miss).  It reads k_trace_closest's per-slot hit flags;
// emitting from the trace kernel's own store would scatter the path indices (rays finish in arbitrary order, r1i).
__global__ void __launch_bounds__(SHADE_THREADS) k_sort_hits(PathArrays P) {
    OHB_TILE_EMIT(qh); OHB_TILE_EMIT(qm);
    const uint32_t n = *P.countIn;
    for (uint32_t tile = blockIdx.x * QTILE; tile < n; tile += gridDim.x * QTILE) {
        qh.begin(); qm.begin();
        // all of the tile's entries and flags are requested before the first round: the rounds are separated by barriers, and
        // loading inside them made every round wait for its own memory round trip (eight in a row per tile)
        uint32_t ent[QTILE_ROUNDS]; uint8_t flg[QTILE_ROUNDS];
#pragma unroll
        for (int r = 0; r < QTILE_ROUNDS; r++) {
            const uint32_t i = tile + uint32_t(r) * SHADE_THREADS + threadIdx.x;
            ent[r] = i < n ? P.queueIn[i] : 0u; flg[r] = i < n ? P.hitFlag[i] : uint8_t(0);
        }
#pragma unroll
        for (int r = 0; r < QTILE_ROUNDS; r++) {
            const uint32_t i = tile + uint32_t(r) * SHADE_THREADS + threadIdx.x;
            const bool valid = i < n, hit = valid && flg[r] != 0u;
            tileEmitRound2(qh, hit, ent[r], qm, valid && !hit, ent[r] | OHB_Q_MISS);
        }
        qh.flush(P.sortCount, P.queueSorted, true, 0u); qm.flush(P.sortCount + 1, P.queueSorted, false, n - 1u);
    }
}
// k_surface: closest-hit / miss shaders of the sorted queue -> payload records (realtime profile; the offline profile fuses it into k_shade)
__global__ void __launch_bounds__(SHADE_THREADS) k_surface(SceneDev sc, FrameParams fr, PathArrays P) {
    const uint32_t n = *P.countIn;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        surfacePath(sc, fr, P, OHB_Q_PATH(P.queueSorted[i]));
}
// k_shade: closest-hit / miss shaders + the raygen's per-bounce body in one kernel, the payload stays in registers
// (shadePath): 376 B of DRAM traffic per path on the 2 M scene against 831 B for a k_surface + k_bounce pair
// (profiles/r1x_ncu_full_synthetic2m_kshade.txt).  The shading kernels are instruction-FETCH limited — straight-line
// bodies of 2-3x the 32 KB L1.5 instruction cache, ncu stall_no_instruction 4-6 of 8 warps per scheduler — so the
// fused kernel only pays off with the code-size work (shared sampler / texture / env-lookup / anisotropy copies, MUFU
// reciprocals: 5 800 -> 3 900 SASS instructions) and with hit-only / miss-only warps (k_sort_hits): profiles/r1p, r1u,
// r1v, r1z sweeps record each step.
// Software prefetch (pf: 0 off, 1 into L2, 2 into L1; OHB_SHADE_PREFETCH).  k_shade waits on memory, not on issue slots (r2m: long
// scoreboard 12-31 warps per issue, issue slots 17-36 % busy, DRAM 20 %): a path's shading is a chain of dependent gathers —
// queue entry -> hit record -> triangle -> vertices / material -> texels, then the path record for the bounce.  Prefetches
// carry no register and no scoreboard, so they widen that chain for free: the bounce body's record of THIS path and the
// ray + hit record of the path this thread shades NEXT round are requested before the closest-hit shader starts, and the
// next path's triangle-level data (indices, instance, material id) once its primitive id has arrived.
__device__ __forceinline__ void shadePrefetch(const void* p, int pf) {
    if (pf == 2) asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); else asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}
// LEAN: the frame uses the Sobol sampler, no anisotropy, no subsurface term (every reference preset) — known at launch, folded at
// compile time so the PCG / anisotropic-GGX / SSS code leaves the kernel
template <int MINB, int PRIM, bool LEAN = false>
__global__ void __launch_bounds__(SHADE_THREADS, MINB) k_shade(SceneDev sc, FrameParams fr, PathArrays P, int pf) {
    if (LEAN) { fr.samplerType = OHB_SAMPLER_SOBOL; fr.aniso = 0.0f; fr.anisoRot = 0.0f; fr.sss = 0.0f; }
    OHB_TILE_EMIT(q);
    const uint32_t n = *P.countIn;
    for (uint32_t tile = blockIdx.x * QTILE; tile < n; tile += gridDim.x * QTILE) {
        q.begin();
        const uint32_t i0 = tile + threadIdx.x;
        uint32_t eNext = i0 < n ? P.queueSorted[i0] : OHB_Q_NONE;
        for (int r = 0; r < QTILE_ROUNDS; r++) {
            const uint32_t eCur = eNext;
            const uint32_t iN = tile + uint32_t(r + 1) * SHADE_THREADS + threadIdx.x;
            eNext = (r + 1 < QTILE_ROUNDS && iN < n) ? P.queueSorted[iN] : OHB_Q_NONE;
            uint32_t primN = OHB_MISS;
            if (pf) {
                if (eCur != OHB_Q_NONE) {
                    const uint32_t pc = OHB_Q_PATH(eCur);
                    shadePrefetch(P.meta + pc, pf); shadePrefetch(P.rad + pc, pf);
                    if (!(eCur & OHB_Q_PRIMARY)) shadePrefetch(P.thr + pc, pf);
                    if (eCur & OHB_Q_PEND_A) shadePrefetch(P.pendA + pc, pf);
                    if (eCur & OHB_Q_PEND_B) shadePrefetch(P.pendB + pc, pf);
                }
                if (eNext != OHB_Q_NONE) {
                    const uint32_t pn = OHB_Q_PATH(eNext);
                    shadePrefetch(P.rayO + pn, pf); shadePrefetch(P.rayD + pn, pf);
                    primN = P.hit[pn].prim;                       // a real load: needed (much later) for the second-level prefetch
                }
            }
            uint32_t e = OHB_Q_NONE;
            if (eCur != OHB_Q_NONE) e = shadePath<PRIM>(sc, fr, P, eCur);
            if (pf && primN != OHB_MISS) { shadePrefetch(sc.indices + size_t(primN) * 3u, pf); shadePrefetch(sc.triInst + primN, pf); shadePrefetch(sc.matIds + primN, pf); }
            q.round(e != OHB_Q_NONE, e);
        }
        q.flush(P.countOut, P.queueOut, true, 0u);
    }
}
// Between iterations: account the rays just traced, clear the queues that are about to be refilled.
__global__ void k_advance(PathArrays P, uint32_t* work, int afterRaygen) {
    if (blockIdx.x || threadIdx.x) return;
    if (afterRaygen) { P.counters[0] += *P.countIn; }
    else { P.counters[1] += *P.countIn; P.counters[2] += *P.shCount; *P.countIn = 0u; }
    *P.shCount = 0u; work[0] = 0u; work[1] = 0u; P.sortCount[0] = 0u; P.sortCount[1] = 0u;
}
__global__ void k_sobol_tab(u4* tab, uint32_t first, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tab[i] = sobolQuad(first + i);
}
__global__ void k_zero_u32(uint32_t* p, uint32_t n) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = 0u; }
__global__ void k_film(FrameParams fr, PathArrays P, FilmArrays F) {
    uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix < P.numPixels) filmPixel(fr, P, F, pix);
}
__global__ void k_resolve(f4* accum, uint32_t* ldr, uint32_t n, int sumMode) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f4 a = accum[i];
    f3 mean = sumMode ? xyz(a) / fmaxf(a.w, 1.0f) : xyz(a);
    ldr[i] = tonemapRGBA8(mean);
}

// One batch of `samplesInBatch` samples for every pixel of the tile: raygen, (1 + 2*maxBounces)
// wavefront iterations, film.  All launches are asynchronous on `st`.
// phases: bit 0 = raygen + wavefront iterations, bit 1 = film (ohb_render runs the two lanes' films in sample order)
void launchOfflineBatch(const SceneDev& sc, const FrameParams& fr, PathArrays P, const FilmArrays& F,
                        uint32_t* work, int numSMs, cudaStream_t st, uint64_t* launches, TimingHooks* th, int phases) {
    if (!(phases & 1)) {
        k_film<<<gridFor(P.numPixels, 128), 128, 0, st>>>(fr, P, F); (*launches)++;
        return;
    }
    uint32_t total = P.numPixels * P.samplesInBatch;
    k_zero_u32<<<1, 32, 0, st>>>(P.countIn, 1); k_zero_u32<<<1, 32, 0, st>>>(P.countOut, 1);
    k_sobol_tab<<<gridFor(P.samplesInBatch, 64), 64, 0, st>>>(const_cast<u4*>(P.sobolTab), P.firstSampleIndex, P.samplesInBatch); (*launches)++;
    k_raygen<false><<<gridFor(total, QTILE), SHADE_THREADS, 0, st>>>(fr, P, total);
    k_advance<<<1, 32, 0, st>>>(P, work, 1);
    *launches += 4;
    unsigned traceGrid = unsigned(numSMs) * 8u;
    unsigned shadeGrid = unsigned(numSMs) * 8u;
    uint32_t iters = 1u + 2u * fr.maxBounces;
    static const int bounceOcc = []() { const char* e = getenv("OHB_BOUNCE_OCC"); return e ? atoi(e) : 8; }();
    const bool bin = octBinOn() && !sc.twoLevel && P.octPerm != nullptr;
    if (!bin) P.octPerm = nullptr;
    static const bool leanOn = []() { const char* e = getenv("OHB_SHADE_LEAN"); return e ? atoi(e) != 0 : true; }();
    const bool lean = leanOn && fr.samplerType == OHB_SAMPLER_SOBOL && fr.aniso == 0.0f && fr.sss == 0.0f;
    static const bool shadeStages = []() { const char* e = getenv("OHB_SHADE_STAGES"); return e ? atoi(e) != 0 : true; }();
    static const int shadePf = []() { const char* e = getenv("OHB_SHADE_PREFETCH"); int v = e ? atoi(e) : OHB_SHADE_PREFETCH_DEFAULT; return v < 0 ? 0 : (v > 2 ? 2 : v); }();
    for (uint32_t it = 0; it < iters; it++) {
        if (th) th->begin(0, st);
        if (bin) launchOctBin<false>(P, st, launches);
        launchTraceClosest(traceGrid, sc, P, work, st);
        if (th) th->end(0, st);
        if (th) th->begin(7, st);
        k_sort_hits<<<shadeGrid, SHADE_THREADS, 0, st>>>(P);
        if (th) { th->end(7, st); th->begin(1, st); }
        // stage-specialised bodies: iteration 0 shades camera rays only, every later iteration none (OHB_SHADE_STAGES=0: one generic body)
        if (!shadeStages) { if (bounceOcc >= 8) k_shade<8, 2><<<shadeGrid, SHADE_THREADS, 0, st>>>(sc, fr, P, shadePf); else k_shade<6, 2><<<shadeGrid, SHADE_THREADS, 0, st>>>(sc, fr, P, shadePf); }
        else if (lean && bounceOcc >= 8) { if (it == 0) k_shade<8, 1, true><<<shadeGrid, SHADE_THREADS, 0, st>>>(sc, fr, P, shadePf); else k_shade<8, 0, true><<<shadeGrid, SHADE_THREADS, 0, st>>>(sc, fr, P, shadePf); }
        else if (it == 0) { if (bounceOcc >= 8) k_shade<8, 1><<<shadeGrid, SHADE_THREADS, 0, st>>>(sc, fr, P, shadePf); else k_shade<6, 1><<<shadeGrid, SHADE_THREADS, 0, st>>>(sc, fr, P, shadePf); }
        else { if (bounceOcc >= 8) k_shade<8, 0><<<shadeGrid, SHADE_THREADS, 0, st>>>(sc, fr, P, shadePf); else k_shade<6, 0><<<shadeGrid, SHADE_THREADS, 0, st>>>(sc, fr, P, shadePf); }
        if (th) th->end(1, st);
        if (th) th->begin(2, st);
        if (bin) launchOctBin<true>(P, st, launches);
        launchTraceShadow(traceGrid, sc, P, work + 1, st);
        if (th) th->end(2, st);
        k_advance<<<1, 32, 0, st>>>(P, work, 0);
        *launches += 5;
        // ping-pong the path queues
        uint32_t* tq = P.queueIn; P.queueIn = P.queueOut; P.queueOut = tq;
        uint32_t* tc = P.countIn; P.countIn = P.countOut; P.countOut = tc;
    }
    if (!(phases & 2)) return;
    if (th) th->begin(3, st);
    k_film<<<gridFor(P.numPixels, 128), 128, 0, st>>>(fr, P, F); (*launches)++;
    if (th) th->end(3, st);
}
// =============================================================================================
// Realtime profile: one frame = N-spp wavefront + per-pixel ReSTIR GI / EMA + a-trous (ohb_realtime.h)
// =============================================================================================
template <bool FUSED>
__global__ void __launch_bounds__(SHADE_THREADS, 8) k_bounce_rt(SceneDev sc, FrameParams fr, PathArrays P) {
    OHB_TILE_EMIT(q);
    const uint32_t n = *P.countIn;
    for (uint32_t tile = blockIdx.x * QTILE; tile < n; tile += gridDim.x * QTILE) {
        q.begin();
        for (int r = 0; r < QTILE_ROUNDS; r++) {
            const uint32_t i = tile + uint32_t(r) * SHADE_THREADS + threadIdx.x;
            uint32_t p = 0u; bool keep = false;
            if (i < n) { p = OHB_Q_PATH(P.queueSorted[i]); keep = bouncePathRT<FUSED>(sc, fr, P, p); }
            q.round(keep, p);
        }
        q.flush(P.countOut, P.queueOut, true, 0u);
    }
}
__global__ void __launch_bounds__(128) k_rt_pixel(SceneDev sc, FrameParams fr, PathArrays P, RTImagesDev im) {
    uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix < P.numPixels) pixelRT(sc, fr, P, im, pix);
}
// PASSES / FASTW: the shipped kernel is <1, true> (ohb_realtime.h denoiseRT); OHB_RT_DENOISE_VAR = 1 runs the shader's three
// passes literally with expf / powf weights (<3, false>), 2 = <3, true>, 3 = <1, false> — A/B only (profiles/r2ae)
template <int PASSES, bool FASTW>
__global__ void __launch_bounds__(128) k_rt_denoise(FrameParams fr, const f4* accum, const f4* normalAOV, uint32_t* ldr, float* dump, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) denoiseRT<PASSES, FASTW>(fr, accum, normalAOV, ldr, dump, i);
}
static int rtDenoiseVar() { static const int v = [] { const char* e = getenv("OHB_RT_DENOISE_VAR"); return e ? atoi(e) : 0; }(); return v; }
void launchRealtimeFrame(const SceneDev& sc, const FrameParams& fr, PathArrays P, const RTImagesDev& im, uint32_t* ldr, float* denoisedDump,
                         uint32_t* work, int numSMs, cudaStream_t st, uint64_t* launches, TimingHooks* th) {
    uint32_t total = P.numPixels * P.samplesInBatch;
    k_zero_u32<<<1, 32, 0, st>>>(P.countIn, 1); k_zero_u32<<<1, 32, 0, st>>>(P.countOut, 1);
    k_sobol_tab<<<gridFor(P.samplesInBatch, 64), 64, 0, st>>>(const_cast<u4*>(P.sobolTab), P.firstSampleIndex, P.samplesInBatch);
    k_raygen<true><<<gridFor(total, QTILE), SHADE_THREADS, 0, st>>>(fr, P, total);
    k_advance<<<1, 32, 0, st>>>(P, work, 1);
    *launches += 5;
    unsigned grid = unsigned(numSMs) * 8u;
    const bool bin = octBinOn() && !sc.twoLevel && P.octPerm != nullptr;
    if (!bin) P.octPerm = nullptr;
    static const bool rtFused = []() { const char* e = getenv("OHB_RT_FUSED"); return e ? atoi(e) != 0 : true; }();
    uint32_t iters = 2u + fr.maxBounces;          // primary + chain B + the ReSTIR GI bounce
    for (uint32_t it = 0; it < iters; it++) {
        if (th) th->begin(0, st);
        if (bin) launchOctBin<false>(P, st, launches);
        launchTraceClosest(grid, sc, P, work, st);
        if (th) th->end(0, st);
        if (th) th->begin(7, st);
        k_sort_hits<<<grid, SHADE_THREADS, 0, st>>>(P); (*launches)++;
        if (th) th->end(7, st);
        if (rtFused) {          // closest-hit / miss shaders + the realtime raygen body in one kernel, the payload in registers
            if (th) th->begin(1, st);
            k_bounce_rt<true><<<grid, SHADE_THREADS, 0, st>>>(sc, fr, P); (*launches)--;
            if (th) th->end(1, st);
        } else {
            if (th) th->begin(4, st);
            k_surface<<<grid, SHADE_THREADS, 0, st>>>(sc, fr, P);
            if (th) th->end(4, st);
            if (th) th->begin(1, st);
            k_bounce_rt<false><<<grid, SHADE_THREADS, 0, st>>>(sc, fr, P);
            if (th) th->end(1, st);
        }
        if (th) th->begin(2, st);
        if (bin) launchOctBin<true>(P, st, launches);
        launchTraceShadow(grid, sc, P, work + 1, st);
        if (th) th->end(2, st);
        k_advance<<<1, 32, 0, st>>>(P, work, 0);
        *launches += 5;
        uint32_t* tq = P.queueIn; P.queueIn = P.queueOut; P.queueOut = tq;
        uint32_t* tc = P.countIn; P.countIn = P.countOut; P.countOut = tc;
    }
    if (th) th->begin(5, st);
    k_rt_pixel<<<gridFor(P.numPixels, 128), 128, 0, st>>>(sc, fr, P, im);
    if (th) th->end(5, st);
    if (th) th->begin(3, st);
    {
        const unsigned g = gridFor(fr.W * fr.H, 128); const uint32_t n = fr.W * fr.H;
        switch (rtDenoiseVar()) {
        case 1: k_rt_denoise<3, false><<<g, 128, 0, st>>>(fr, im.accumCurr, P.normalAOV, ldr, denoisedDump, n); break;
        case 2: k_rt_denoise<3, true><<<g, 128, 0, st>>>(fr, im.accumCurr, P.normalAOV, ldr, denoisedDump, n); break;
        case 3: k_rt_denoise<1, false><<<g, 128, 0, st>>>(fr, im.accumCurr, P.normalAOV, ldr, denoisedDump, n); break;
        default: k_rt_denoise<1, true><<<g, 128, 0, st>>>(fr, im.accumCurr, P.normalAOV, ldr, denoisedDump, n); break;
        }
    }
    if (th) th->end(3, st);
    *launches += 2;
}

// =============================================================================================
// SVGF denoiser (DenoiseMode::Atrous): AtrousDenoiser::dispatch (atrous_denoise.cpp:392-537).  32x8 pixel blocks:
// a warp covers one row segment, so the 25 taps of neighbouring lanes share L1 lines.
// =============================================================================================
__global__ void __launch_bounds__(256) k_svgf_temporal(SvgfTemporalArgs a) {
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < a.W && y < a.H) svgfTemporalPixel(a, x, y);
}
__global__ void __launch_bounds__(256) k_svgf_atrous(SvgfAtrousArgs a) {
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < a.W && y < a.H) svgfAtrousPixel(a, x, y);
}
void launchSvgf(const SvgfBuffers& b, uint32_t W, uint32_t H, int cur, bool reset, cudaStream_t st, uint64_t* launches, TimingHooks* th) {
    const int prev = 1 - cur;
    dim3 grid((W + 31u) / 32u, (H + 7u) / 8u);
    if (th) th->begin(6, st);
    SvgfTemporalArgs t{};
    t.beauty = b.beauty; t.motion = b.motion; t.depth = b.depth; t.normal = b.normal;
    t.prevColor = b.histColor[prev]; t.prevMoments = b.histMoments[prev]; t.prevGeom = b.histGeom[prev];
    t.outColor = b.color[0]; t.outMoments = b.histMoments[cur]; t.outVariance = b.var[0]; t.outGeom = b.histGeom[cur];
    t.W = int(W); t.H = int(H); t.reset = reset ? 1 : 0;
    k_svgf_temporal<<<grid, 256, 0, st>>>(t);
    // ping-pong schedule of atrous_denoise.cpp:477-486: iteration 0 writes next frame's colour history
    h4* A = b.color[0]; h4* B = b.color[1]; h4* HC = b.histColor[cur];
    h4* inC[OHB_SVGF_ITERATIONS] = {A, HC, B, A, B}; h4* outC[OHB_SVGF_ITERATIONS] = {HC, B, A, B, A};
    for (int it = 0; it < OHB_SVGF_ITERATIONS; it++) {
        SvgfAtrousArgs a{};
        a.inColor = inC[it]; a.outColor16 = outC[it]; a.normal = b.normal; a.depth = b.depth; a.inVar = b.var[it & 1]; a.outVar = b.var[1 - (it & 1)]; a.outLDR = b.beauty;
        a.W = int(W); a.H = int(H); a.stepSize = 1 << it; a.isFinal = it == OHB_SVGF_ITERATIONS - 1;
        a.sigmaL = b.sigmaL; a.sigmaNormal = b.sigmaNormal; a.sigmaDepth = b.sigmaDepth;
        k_svgf_atrous<<<grid, 256, 0, st>>>(a);
    }
    if (th) th->end(6, st);
    *launches += 1 + OHB_SVGF_ITERATIONS;
}

void launchResolve(f4* accum, uint32_t* ldr, uint32_t n, int sumMode, cudaStream_t st, uint64_t* launches) {
    k_resolve<<<gridFor(n, 256), 256, 0, st>>>(accum, ldr, n, sumMode); (*launches)++;
}

// ---- parity hooks: trace caller-supplied rays ---------------------------------------------------
struct HookIO {
    const ohb_ray* rays; ohb_hit* hits; uint8_t* occ;
    __device__ __forceinline__ void load(uint32_t i, f3& o, f3& d, float& tmin, float& tmax) {
        ohb_ray r = rays[i]; o = mk3(r.origin[0], r.origin[1], r.origin[2]); d = mk3(r.dir[0], r.dir[1], r.dir[2]); tmin = r.tmin; tmax = r.tmax;
    }
    __device__ __forceinline__ void store(uint32_t i, Trav& t) {
        if (hits) hits[i] = travResult(t);
        else occ[i] = t.anyHit ? 1 : 0;
    }
};
template <bool TL>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_batch(SceneDev sc, const ohb_ray* rays, uint32_t n, ohb_hit* hits, uint32_t* work, int minActive, int postponeDen) {
    HookIO io{rays, hits, nullptr};
    persistentTrace<false, 0, HookIO, TL>(sc, n, work, io, minActive, postponeDen);
}
template <bool TL>
__global__ void __launch_bounds__(TRACE_THREADS) k_occluded_batch(SceneDev sc, const ohb_ray* rays, uint32_t n, uint8_t* occ, uint32_t* work, int minActive, int postponeDen) {
    HookIO io{rays, nullptr, occ};
    persistentTrace<true, 0, HookIO, TL>(sc, n, work, io, minActive, postponeDen);
}
void launchTraceBatch(const SceneDev& sc, const ohb_ray* rays, uint32_t n, ohb_hit* hits, uint8_t* occ, uint32_t* work, int numSMs, cudaStream_t st, uint64_t* launches) {
    k_zero_u32<<<1, 32, 0, st>>>(work, 2);
    const TraceKnobs k = traceKnobs();
    if (sc.twoLevel) {
        if (hits) k_trace_batch<true><<<numSMs * 6, TRACE_THREADS, 0, st>>>(sc, rays, n, hits, work, k.minActive, k.postponeDen & 0xFF);
        else      k_occluded_batch<true><<<numSMs * 6, TRACE_THREADS, 0, st>>>(sc, rays, n, occ, work, k.minActive, k.postponeDen & 0xFF);
    } else {
        if (hits) k_trace_batch<false><<<numSMs * 8, TRACE_THREADS, 0, st>>>(sc, rays, n, hits, work, k.minActive, k.postponeDen);
        else      k_occluded_batch<false><<<numSMs * 8, TRACE_THREADS, 0, st>>>(sc, rays, n, occ, work, k.minActive, k.postponeDen);
    }
    *launches += 2;
}

}  // namespace ohb

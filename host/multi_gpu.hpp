// host/multi_gpu.hpp — the sharded offline render of SURVEY §8e driven from the C++20 host: one thread per GPU of the box,
// scene replicated, GPU g renders its contiguous block of sample indices in SUM mode (ohb_set_accum_mode), ONE
// ncclReduce(sum) of the RGBA32F accumulation images over NVLink onto GPU 0 (in place, on the library's own device
// pointer, ohb_accum_dev_ptr), resolve + readback there.  The sample sequence is a pure function of (pixel, sample index)
// (sampler_sobol.glsl:60-82), so the sharded image holds exactly the samples of a one-GPU render; only the fp32 summation
// order differs.  Replaces nothing in the reference (its renderer drives one VkDevice); this is the multi-GPU seam north_star adds.
#pragma once
#include "ohao_b200_host.hpp"
#include <cuda_runtime_api.h>
#include <nccl.h>
#include <chrono>
#include <functional>
#include <thread>

namespace ohao {

struct ShardedResult { std::vector<uint8_t> pixels; double renderMs = 0, reduceMs = 0, totalMs = 0; bool ok = false; };

// buildScene(renderer) sets scene / camera / modes on a fresh Renderer bound to one GPU; it is called once per GPU.
inline ShardedResult renderSharded(int nGpus, uint32_t W, uint32_t H, uint32_t spp, uint32_t seed, const std::function<bool(Renderer&)>& buildScene) {
    ShardedResult out;
    int ndev = 0; cudaGetDeviceCount(&ndev);
    if (nGpus < 1 || nGpus > ndev) { std::cerr << "renderSharded: " << nGpus << " GPUs requested, " << ndev << " present\n"; return out; }
    const size_t n = size_t(nGpus);
    std::vector<int> devs(n); for (int g = 0; g < nGpus; g++) devs[size_t(g)] = g;
    std::vector<ncclComm_t> comms(n);
    if (nGpus > 1 && ncclCommInitAll(comms.data(), nGpus, devs.data()) != ncclSuccess) { std::cerr << "renderSharded: ncclCommInitAll failed\n"; return out; }
    std::vector<std::unique_ptr<Renderer>> rs(n);
    std::vector<cudaStream_t> streams(n);
    std::vector<int> fail(n, 0);
    // contiguous sample-index blocks whose sizes differ by at most one (sharding.py sample_blocks)
    std::vector<uint32_t> first(n), count(n);
    for (int g = 0, f = int(seed); g < nGpus; g++) { count[size_t(g)] = spp / uint32_t(nGpus) + (uint32_t(g) < spp % uint32_t(nGpus) ? 1u : 0u); first[size_t(g)] = uint32_t(f); f += int(count[size_t(g)]); }
    auto setup = [&](int g) {
        cudaSetDevice(g); cudaStreamCreateWithFlags(&streams[size_t(g)], cudaStreamNonBlocking);
        rs[size_t(g)] = std::make_unique<Renderer>(W, H, g);
        Renderer& r = *rs[size_t(g)];
        if (!r.initialize() || !buildScene(r) || !r.updateSceneBuffers()) { fail[size_t(g)] = 1; return; }
        ohb_ctx* c = r.rtRenderer()->ctx();
        if (ohb_set_accum_mode(c, 1) || ohb_clear_accum(c)) { fail[size_t(g)] = 1; return; }
        r.setRenderSeed(first[size_t(g)]);
        r.render(1); ohb_synchronize(c);                       // warm-up: module load, path-state allocation
        ohb_clear_accum(c); r.setRenderSeed(first[size_t(g)]);
        if (nGpus > 1) {                                       // ... and NCCL's lazy connection set-up (60-300 ms on the first collective): all-reduce of zeros
            size_t bytes = 0; float* acc = static_cast<float*>(ohb_accum_dev_ptr(c, &bytes)); ohb_synchronize(c);
            if (ncclAllReduce(acc, acc, 1024, ncclFloat32, ncclSum, comms[size_t(g)], streams[size_t(g)]) != ncclSuccess) { fail[size_t(g)] = 1; return; }
            cudaStreamSynchronize(streams[size_t(g)]);
        }
    };
    {
        std::vector<std::thread> th; for (int g = 0; g < nGpus; g++) th.emplace_back(setup, g);
        for (auto& t : th) t.join();
    }
    for (int f : fail) if (f) return out;
    auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<double> tRender(n, 0.0);
    auto work = [&](int g) {
        cudaSetDevice(g);
        Renderer& r = *rs[size_t(g)]; ohb_ctx* c = r.rtRenderer()->ctx();
        if (count[size_t(g)]) r.render(count[size_t(g)]);
        if (ohb_synchronize(c)) { fail[size_t(g)] = 1; return; }
        tRender[size_t(g)] = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
        if (nGpus > 1) {
            size_t bytes = 0; float* acc = static_cast<float*>(ohb_accum_dev_ptr(c, &bytes));
            if (ncclReduce(acc, acc, bytes / 4, ncclFloat32, ncclSum, 0, comms[size_t(g)], streams[size_t(g)]) != ncclSuccess) { fail[size_t(g)] = 1; return; }
            cudaStreamSynchronize(streams[size_t(g)]);
        }
    };
    {
        std::vector<std::thread> th; for (int g = 0; g < nGpus; g++) th.emplace_back(work, g);
        for (auto& t : th) t.join();
    }
    for (int f : fail) if (f) return out;
    double tr = 0; for (double t : tRender) tr = std::max(tr, t);
    out.renderMs = tr;
    out.reduceMs = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count() - tr;
    cudaSetDevice(0);
    ohb_ctx* c0 = rs[0]->rtRenderer()->ctx();
    if (ohb_resolve(c0)) return out;
    auto px = rs[0]->getPixelSpan();
    out.pixels.assign(px.begin(), px.end());
    out.totalMs = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    for (int g = 0; g < nGpus; g++) { cudaSetDevice(g); rs[size_t(g)].reset(); cudaStreamDestroy(streams[size_t(g)]); if (nGpus > 1) ncclCommDestroy(comms[size_t(g)]); }
    out.ok = !out.pixels.empty();
    return out;
}

}  // namespace ohao

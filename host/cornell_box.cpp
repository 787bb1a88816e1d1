// host/cornell_box.cpp — the reference's examples/cornell_box.cpp:57-170 on the B200 path tracer.
//   cornell_box [out.png] [spp] [--denoise=none] [--width W --height H] [--realtime] [--device N] [--seed S] [--gpus N]
// --gpus N (offline): the spp are split into N contiguous sample-index blocks, one GPU each, summed by ONE ncclReduce (multi_gpu.hpp).
// Same scene, camera and call sequence; differences: resolution is a flag (the reference hard-codes 1920x1080, quirk Q11)
// and `spp` render() calls are made instead of spp+3 (the Vulkan readback ring lags 3 frames, Q10; getPixelSpan here is
// current), so the image holds exactly `spp` samples in both.
#include "ohao_b200_host.hpp"
#ifdef OHB_HOST_NCCL
#include "multi_gpu.hpp"
#endif
#include <chrono>
using namespace ohao;

static void addWall(Scene* scene, std::string_view name, vec3 a, vec3 b, vec3 c, vec3 d, vec3 normal, vec3 color) {
    Actor* actor = scene->createActor(name);
    actor->model = std::make_shared<Model>();
    addQuad(*actor->model, a, b, c, d, normal, color);
    actor->material.baseColor = color; actor->material.roughness = 0.95f;
}

int main(int argc, char** argv) {
    const std::string output = argOr(argc, argv, 1, "cornell_box.png");
    const int samples = int(std::max(1L, std::atol(argOr(argc, argv, 2, "1024").c_str())));
    const uint32_t W = uint32_t(flagValue(argc, argv, "width", 1920)), H = uint32_t(flagValue(argc, argv, "height", 1080));
    const bool realtime = hasFlag(argc, argv, "realtime");
    std::cout << "OHAO Cornell Box — " << W << "x" << H << " @ " << samples << " spp\n";

    const int gpus = int(flagValue(argc, argv, "gpus", 1));
    const DenoiseMode denoise = denoiseFlag(argc, argv, realtime ? DenoiseMode::Atrous : DenoiseMode::None);   // Atrous = the SVGF denoiser (realtime profile)
    auto scene = std::make_unique<Scene>("Cornell Box");
    const float S = 5.0f;
    const vec3 white{0.73f, 0.73f, 0.73f}, red{0.65f, 0.05f, 0.05f}, green{0.12f, 0.45f, 0.15f};
    vec3 LBB{-S, -S, -S}, RBB{S, -S, -S}, LTB{-S, S, -S}, RTB{S, S, -S}, LBF{-S, -S, S}, RBF{S, -S, S}, LTF{-S, S, S}, RTF{S, S, S};
    addWall(scene.get(), "Back", LBB, RBB, RTB, LTB, {0, 0, 1}, white);
    addWall(scene.get(), "Left", LBB, LTB, LTF, LBF, {1, 0, 0}, red);
    addWall(scene.get(), "Right", RBB, RBF, RTF, RTB, {-1, 0, 0}, green);
    addWall(scene.get(), "Floor", LBB, LBF, RBF, RBB, {0, 1, 0}, white);
    addWall(scene.get(), "Ceiling", LTB, RTB, RTF, LTF, {0, -1, 0}, white);

    Actor* metal = scene->createActor("MetalSphere");
    metal->model = generateSphereMesh(); metal->position = {-2.0f, -S + 2.0f, 0.0f}; metal->scale = {2.0f, 2.0f, 2.0f};
    metal->material.baseColor = {0.95f, 0.93f, 0.88f}; metal->material.roughness = 0.05f; metal->material.metallic = 1.0f;
    Actor* glass = scene->createActor("GlassSphere");
    glass->model = generateSphereMesh(); glass->position = {2.5f, -S + 1.8f, 1.5f}; glass->scale = {1.8f, 1.8f, 1.8f};
    glass->material.baseColor = {0.9f, 0.95f, 1.0f}; glass->material.roughness = 0.02f;

    const vec3 lightColors[12] = {{1.0f, 0.3f, 0.2f}, {0.2f, 1.0f, 0.3f}, {0.3f, 0.4f, 1.0f}, {1.0f, 0.9f, 0.3f}, {1.0f, 0.5f, 0.0f}, {0.8f, 0.2f, 1.0f},
                                  {0.0f, 1.0f, 1.0f}, {1.0f, 0.0f, 0.5f}, {1.0f, 1.0f, 1.0f}, {0.5f, 1.0f, 0.5f}, {1.0f, 0.7f, 0.5f}, {0.5f, 0.7f, 1.0f}};
    const float lightPositions[12][3] = {{-3, 4, -3}, {0, 4, -3}, {3, 4, -3}, {-3, 4, 0}, {0, 4, 0}, {3, 4, 0}, {-3, 4, 3}, {0, 4, 3}, {3, 4, 3}, {-4, 0, 0}, {4, 0, 0}, {0, 0, -4}};
    for (int i = 0; i < 12; i++) {
        Actor* l = scene->createActor("Light" + std::to_string(i));
        l->hasLight = true; l->light.type = LightType::Sphere; l->light.color = lightColors[i]; l->light.intensity = 5.0f; l->light.radius = 0.3f;
        l->position = {lightPositions[i][0], lightPositions[i][1], lightPositions[i][2]};
    }
    // scene / camera / mode wiring shared by the one-GPU path and every rank of the sharded path (the Scene is read-only after this)
    auto wire = [&](Renderer& r) {
        r.setScene(scene.get());
        auto& camera = r.getCamera();
        camera.setPosition({0.0f, 0.0f, 13.0f}); camera.setFov(38.0f); camera.setRotation(0.0f, -90.0f);
        r.setRenderMode(realtime ? RenderMode::RTRealtime : RenderMode::RTOffline);
        r.setDenoiseMode(denoise);
        return true;
    };
    if (gpus > 1) {
#ifdef OHB_HOST_NCCL
        if (realtime) { std::cerr << "FATAL: --gpus shards the offline profile (the realtime profile's ReSTIR reuse reads neighbouring pixels of the previous frame)\n"; return 1; }
        ShardedResult sr = renderSharded(gpus, W, H, uint32_t(samples), uint32_t(flagValue(argc, argv, "seed", 0)), wire);
        if (!sr.ok) { std::cerr << "FATAL: sharded render failed\n"; return 1; }
        std::cout << "Done: " << long(sr.totalMs) << " ms on " << gpus << " GPUs (render " << sr.renderMs << " ms, ncclReduce + sync " << sr.reduceMs << " ms; "
                  << double(W) * H * samples / (sr.totalMs * 1e3) << " Msamples/s incl. reduce + readback)\n";
        if (!writePNG(output, sr.pixels.data(), W, H)) { std::cerr << "FATAL: cannot write " << output << "\n"; return 1; }
        std::cout << "Saved " << output << "\n";
        return 0;
#else
        std::cerr << "FATAL: built without NCCL (make NCCL=1)\n"; return 1;
#endif
    }
    Renderer renderer(W, H, int(flagValue(argc, argv, "device", 0)));
    if (!renderer.initialize()) { std::cerr << "FATAL: renderer init failed\n"; return 1; }
    wire(renderer);
    renderer.setRenderSeed(uint32_t(flagValue(argc, argv, "seed", 0)));
    if (!renderer.updateSceneBuffers()) { std::cerr << "FATAL: scene upload failed\n"; return 1; }

    std::cout << "Rendering (" << (realtime ? "RTRealtime" : "RTOffline") << ")...\n";
    auto start = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < samples; i++) renderer.render();
    const auto pixels = renderer.getPixelSpan();                // blocking readback, like vkDeviceWaitIdle + map
    const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - start).count();
    std::cout << "Done: " << ms << " ms  (" << double(W) * H * samples / (double(ms) * 1e3) << " Msamples/s incl. readback)\n";
    if (pixels.empty() || !writePNG(output, pixels.data(), W, H)) { std::cerr << "FATAL: no pixels / cannot write " << output << "\n"; return 1; }
    std::cout << "Saved " << output << "\n";
    return 0;
}

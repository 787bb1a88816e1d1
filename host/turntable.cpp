// host/turntable.cpp — the reference's examples/turntable.cpp:55-235 on the B200 path tracer.
//   turntable <model.glb|blob[:ntris]|sphere> <cornell|dark|env|mirror> [spp] [frames] [--realtime] [--hdr FILE] [--width W --height H] [--outdir DIR] [--device N]
// Same rooms, lights, orbit (radius / height / 15 deg tilt / fov 70) and per-frame call sequence (resetAccumulation, spp x render,
// getPixelSpan).  The hero mesh is a seeded procedural blob instead of a glTF file (asset loading is out of scope), the "env" mode
// uses the procedural outdoor HDRI.  Frames are written as PNG into --outdir (default renders/turntable) unless --no-write.
#include "procedural.hpp"
#include "asset_io.hpp"
#include <chrono>
#include <filesystem>
using namespace ohao;

static void addWall(Scene* scene, std::string_view name, vec3 a, vec3 b, vec3 c, vec3 d, vec3 n, vec3 color, float rough = 0.95f, float metal = 0.0f) {
    Actor* actor = scene->createActor(name);
    actor->model = std::make_shared<Model>();
    addQuad(*actor->model, a, b, c, d, n, color);
    actor->material.baseColor = color; actor->material.roughness = rough; actor->material.metallic = metal;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::cout << "Usage: turntable <model.glb|blob[:ntris]|sphere> <cornell|dark|env|mirror> [spp] [frames] [--realtime] [--hdr FILE] [--width W --height H]\n"; return 1; }
    const std::string modelSpec = argv[1], mode = argv[2];
    const int spp = int(std::max(1L, std::atol(argOr(argc, argv, 3, "64").c_str()))), totalFrames = int(std::max(1L, std::atol(argOr(argc, argv, 4, "120").c_str())));
    const bool realtime = hasFlag(argc, argv, "realtime");
    const uint32_t W = uint32_t(flagValue(argc, argv, "width", 1280)), H = uint32_t(flagValue(argc, argv, "height", 720));
    const std::string outdir = flagString(argc, argv, "outdir", "renders/turntable");
    std::cout << "OHAO Turntable — " << mode << " mode, " << spp << " spp, " << totalFrames << " frames, " << (realtime ? "RTRealtime" : "RTOffline") << "\n";

    Renderer renderer(W, H, int(flagValue(argc, argv, "device", 0)));
    if (!renderer.initialize()) return 1;
    auto scene = std::make_unique<Scene>("Turntable");
    const float S = 5.0f;
    const vec3 white{0.73f, 0.73f, 0.73f}, red{0.65f, 0.05f, 0.05f}, green{0.12f, 0.45f, 0.15f}, mirror{0.95f, 0.95f, 0.95f};
    vec3 LBB{-S, -S, -S}, RBB{S, -S, -S}, LTB{-S, S, -S}, RTB{S, S, -S}, LBF{-S, -S, S}, RBF{S, -S, S}, LTF{-S, S, S}, RTF{S, S, S};
    if (mode == "env") {
        // the reference loads assets/test_models/env_outdoor.hdr (turntable.cpp:79); --hdr names the file, without it a procedural sky stands in
        const std::string hdr = flagString(argc, argv, "hdr", "");
        std::vector<float> env; uint32_t ew = 0, eh = 0;
        if (!hdr.empty()) { if (!loadHDR(hdr, env, ew, eh)) return 1; renderer.setEnvironmentMap(std::move(env), ew, eh); }
        else renderer.setEnvironmentMap(proceduralEnv(1024, 512), 1024, 512);
    } else {
        addWall(scene.get(), "Left", LBB, LTB, LTF, LBF, {1, 0, 0}, red); addWall(scene.get(), "Right", RBB, RBF, RTF, RTB, {-1, 0, 0}, green);
        addWall(scene.get(), "Floor", LBB, LBF, RBF, RBB, {0, 1, 0}, white); addWall(scene.get(), "Ceiling", LTB, RTB, RTF, LTF, {0, -1, 0}, white);
        if (mode == "mirror") { addWall(scene.get(), "BackMirror", LBB, RBB, RTB, LTB, {0, 0, 1}, mirror, 0.02f, 1.0f); addWall(scene.get(), "FrontMirror", LBF, RBF, RTF, LTF, {0, 0, -1}, mirror, 0.02f, 1.0f); }
        else addWall(scene.get(), "Back", LBB, RBB, RTB, LTB, {0, 0, 1}, white);
    }
    // hero: centred, resting on the floor like the reference's auto-placed model
    uint32_t ntris = 50000; if (auto c = modelSpec.find(':'); c != std::string::npos) ntris = uint32_t(std::atol(modelSpec.c_str() + c + 1));
    Actor* hero = scene->createActor("Hero");
    const bool isGlb = modelSpec.size() > 4 && (modelSpec.substr(modelSpec.size() - 4) == ".glb" || modelSpec.substr(modelSpec.size() - 4) == ".GLB");
    if (isGlb) {
        // turntable.cpp:111-156: load, scale to a height of 4 (5.5 in the mirror room), rotate 180 degrees about Y, centre, feet on the floor
        hero->model = std::make_shared<Model>();
        if (!loadGLB(modelSpec, *hero->model, hero->material)) return 1;
        vec3 bmin{FLT_MAX, FLT_MAX, FLT_MAX}, bmax{-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (const Vertex& v : hero->model->vertices) {
            bmin = {std::min(bmin.x, v.position[0]), std::min(bmin.y, v.position[1]), std::min(bmin.z, v.position[2])};
            bmax = {std::max(bmax.x, v.position[0]), std::max(bmax.y, v.position[1]), std::max(bmax.z, v.position[2])};
        }
        if (bmax.y - bmin.y < bmax.z - bmin.z) { std::cerr << "turntable: Z-up models are not supported\n"; return 1; }
        const float scale = (mode == "mirror" ? 5.5f : 4.0f) / std::max(bmax.y - bmin.y, bmax.z - bmin.z), floorY = mode == "env" ? 0.0f : -5.0f;
        const vec3 c = (bmin + bmax) * 0.5f;
        hero->yaw180 = true; hero->scale = {scale, scale, scale}; hero->position = {-c.x * scale, floorY - bmin.y * scale, -c.z * scale};
    } else {
        hero->model = modelSpec.rfind("sphere", 0) == 0 ? generateSphereMesh(64, 32, 1.0f) : generateBlobMesh(ntris, 1);
        hero->scale = {1.5f, 1.5f, 1.5f}; hero->position = {0.0f, mode == "env" ? 0.0f : -S + 1.8f, 0.0f};
        hero->material.baseColor = {1, 1, 1}; hero->material.roughness = 1.0f; hero->material.metallic = 1.0f;
        hero->material.albedoTex = proceduralTexture(1024, 11, 0); hero->material.normalTex = proceduralTexture(1024, 12, 1);
        hero->material.roughMetalTex = proceduralTexture(1024, 13, 2); hero->material.emissiveTex = proceduralTexture(1024, 14, 3);
    }
    auto sphereLight = [&](const char* name, vec3 pos, vec3 color, float intensity, float radius) {
        Actor* l = scene->createActor(name); l->hasLight = true; l->light.color = color; l->light.intensity = intensity; l->light.radius = radius; l->position = pos;
    };
    if (mode == "dark") sphereLight("Fill", {4, 4, 4}, {1, 1, 1}, 2.0f, 0.3f);
    else {
        sphereLight("Key", {3, 4, 3}, {1, 0.95f, 0.9f}, mode == "env" ? 8.0f : (mode == "mirror" ? 20.0f : 25.0f), 1.0f);
        if (mode == "cornell") {
            Actor* a = scene->createActor("CeilingPanel"); a->hasLight = true; a->light.type = LightType::AreaRect; a->light.color = {1, 0.98f, 0.92f}; a->light.intensity = 12.0f;
            a->light.edge1 = {3, 0, 0}; a->light.edge2 = {0, 0, 3}; a->position = {-1.5f, 4.99f, -1.5f};
        }
    }
    renderer.setScene(scene.get());
    renderer.setRenderMode(realtime ? RenderMode::RTRealtime : RenderMode::RTOffline);
    renderer.setDenoiseMode(denoiseFlag(argc, argv, realtime ? DenoiseMode::Atrous : DenoiseMode::None));   // Atrous = the SVGF denoiser (realtime profile)
    if (const std::string dump = flagString(argc, argv, "dump-scene", ""); !dump.empty()) {
        // no GPU needed: write the packed arrays (the §3.2 data contract) for the tests to compare with the Python packer
        return dumpPackedScene(packScene(*scene, mode == "env", 1.0f).arrays, dump) ? 0 : 1;
    }
    if (!renderer.updateSceneBuffers()) { std::cerr << "FATAL: scene upload failed\n"; return 1; }
    const float orbitRadius = mode == "env" ? 8.0f : (mode == "mirror" ? 4.5f : 4.2f), orbitHeight = mode == "env" ? 1.0f : -3.5f;
    const bool write = !hasFlag(argc, argv, "no-write");
    // --two-level: BLAS per actor + TLAS instead of the flattened tree; --bob: the hero also bobs up and down every frame and the
    // structure is refit (buildTLAS in MODE_UPDATE) instead of rebuilt
    const bool bob = hasFlag(argc, argv, "bob"); const float heroY0 = hero->position.y;
    if (hasFlag(argc, argv, "two-level")) { renderer.setTwoLevelAccel(true); if (!renderer.updateSceneBuffers()) { std::cerr << "FATAL: scene upload failed\n"; return 1; } }
    if (write) std::filesystem::create_directories(outdir);
    auto t0 = std::chrono::high_resolution_clock::now();
    for (int frame = 0; frame < totalFrames; frame++) {
        float t = float(frame) / float(totalFrames), angle = t * 2.0f * 3.14159f;
        float cx = orbitRadius * std::cos(angle), cz = orbitRadius * std::sin(angle);
        auto& camera = renderer.getCamera();
        camera.setPosition({cx, orbitHeight, cz});
        camera.setRotation(15.0f, std::atan2(-cz, -cx) * 57.29577951308232f);
        camera.setFov(70.0f);
        if (bob) { hero->position.y = heroY0 + 0.3f * std::sin(angle * 4.0f); if (!renderer.updateInstanceTransforms()) { std::cerr << "FATAL: TLAS update failed\n"; return 1; } }
        if (realtime) { renderer.notifyCameraChanged(); renderer.render(); }            // realtime: one frame per orbit step, history carried over
        else { renderer.resetAccumulation(); renderer.render(uint32_t(spp)); }            // offline: spp render() calls, batched into one submission
        const auto pixels = renderer.getPixelSpan();
        if (pixels.empty()) { std::cerr << "FATAL: no pixels\n"; return 1; }
        if (write) { char fn[512]; std::snprintf(fn, sizeof(fn), "%s/%s_%04d.png", outdir.c_str(), mode.c_str(), frame); writePNG(fn, pixels.data(), W, H); }
    }
    double s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    std::cout << "Done: " << totalFrames << " frames in " << s << " s (" << totalFrames / s << " fps, " << double(W) * H * (realtime ? 1 : spp) * totalFrames / s / 1e6 << " Msamples/s)\n";
    return 0;
}

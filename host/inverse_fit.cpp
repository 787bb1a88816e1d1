// host/inverse_fit.cpp — `inverse_fit --backend pt --preset lantern --quality draft` (examples/inverse_fit.cpp:26,
// ohao/inverse/{fit_engine,staged_fit,param_space,optimizer,render_session,quality}.hpp) on the B200 path tracer.
//
// What is kept from the reference: the studio scene (hero + ground + pedestal + key / fill / rim sphere lights + env,
// scene_builder.hpp), theta in R^12 = albedo RGB, roughness, metallic | pedestal RGB | key, fill, rim intensity / 40 | env
// scale (scene_builder.hpp:160-200), central finite differences with projection (param_space.hpp:54-74), Adam
// (optimizer.hpp:12-40, lr 0.06, eps 0.04 = fit_config.hpp:32-34), the FIT budget of the draft preset (384x216 @ 32 spp, 3
// views, quality.hpp:23-24), RenderSession::render's call order per probe (seed + view*9973, material / light edits without a
// BLAS rebuild, resetAccumulation, spp renders, RGBA8 read-back, render_session.hpp:29-62) and the MSE-on-LDR8 loss.
// Every probe uses the same seed (common random numbers), so FD noise cancels regardless of which GPU renders it.
// What is new: the 2*12*3 = 72 probe renders of an iteration are independent and are distributed round-robin over
// `--gpus G` devices (one context + one host thread per GPU); only one double per probe comes back.
// Not restated: the staged multi-start schedule, priors and the "show" render / OIDN (out of the hot path).
//   inverse_fit --backend pt --preset lantern --quality draft [--iters N] [--gpus G] [--seed S] [--tris N]
#include "procedural.hpp"
#include <chrono>
#include <thread>
using namespace ohao;

struct Budget { uint32_t w, h; int spp; };
struct InverseScene {                      // scene_builder.hpp: buildStudio
    std::unique_ptr<Scene> scene; Actor *hero, *ground, *pedestal, *key, *fill, *rim; float envScale = 1.0f;
    static constexpr float kKeyIScale = 40.0f;
    void applyTheta(const std::vector<double>& t) {
        hero->material.baseColor = {float(t[0]), float(t[1]), float(t[2])}; hero->material.roughness = float(t[3]); hero->material.metallic = float(t[4]);
        pedestal->material.baseColor = {float(t[5]), float(t[6]), float(t[7])};
        key->light.intensity = float(t[8]) * kKeyIScale; fill->light.intensity = float(t[9]) * kKeyIScale; rim->light.intensity = float(t[10]) * kKeyIScale;
        envScale = float(t[11]);
    }
    void applyCamera(Camera& c, int view) const {          // 3 fixed views around the hero
        const float yaw[3] = {-90.0f, -55.0f, -125.0f}; const vec3 pos[3] = {{0, 1.4f, 4.2f}, {-2.4f, 1.6f, 3.4f}, {2.4f, 1.2f, 3.4f}};
        c.setPosition(pos[view]); c.setRotation(-8.0f, yaw[view]); c.setFov(40.0f);
    }
};
static InverseScene buildStudio(uint32_t heroTris) {
    InverseScene s; s.scene = std::make_unique<Scene>("InverseStudio");
    s.ground = s.scene->createActor("Ground"); s.ground->model = std::make_shared<Model>();
    addQuad(*s.ground->model, {-8, 0, -8}, {-8, 0, 8}, {8, 0, 8}, {8, 0, -8}, {0, 1, 0}, {0.5f, 0.5f, 0.5f});
    s.ground->material.baseColor = {0.45f, 0.45f, 0.47f}; s.ground->material.roughness = 0.8f;
    s.pedestal = s.scene->createActor("Pedestal"); s.pedestal->model = std::make_shared<Model>();
    addBox(*s.pedestal->model, {-0.7f, 0.0f, -0.7f}, {0.7f, 0.5f, 0.7f}, {0.6f, 0.6f, 0.6f}); s.pedestal->material.roughness = 0.6f;
    s.hero = s.scene->createActor("Hero"); s.hero->model = generateBlobMesh(heroTris, 7, 0.75f); s.hero->position = {0, 1.3f, 0};
    auto light = [&](const char* n, vec3 p, vec3 c, float r) { Actor* a = s.scene->createActor(n); a->hasLight = true; a->light.color = c; a->light.radius = r; a->position = p; return a; };
    s.key = light("Key", {3.0f, 4.0f, 3.0f}, {1.0f, 0.95f, 0.9f}, 0.8f); s.fill = light("Fill", {-3.5f, 2.5f, 2.0f}, {0.85f, 0.9f, 1.0f}, 1.0f); s.rim = light("Rim", {0.0f, 3.5f, -3.5f}, {1, 1, 1}, 0.6f);
    return s;
}
struct ParamSpace {                       // param_space.hpp
    std::vector<double> values, lo, hi;
    size_t size() const { return values.size(); }
    double project(size_t i, double v) const { return std::min(std::max(v, lo[i]), hi[i]); }
};
struct AdamState {                        // optimizer.hpp:12-40
    std::vector<double> m, v; int t = 0; double beta1 = 0.9, beta2 = 0.999, eps = 1e-8;
    void step(ParamSpace& sp, const std::vector<double>& g, double lr) {
        if (m.size() != sp.size()) { m.assign(sp.size(), 0.0); v.assign(sp.size(), 0.0); t = 0; }
        ++t; const double b1t = 1.0 - std::pow(beta1, t), b2t = 1.0 - std::pow(beta2, t);
        for (size_t i = 0; i < sp.size(); i++) {
            m[i] = beta1 * m[i] + (1.0 - beta1) * g[i]; v[i] = beta2 * v[i] + (1.0 - beta2) * g[i] * g[i];
            sp.values[i] = sp.project(i, sp.values[i] - lr * (m[i] / b1t) / (std::sqrt(v[i] / b2t) + eps));
        }
    }
};
static double mseRGB(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {     // image_loss.hpp:59: mean squared error over RGB in [0,1]
    double s = 0.0; size_t n = a.size() / 4;
    for (size_t i = 0; i < n; i++) for (int c = 0; c < 3; c++) { double d = (double(a[i * 4 + c]) - double(b[i * 4 + c])) / 255.0; s += d * d; }
    return s / double(n * 3);
}
// One GPU worker = one Renderer (context) + its own copy of the scene description.
struct Worker {
    InverseScene inv; std::unique_ptr<Renderer> renderer; Budget fit; bool bound = false;
    bool init(int device, uint32_t heroTris, Budget b, const std::vector<float>& env, uint32_t ew, uint32_t eh) {
        inv = buildStudio(heroTris); fit = b;
        renderer = std::make_unique<Renderer>(b.w, b.h, device);
        if (!renderer->initialize()) return false;
        renderer->setRenderMode(RenderMode::RTOffline); renderer->setDenoiseMode(DenoiseMode::None);
        renderer->setEnvironmentMap(env, ew, eh);
        return true;
    }
    bool render(const std::vector<double>& theta, int view, uint32_t seed, std::vector<uint8_t>& out) {     // RenderSession::render
        inv.applyTheta(theta);
        inv.applyCamera(renderer->getCamera(), view);
        renderer->setRenderSeed(seed + uint32_t(view) * 9973u);
        renderer->setEnvIntensityScale(inv.envScale);
        if (!bound) { renderer->setScene(inv.scene.get()); if (!renderer->updateSceneBuffers()) return false; bound = true; }
        else if (!renderer->updateRTMaterialParams() || !renderer->updateRTLightParams()) return false;
        renderer->resetAccumulation();
        renderer->render(uint32_t(fit.spp));
        auto px = renderer->getPixelSpan();
        if (px.empty()) return false;
        out.assign(px.begin(), px.end());
        return true;
    }
};

int main(int argc, char** argv) {
    const std::string backend = flagString(argc, argv, "backend", "pt"), preset = flagString(argc, argv, "preset", "lantern"), quality = flagString(argc, argv, "quality", "draft");
    if (backend != "pt") { std::cerr << "only --backend pt exists on this path\n"; return 1; }
    const Budget fit = quality == "draft" ? Budget{384, 216, 32} : (quality == "ultra" || quality == "cinema") ? Budget{960, 540, 256} : Budget{640, 360, 128};   // quality.hpp:23-30
    const int iters = int(flagValue(argc, argv, "iters", 40)), G = int(std::max(1L, flagValue(argc, argv, "gpus", 1)));
    const uint32_t seed = uint32_t(flagValue(argc, argv, "seed", 1234)), heroTris = uint32_t(flagValue(argc, argv, "tris", 15452));   // the in-tree helmet's triangle count
    const double lr = 0.06, eps = 0.04; const int nViews = 3;
    std::cout << "OHAO inverse_fit — backend pt, preset " << preset << ", quality " << quality << " (FIT " << fit.w << "x" << fit.h << " @ " << fit.spp << " spp, " << nViews << " views), " << G << " GPU(s)\n";
    std::vector<float> env = proceduralEnv(512, 256, {0.2f, 0.8f, 0.4f}, 600.0f);
    std::vector<Worker> workers(static_cast<size_t>(G));
    for (int g = 0; g < G; g++) if (!workers[size_t(g)].init(g, heroTris, fit, env, 512, 256)) { std::cerr << "FATAL: worker " << g << " init failed\n"; return 1; }

    const std::vector<double> truth = {0.80, 0.45, 0.25, 0.35, 0.10, 0.55, 0.50, 0.45, 12.0 / 40.0, 5.0 / 40.0, 8.0 / 40.0, 1.0};
    ParamSpace sp; sp.values = {0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.2, 0.2, 0.2, 0.7};
    sp.lo = {0.02, 0.02, 0.02, 0.05, 0.0, 0.02, 0.02, 0.02, 0.0, 0.0, 0.0, 0.0}; sp.hi = {1, 1, 1, 1, 1, 1, 1, 1, 1.5, 1.5, 1.5, 3.0};
    // Bind every worker's scene at the SAME theta (the initial guess).  updateRTMaterialParams only rewrites the 3 vec4 per
    // material (rt_build.cpp:188-389); the 1x1 solid-colour diffuse layer each untextured material received at upload time
    // (quirk Q1) keeps the colour it was bound with, so the image depends on the theta at bind time — with one renderer the
    // reference is self-consistent, with G renderers they must all bind alike or their probes disagree.
    { std::vector<uint8_t> scratch; for (int g = 0; g < G; g++) if (!workers[size_t(g)].render(sp.values, 0, seed, scratch)) { std::cerr << "FATAL: bind render failed on GPU " << g << "\n"; return 1; } }
    // targets: truth rendered with the same budget and seeds (worker 0)
    std::vector<std::vector<uint8_t>> target(static_cast<size_t>(nViews));
    for (int v = 0; v < nViews; v++) if (!workers[0].render(truth, v, seed, target[size_t(v)])) { std::cerr << "FATAL: target render failed\n"; return 1; }

    struct Job { std::vector<double> theta; int view; double loss; };
    auto runJobs = [&](std::vector<Job>& jobs) -> bool {
        std::vector<std::thread> th; std::vector<int> ok(static_cast<size_t>(G), 1);
        for (int g = 0; g < G; g++) th.emplace_back([&, g] {
            std::vector<uint8_t> img;
            for (size_t j = size_t(g); j < jobs.size(); j += size_t(G)) {          // round-robin == sharding.jobs_for_rank
                if (!workers[size_t(g)].render(jobs[j].theta, jobs[j].view, seed, img)) { ok[size_t(g)] = 0; return; }
                jobs[j].loss = mseRGB(img, target[size_t(jobs[j].view)]);
            }
        });
        for (auto& t : th) t.join();
        return std::all_of(ok.begin(), ok.end(), [](int x) { return x != 0; });
    };
    auto lossAt = [&](const std::vector<double>& theta, double& out) {
        std::vector<Job> jobs; for (int v = 0; v < nViews; v++) jobs.push_back({theta, v, 0.0});
        if (!runJobs(jobs)) return false;
        out = 0.0; for (auto& j : jobs) out += j.loss / nViews;
        return true;
    };
    double L0 = 0.0; if (!lossAt(sp.values, L0)) return 1;
    std::cout << "iter 0  loss " << L0 << "\n";
    AdamState adam; size_t probes = 0; double probeSeconds = 0.0;
    for (int it = 1; it <= iters; it++) {
        // finiteDiffGradient: 2 * |theta| * nViews independent probes
        std::vector<Job> jobs; std::vector<double> denom(sp.size(), 0.0);
        for (size_t i = 0; i < sp.size(); i++) {
            double v0 = sp.values[i], hi = sp.project(i, v0 + eps), lo = sp.project(i, v0 - eps);
            denom[i] = hi - lo;
            for (int side = 0; side < 2; side++) for (int v = 0; v < nViews; v++) { Job j{sp.values, v, 0.0}; j.theta[i] = side ? lo : hi; jobs.push_back(std::move(j)); }
        }
        auto t0 = std::chrono::high_resolution_clock::now();
        if (!runJobs(jobs)) { std::cerr << "FATAL: probe render failed\n"; return 1; }
        probeSeconds += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count(); probes += jobs.size();
        std::vector<double> g(sp.size(), 0.0);
        for (size_t i = 0; i < sp.size(); i++) {
            if (denom[i] < 1e-12) continue;
            double Lh = 0, Ll = 0;
            for (int v = 0; v < nViews; v++) { Lh += jobs[(i * 2 + 0) * size_t(nViews) + size_t(v)].loss / nViews; Ll += jobs[(i * 2 + 1) * size_t(nViews) + size_t(v)].loss / nViews; }
            g[i] = (Lh - Ll) / denom[i];
        }
        adam.step(sp, g, lr);
        if (it % 5 == 0 || it == iters) { double L; if (!lossAt(sp.values, L)) return 1; std::cout << "iter " << it << "  loss " << L << "\n"; }
    }
    double Lf = 0.0; lossAt(sp.values, Lf);
    double err = 0.0; for (size_t i = 0; i < sp.size(); i++) err += (sp.values[i] - truth[i]) * (sp.values[i] - truth[i]);
    std::cout << "final loss " << Lf << " (initial " << L0 << "), |theta - truth| = " << std::sqrt(err) << "\n";
    std::cout << "probes: " << probes << " in " << probeSeconds << " s = " << double(probes) / probeSeconds << " probes/s, " << probeSeconds / iters << " s per FD iteration, "
              << double(probes) * fit.w * fit.h * fit.spp / probeSeconds / 1e6 << " Msamples/s on " << G << " GPU(s)\n";
    std::cout << "theta:"; for (double v : sp.values) std::cout << " " << v; std::cout << "\n";
    return Lf < L0 ? 0 : 2;
}

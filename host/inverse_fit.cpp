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
// --schedule staged (default) restates StagedFitter (staged_fit.hpp:150-780) for the synthetic-target product presets: the
// multi-start probe (:398-520), the stage list env -> lights -> albedo -> brdf -> brdf2 -> pedestal -> lights2 -> refine
// (:700-729) with per-stage Adam, lr / eps / spp multipliers and early stops (runStage, :296-395), the diffuse-target metal
// lock (:765-778), and the FIT loss (lossAt, :206-294): hybridSpecularRGB (image_loss.hpp:123-171) over the cropped frame
// (maskYMin 0.22), views weighted 1 / 0.5 / 0.5, + the multi-light regulariser (:160-201) + the preset's metal / rough prior.
// --schedule flat is the plain all-parameter FD + Adam loop (72 probes per iteration) the config-5 probes/s figure is quoted on.
// Not restated: the neural theta prior (C1), external photo / lab-bundle targets, the visual polish and the "show" render / OIDN.
//   inverse_fit --backend pt --preset lantern --quality draft [--iters N] [--gpus G] [--seed S] [--tris N] [--schedule staged|flat]
#include "procedural.hpp"
#include <chrono>
#include <limits>
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
// image_loss.hpp:37-46, :123-171 — highlight-weighted MSE + 0.35 MAE over the cropped frame
static double hybridSpecularRGB(const std::vector<uint8_t>& pred, const std::vector<uint8_t>& target, uint32_t W, uint32_t H, double xMaxFrac, double yMinFrac, double maeWeight, double specularWeight) {
    if (pred.size() != target.size() || pred.empty()) return std::numeric_limits<double>::infinity();
    const double sw = std::min(std::max(specularWeight, 0.0), 1.0);
    uint32_t xLim = xMaxFrac >= 1.0 ? W : uint32_t(std::ceil(xMaxFrac * double(W))), y0 = yMinFrac <= 0.0 ? 0u : uint32_t(std::floor(yMinFrac * double(H)));
    xLim = std::min(xLim, W); y0 = std::min(y0, H);
    double sumMse = 0.0, sumMae = 0.0, sumW = 0.0; size_t count = 0;
    for (uint32_t y = y0; y < H; ++y) for (uint32_t x = 0; x < xLim; ++x) {
        const size_t o = (size_t(y) * W + x) * 4;
        const double tr = target[o] / 255.0, tg = target[o + 1] / 255.0, tb = target[o + 2] / 255.0, pr = pred[o] / 255.0, pg = pred[o + 1] / 255.0, pb = pred[o + 2] / 255.0;
        const double dr = pr - tr, dg = pg - tg, db = pb - tb;
        const double mse = dr * dr + dg * dg + db * db, mae = std::abs(dr) + std::abs(dg) + std::abs(db);
        const double lumaT = 0.2126 * tr + 0.7152 * tg + 0.0722 * tb, maxT = std::max({tr, tg, tb}), maxP = std::max({pr, pg, pb});
        const double highlight = std::max(maxT, 0.65 * maxP);
        const double wSpec = 0.18 + std::pow(std::max(highlight, lumaT), 1.35);
        const double w = sw <= 1e-9 ? 1.0 : (1.0 - sw) + sw * wSpec;
        sumMse += w * mse; sumMae += w * mae; sumW += w; ++count;
    }
    if (count == 0 || sumW <= 0.0) return 0.0;
    const double inv = 1.0 / (sumW * 3.0);
    return sumMse * inv + maeWeight * sumMae * inv;
}
// io.hpp:23-57 — how specular the target looks (bright fraction, peak contrast, mean max channel)
static double targetHighlightScore(const std::vector<uint8_t>& img, uint32_t W, uint32_t H, double xMaxFrac, double yMinFrac) {
    if (img.empty()) return 0.0;
    const uint32_t xLim = xMaxFrac >= 1.0 ? W : uint32_t(std::ceil(xMaxFrac * W)), y0 = yMinFrac <= 0.0 ? 0u : uint32_t(std::floor(yMinFrac * H));
    size_t bright = 0, count = 0; double sumL = 0.0, maxL = 0.0, sumMax = 0.0;
    for (uint32_t y = y0; y < H; ++y) for (uint32_t x = 0; x < xLim; ++x) {
        const size_t o = (size_t(y) * W + x) * 4;
        const double r = img[o] / 255.0, g = img[o + 1] / 255.0, b = img[o + 2] / 255.0, mx = std::max({r, g, b}), luma = 0.2126 * r + 0.7152 * g + 0.0722 * b;
        if (mx > 0.55 || luma > 0.50) ++bright;
        sumL += luma; sumMax += mx; maxL = std::max(maxL, luma); ++count;
    }
    if (count == 0) return 0.0;
    const double frac = double(bright) / double(count), meanL = sumL / double(count), meanMax = sumMax / double(count), contrast = maxL / (meanL + 1e-3);
    const double cScore = std::min(std::max((contrast - 1.4) / 2.5, 0.0), 1.0), mScore = std::min(std::max((meanMax - 0.25) / 0.55, 0.0), 1.0);
    return std::min(std::max(0.40 * frac + 0.35 * cScore + 0.25 * mScore, 0.0), 1.0);
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
    bool render(const std::vector<double>& theta, int view, uint32_t seed, std::vector<uint8_t>& out, int spp = 0) {     // RenderSession::render
        inv.applyTheta(theta);
        inv.applyCamera(renderer->getCamera(), view);
        renderer->setRenderSeed(seed + uint32_t(view) * 9973u);
        renderer->setEnvIntensityScale(inv.envScale);
        if (!bound) { renderer->setScene(inv.scene.get()); if (!renderer->updateSceneBuffers()) return false; bound = true; }
        else if (!renderer->updateRTMaterialParams() || !renderer->updateRTLightParams()) return false;
        renderer->resetAccumulation();
        renderer->render(uint32_t(spp > 0 ? spp : fit.spp));
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

    const bool staged = flagString(argc, argv, "schedule", "staged") == "staged";
    const double maskX = 1.0, maskYMin = 0.22, specularWeight = 0.45, lightReg = 0.035; const float brdfSppMul = 2.0f;      // fit_config.hpp:35-36, 72-75
    size_t probes = 0; double probeSeconds = 0.0; uint64_t probeSamples = 0;
    struct Job { std::vector<double> theta; int view; double loss; int spp = 0; double sw = -1.0; };     // sw < 0: plain MSE (flat schedule)
    auto runJobs = [&](std::vector<Job>& jobs) -> bool {
        auto t0 = std::chrono::high_resolution_clock::now();
        std::vector<std::thread> th; std::vector<int> ok(static_cast<size_t>(G), 1);
        for (int g = 0; g < G; g++) th.emplace_back([&, g] {
            std::vector<uint8_t> img;
            for (size_t j = size_t(g); j < jobs.size(); j += size_t(G)) {          // round-robin == sharding.jobs_for_rank
                if (!workers[size_t(g)].render(jobs[j].theta, jobs[j].view, seed, img, jobs[j].spp)) { ok[size_t(g)] = 0; return; }
                jobs[j].loss = jobs[j].sw < 0.0 ? mseRGB(img, target[size_t(jobs[j].view)])
                                                : hybridSpecularRGB(img, target[size_t(jobs[j].view)], fit.w, fit.h, maskX, maskYMin, 0.35, jobs[j].sw);
            }
        });
        for (auto& t : th) t.join();
        probeSeconds += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count(); probes += jobs.size();
        for (auto& j : jobs) probeSamples += uint64_t(fit.w) * fit.h * uint64_t(j.spp > 0 ? j.spp : fit.spp);
        return std::all_of(ok.begin(), ok.end(), [](int x) { return x != 0; });
    };
    if (staged) {
        // ---- StagedFitter (staged_fit.hpp) for a product preset with a synthetic target -----------------------------------
        const size_t roughIdx = 3, metalIdx = 4;
        const double highlightScore = targetHighlightScore(target[0], fit.w, fit.h, maskX, maskYMin);
        auto lightRegularizer = [&](const std::vector<double>& th) {                                   // :160-201
            const double key = th[8], fill = th[9], rim = th[10], envS = th[11];
            double reg = (key - 0.50) * (key - 0.50) + 0.6 * (fill - 0.22) * (fill - 0.22) + 0.6 * (rim - 0.25) * (rim - 0.25) + 0.4 * (envS - 0.95) * (envS - 0.95);
            if (fill > key) reg += 2.0 * (fill - key) * (fill - key);
            if (rim > key) reg += 2.0 * (rim - key) * (rim - key);
            const double total = key + fill + rim, targetTotal = 0.50 + 0.22 + 0.25;
            reg += 0.35 * (total - targetTotal) * (total - targetTotal);
            return lightReg * reg;
        };
        auto metalPrior = [&](const std::vector<double>& th) {                                         // :268-291, diffuse (product) target
            const double metal = th[metalIdx], rough = th[roughIdx];
            double pr = 0.12 * (metal - 0.10) * (metal - 0.10);
            if (metal > 0.35) pr += 0.25 * (metal - 0.35) * (metal - 0.35);
            return pr + 0.03 * (rough - 0.40) * (rough - 0.40);
        };
        // the image part of lossAt for many thetas at once: every (theta, view) render is one independent job
        auto imageLosses = [&](const std::vector<std::vector<double>>& thetas, float sppScale, double sw, std::vector<double>& out) -> bool {
            const int spp = sppScale > 1.001f ? std::max(1, int(std::lround(fit.spp * sppScale))) : fit.spp;
            std::vector<Job> jobs;
            for (const auto& th : thetas) for (int v = 0; v < nViews; v++) { Job j{th, v, 0.0}; j.spp = spp; j.sw = sw; jobs.push_back(std::move(j)); }
            if (!runJobs(jobs)) return false;
            out.assign(thetas.size(), 0.0);
            for (size_t k = 0; k < thetas.size(); k++) {
                double L = 0.0, wSum = 0.0;
                for (int v = 0; v < nViews; v++) { const double w = v == 0 ? 1.0 : 0.5; L += w * jobs[k * size_t(nViews) + size_t(v)].loss; wSum += w; }      // :245-247
                out[k] = L / wSum + lightRegularizer(thetas[k]) + metalPrior(thetas[k]);
            }
            return true;
        };
        auto lossAtS = [&](const std::vector<double>& th, float sppScale, double sw, double& out) { std::vector<double> o; if (!imageLosses({th}, sppScale, sw, o)) return false; out = o[0]; return true; };
        double loss = 0.0; if (!lossAtS(sp.values, 1.0f, specularWeight, loss)) return 1;
        const double L0s = loss;
        std::cout << "highlight score " << highlightScore << ", initial loss " << loss << "\n";
        // multi-start probe (:398-520): init, mid-grey, conductor, rough dielectric + one seeded random candidate
        {
            std::vector<std::vector<double>> cand; cand.push_back(sp.values);
            { auto mid = sp.values; mid[0] = mid[1] = mid[2] = 0.45; mid[3] = 0.45; mid[4] = 0.25; mid[5] = mid[6] = mid[7] = 0.25; mid[8] = 0.50; mid[9] = 0.22; mid[10] = 0.25; mid[11] = 0.90; cand.push_back(mid); }
            { auto m = sp.values; m[3] = 0.12; m[4] = 0.85; cand.push_back(m); }
            { auto m = sp.values; m[3] = 0.80; m[4] = 0.05; cand.push_back(m); }
            { auto m = sp.values; uint32_t st = seed + 17u * 7u; auto rnd = [&]() { st = st * 1664525u + 1013904223u; return double(st >> 8) / 16777216.0; };
              for (size_t i = 0; i < m.size(); i++) { m[i] = sp.project(i, sp.lo[i] + rnd() * (sp.hi[i] - sp.lo[i])); }
              cand.push_back(m); }
            for (auto& th : cand) for (size_t i = 0; i < th.size(); i++) th[i] = sp.project(i, th[i]);
            const float probeSpp = highlightScore > 0.16 ? 2.0f : 1.25f; const double probeSw = highlightScore > 0.16 ? std::min(1.0, specularWeight + 0.2) : specularWeight * 0.6;
            std::vector<double> Lc; if (!imageLosses(cand, probeSpp, probeSw, Lc)) return 1;
            double bestStartLoss = loss; std::vector<double> bestStart = sp.values;
            for (size_t c = 0; c < cand.size(); c++) {
                if (highlightScore > 0.16) { if (cand[c][metalIdx] > 0.55 && cand[c][roughIdx] < 0.35) Lc[c] *= 0.82; if (cand[c][metalIdx] < 0.30) Lc[c] *= 1.18; }
                std::cout << "  candidate " << c + 1 << "/" << cand.size() << "  loss=" << Lc[c] << "\n";
                if (Lc[c] < bestStartLoss) { bestStartLoss = Lc[c]; bestStart = cand[c]; }
            }
            sp.values = bestStart; if (!lossAtS(sp.values, 1.0f, specularWeight, loss)) return 1;
        }
        double bestLoss = loss; std::vector<double> bestTheta = sp.values;
        auto runStage = [&](const char* name, const std::vector<size_t>& active, int stageIters, double lrMul, double epsMul, float sppScale = 1.0f, double sw = -1.0, int patience = 5) -> bool {   // :296-395
            if (active.empty() || stageIters <= 0) return true;
            if (sw < 0.0) sw = specularWeight;
            std::cout << "-- stage " << name << " (" << active.size() << " params, " << stageIters << " iters, lr x" << lrMul << (sppScale > 1.001f ? ", spp x" + std::to_string(sppScale) : std::string()) << ") --\n";
            AdamState adamS; int worse = 0;
            double stageBest; if (!lossAtS(sp.values, sppScale, sw, stageBest)) return false;
            std::vector<double> stageBestTh = sp.values;
            const double stageLr = lr * lrMul, stageEps = eps * epsMul;
            for (int it = 0; it < stageIters; ++it) {
                std::vector<std::vector<double>> thetas; std::vector<double> denom(active.size(), 0.0);
                for (size_t a = 0; a < active.size(); a++) {
                    const size_t ai = active[a]; const double v0 = sp.values[ai], span = std::max(1e-3, sp.hi[ai] - sp.lo[ai]), epsI = std::max(stageEps, 0.02 * span);
                    const double hi = sp.project(ai, v0 + epsI), lo = sp.project(ai, v0 - epsI); denom[a] = hi - lo;
                    auto th = sp.values; th[ai] = hi; thetas.push_back(th); th[ai] = lo; thetas.push_back(th);
                }
                std::vector<double> Ls; if (!imageLosses(thetas, sppScale, sw, Ls)) return false;      // 2 |active| nViews probes, fanned out over the GPUs
                std::vector<double> g(sp.size(), 0.0);
                for (size_t a = 0; a < active.size(); a++) if (denom[a] >= 1e-12) g[active[a]] = (Ls[2 * a] - Ls[2 * a + 1]) / denom[a];
                const std::vector<double> before = sp.values;
                adamS.step(sp, g, stageLr);
                for (size_t i = 0; i < sp.size(); i++) if (std::find(active.begin(), active.end(), i) == active.end()) sp.values[i] = before[i];
                if (!lossAtS(sp.values, sppScale, sw, loss)) return false;
                std::cout << "  [" << name << "] " << it + 1 << "/" << stageIters << "  loss=" << loss << "\n";
                if (loss + 1e-9 < stageBest) { stageBest = loss; stageBestTh = sp.values; worse = 0; } else ++worse;
                if (loss > 0.0 && loss < 5e-5) break;
                if (worse >= patience && stageBest < 2e-3) break;
            }
            sp.values = stageBestTh; loss = stageBest;
            if (stageBest < bestLoss) { bestLoss = stageBest; bestTheta = stageBestTh; }
            return true;
        };
        const std::vector<size_t> gAlbedo{0, 1, 2}, gBrdf{3, 4}, gPed{5, 6, 7}, gLight{8, 9, 10}, gEnv{11};
        const int envIters = std::max(6, iters * 15 / 100), lightIters = std::max(8, iters * 22 / 100), albedoIters = std::max(10, iters * 22 / 100), brdfIters = std::max(10, iters * 20 / 100),
                  pedestalIters = std::max(5, iters * 12 / 100), refineIters = std::max(5, iters * 12 / 100);      // :583-589
        bool ok = runStage("env", gEnv, envIters, 1.0, 1.0) && runStage("lights", gLight, lightIters, 0.85, 1.0);
        if (ok && highlightScore > 0.24) ok = runStage("brdf_pre", gBrdf, std::max(6, brdfIters / 2), 0.80, 0.65, std::max(brdfSppMul, 2.0f), std::min(1.0, specularWeight + 0.30));
        ok = ok && runStage("albedo", gAlbedo, albedoIters, 0.8, 0.9) && runStage("brdf", gBrdf, brdfIters, 0.75, 0.7, brdfSppMul, std::min(1.0, specularWeight + 0.1))
                && runStage("brdf2", gBrdf, std::max(4, brdfIters / 2), 0.45, 0.55, brdfSppMul, std::min(1.0, specularWeight + 0.15))
                && runStage("pedestal", gPed, pedestalIters, 0.65, 1.0) && runStage("lights2", gLight, std::max(3, lightIters / 2), 0.50, 0.75);
        { std::vector<size_t> refine = gAlbedo; refine.insert(refine.end(), gBrdf.begin(), gBrdf.end()); refine.insert(refine.end(), gEnv.begin(), gEnv.end()); refine.push_back(gLight.front());
          ok = ok && runStage("refine", refine, refineIters, 0.35, 0.55, 1.25f, specularWeight * 0.7); }
        if (ok && sp.values[metalIdx] > 0.35) {                                                           // metal lock, diffuse target (:765-778)
            std::cout << "  metal lock (diffuse): metal " << sp.values[metalIdx] << " -> 0.12\n";
            sp.values[metalIdx] = sp.project(metalIdx, 0.12); sp.values[roughIdx] = sp.project(roughIdx, std::max(sp.values[roughIdx], 0.32));
            ok = lossAtS(sp.values, 1.5f, specularWeight * 0.3, loss);
            if (ok && loss < bestLoss) { bestLoss = loss; bestTheta = sp.values; }
            ok = ok && runStage("metal_lock", gBrdf, std::max(4, brdfIters / 3), 0.40, 0.5, 1.25f, specularWeight * 0.35);
        }
        if (!ok) { std::cerr << "FATAL: probe render failed\n"; return 1; }
        sp.values = bestTheta;
        double Lf = 0.0; if (!lossAtS(sp.values, 1.0f, specularWeight, Lf)) return 1;
        double err = 0.0; for (size_t i = 0; i < sp.size(); i++) err += (sp.values[i] - truth[i]) * (sp.values[i] - truth[i]);
        std::cout << "final loss " << Lf << " (initial " << L0s << "), |theta - truth| = " << std::sqrt(err) << "\n";
        std::cout << "probes: " << probes << " in " << probeSeconds << " s = " << double(probes) / probeSeconds << " probes/s, " << double(probeSamples) / probeSeconds / 1e6 << " Msamples/s on " << G << " GPU(s)\n";
        std::cout << "theta:"; for (double v : sp.values) std::cout << " " << v; std::cout << "\n";
        return Lf < L0s ? 0 : 2;
    }
    auto lossAt = [&](const std::vector<double>& theta, double& out) {
        std::vector<Job> jobs; for (int v = 0; v < nViews; v++) jobs.push_back({theta, v, 0.0});
        if (!runJobs(jobs)) return false;
        out = 0.0; for (auto& j : jobs) out += j.loss / nViews;
        return true;
    };
    double L0 = 0.0; if (!lossAt(sp.values, L0)) return 1;
    std::cout << "iter 0  loss " << L0 << "\n";
    AdamState adam; probes = 0; probeSeconds = 0.0; probeSamples = 0;
    for (int it = 1; it <= iters; it++) {
        // finiteDiffGradient: 2 * |theta| * nViews independent probes
        std::vector<Job> jobs; std::vector<double> denom(sp.size(), 0.0);
        for (size_t i = 0; i < sp.size(); i++) {
            double v0 = sp.values[i], hi = sp.project(i, v0 + eps), lo = sp.project(i, v0 - eps);
            denom[i] = hi - lo;
            for (int side = 0; side < 2; side++) for (int v = 0; v < nViews; v++) { Job j{sp.values, v, 0.0}; j.theta[i] = side ? lo : hi; jobs.push_back(std::move(j)); }
        }
        if (!runJobs(jobs)) { std::cerr << "FATAL: probe render failed\n"; return 1; }
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

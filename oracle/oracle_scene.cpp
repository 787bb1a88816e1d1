// oracle/oracle_scene.cpp — TEST INFRASTRUCTURE ONLY (see oracle.cpp header).
#include "oracle_scene.h"
#include <numeric>
#include <limits>

namespace orc {

static const SobolTable g_sobol = {{
#include "sobol_dirs.inc"
}};
const SobolTable& sobolTable() { return g_sobol; }

// ---------------------------------------------------------------------------------------------
// EnvCDF::build restated (ohao/render/rt/env_cdf.cpp:13-61).  Sequential fp32 sums: the order
// is part of the contract (north_star: CDFs agree to 1e-6 relative).
// ---------------------------------------------------------------------------------------------
void buildEnvCDF(const float* rgba, int W, int H, std::vector<float>& marg, std::vector<float>& cond, float& integral) {
    cond.assign(size_t(W) * size_t(H), 0.0f);
    marg.assign(size_t(H), 0.0f);
    std::vector<float> rowTotal(size_t(H), 0.0f);
    const float pi = 3.14159265358979323846f;
    for (int y = 0; y < H; y++) {
        float sinTheta = std::sin(pi * (float(y) + 0.5f) / float(H));
        float* row = &cond[size_t(y) * size_t(W)];
        float run = 0.0f;
        for (int x = 0; x < W; x++) {
            const float* p = rgba + (size_t(y) * size_t(W) + size_t(x)) * 4u;
            float lum = 0.2126f * p[0] + 0.7152f * p[1] + 0.0722f * p[2];
            run += lum * sinTheta;
            row[x] = run;
        }
        if (run > 0.0f) for (int x = 0; x < W; x++) row[x] /= run;
        else            for (int x = 0; x < W; x++) row[x] = float(x + 1) / float(W);
        rowTotal[size_t(y)] = run;
    }
    float total = 0.0f;
    for (int y = 0; y < H; y++) { total += rowTotal[size_t(y)]; marg[size_t(y)] = total; }
    integral = total;
    if (total > 0.0f) for (int y = 0; y < H; y++) marg[size_t(y)] /= total;
    else              for (int y = 0; y < H; y++) marg[size_t(y)] = float(y + 1) / float(H);
}

// env_sampling.glsl:16-38 — lower-bound binary searches.
static int lowerBound(const float* a, int n, float u) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (a[mid] < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}
static const float kPi = 3.14159265358979f, kTwoPi = 6.28318530717959f;

static float texelPdf(const Scene& s, int x, int y, float sinT) {
    uint32_t W = s.envW, H = s.envH;
    size_t rowBase = size_t(y) * W;
    float condDiff = s.cond[rowBase + x] - (x > 0 ? s.cond[rowBase + x - 1] : 0.0f);
    float margDiff = s.marg[y] - (y > 0 ? s.marg[y - 1] : 0.0f);
    float pdfUV = condDiff * margDiff * float(W) * float(H);
    return std::max(pdfUV / (kTwoPi * kPi * sinT), 0.0f);
}
void sampleEnvMap(const Scene& s, float u1, float u2, V3& dir, float& pdf) {   // env_sampling.glsl:53-76
    uint32_t W = s.envW, H = s.envH;
    int y = lowerBound(s.marg.data(), int(H), u1);
    int x = lowerBound(s.cond.data() + size_t(y) * W, int(W), u2);
    float u = (float(x) + 0.5f) / float(W), v = (float(y) + 0.5f) / float(H);
    float phi = (u - 0.5f) * kTwoPi, theta = v * kPi;
    float sinT = std::sin(theta);
    dir = {sinT * std::cos(phi), std::cos(theta), sinT * std::sin(phi)};
    float theta2 = (float(y) + 0.5f) / float(H) * kPi;
    pdf = texelPdf(s, x, y, std::max(std::sin(theta2), 1e-4f));
}
float pdfEnvMap(const Scene& s, V3 dir) {                                      // env_sampling.glsl:79-94
    uint32_t W = s.envW, H = s.envH;
    float theta = std::acos(clampf(dir.y, -1.0f, 1.0f));
    float phi = std::atan2(dir.z, dir.x);
    float u = phi / kTwoPi + 0.5f, v = theta / kPi;
    int x = std::min(std::max(int(u * float(W)), 0), int(W) - 1);
    int y = std::min(std::max(int(v * float(H)), 0), int(H) - 1);
    return texelPdf(s, x, y, std::max(std::sin(theta), 1e-4f));
}

// ---------------------------------------------------------------------------------------------
// Texture fetch: VkSampler LINEAR/LINEAR, REPEAT (rt_build.cpp:627-639), unnormalised texel
// coordinate = uv*size - 0.5, weights from its fraction, wrap by modulo.  Full fp32 weights.
// ---------------------------------------------------------------------------------------------
struct Bilin { int x0, x1, y0, y1; float fx, fy; };
static Bilin bilin(V2 uv, int W, int H) {
    float u = uv.x - std::floor(uv.x), v = uv.y - std::floor(uv.y);
    float x = u * float(W) - 0.5f, y = v * float(H) - 0.5f;
    float xf = std::floor(x), yf = std::floor(y);
    Bilin b; b.fx = x - xf; b.fy = y - yf;
    int xi = int(xf), yi = int(yf);
    auto wrap = [](int i, int n) { i %= n; return i < 0 ? i + n : i; };
    b.x0 = wrap(xi, W); b.x1 = wrap(xi + 1, W); b.y0 = wrap(yi, H); b.y1 = wrap(yi + 1, H);
    return b;
}
static V4 lerp4(V4 a, V4 b, float t) { return {a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t}; }
V4 sampleLayer(const Scene& s, uint32_t layer, V2 uv) {
    if (layer >= s.texLayers) return {0, 0, 0, 0};
    int W = int(s.texW), H = int(s.texH);
    const uint8_t* base = s.tex.data() + size_t(layer) * size_t(W) * size_t(H) * 4u;
    Bilin b = bilin(uv, W, H);
    auto px = [&](int x, int y) { const uint8_t* p = base + (size_t(y) * W + x) * 4u;
        const float k = 1.0f / 255.0f; return V4{p[0] * k, p[1] * k, p[2] * k, p[3] * k}; };
    return lerp4(lerp4(px(b.x0, b.y0), px(b.x1, b.y0), b.fx), lerp4(px(b.x0, b.y1), px(b.x1, b.y1), b.fx), b.fy);
}
V4 sampleEnvTexture(const Scene& s, V2 uv) {
    int W = int(s.envW), H = int(s.envH);
    Bilin b = bilin(uv, W, H);
    auto px = [&](int x, int y) { const float* p = s.env.data() + (size_t(y) * W + x) * 4u; return V4{p[0], p[1], p[2], p[3]}; };
    return lerp4(lerp4(px(b.x0, b.y0), px(b.x1, b.y0), b.fx), lerp4(px(b.x0, b.y1), px(b.x1, b.y1), b.fx), b.fy);
}

// ---------------------------------------------------------------------------------------------
// Oracle-private BVH: top-down binned SAH over world-space triangles, leaves <= 4 triangles.
// Deliberately a different builder from the product's LBVH so a builder bug cannot cancel out.
// ---------------------------------------------------------------------------------------------
namespace {
struct Box { V3 lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
    void grow(V3 p) { lo = vmin(lo, p); hi = vmax(hi, p); }
    void grow(const Box& b) { lo = vmin(lo, b.lo); hi = vmax(hi, b.hi); }
    float area() const { V3 e = hi - lo; return (e.x < 0) ? 0.0f : 2.0f * (e.x * e.y + e.y * e.z + e.z * e.x); } };
}
static int buildRec(Scene& s, std::vector<Box>& tb, std::vector<V3>& cen, uint32_t first, uint32_t count) {
    int me = int(s.nodes.size());
    s.nodes.push_back({});
    Box b, cb;
    for (uint32_t i = first; i < first + count; i++) { b.grow(tb[s.bvhTris[i]]); cb.grow(cen[s.bvhTris[i]]); }
    // pad generously: the oracle only has to be conservative, not fast
    V3 ext = b.hi - b.lo; float pad = 1e-5f * (std::max(ext.x, std::max(ext.y, ext.z)) + maxcomp(vabs(b.lo)) + maxcomp(vabs(b.hi))) + 1e-7f;
    s.nodes[me].lo = b.lo - v3(pad); s.nodes[me].hi = b.hi + v3(pad);
    auto makeLeaf = [&]() { s.nodes[me].first = first; s.nodes[me].count = count; s.nodes[me].left = s.nodes[me].right = -1; return me; };
    if (count <= 4) return makeLeaf();
    V3 ce = cb.hi - cb.lo;
    int axis = (ce.x >= ce.y && ce.x >= ce.z) ? 0 : (ce.y >= ce.z ? 1 : 2);
    float cmin = get(cb.lo, axis), cext = get(ce, axis);
    uint32_t mid = first + count / 2;
    if (cext > 0.0f) {
        const int NB = 16; Box bb[NB]; int bc[NB] = {0};
        for (uint32_t i = first; i < first + count; i++) {
            uint32_t t = s.bvhTris[i];
            int k = std::min(NB - 1, int(NB * ((get(cen[t], axis) - cmin) / cext)));
            bb[k].grow(tb[t]); bc[k]++;
        }
        float lA[NB], rA[NB]; int lC[NB], rC[NB]; Box acc; int c = 0;
        for (int k = 0; k < NB; k++) { acc.grow(bb[k]); c += bc[k]; lA[k] = acc.area(); lC[k] = c; }
        acc = Box(); c = 0;
        for (int k = NB - 1; k >= 0; k--) { acc.grow(bb[k]); c += bc[k]; rA[k] = acc.area(); rC[k] = c; }
        float best = 1e30f; int bk = -1;
        for (int k = 0; k < NB - 1; k++) { if (!lC[k] || !rC[k + 1]) continue; float cost = lA[k] * lC[k] + rA[k + 1] * rC[k + 1]; if (cost < best) { best = cost; bk = k; } }
        if (bk >= 0) {
            auto it = std::partition(s.bvhTris.begin() + first, s.bvhTris.begin() + first + count, [&](uint32_t t) {
                return std::min(NB - 1, int(NB * ((get(cen[t], axis) - cmin) / cext))) <= bk; });
            mid = uint32_t(it - s.bvhTris.begin());
        }
    }
    if (mid == first || mid == first + count) {
        mid = first + count / 2;
        std::nth_element(s.bvhTris.begin() + first, s.bvhTris.begin() + mid, s.bvhTris.begin() + first + count,
                         [&](uint32_t a, uint32_t bq) { return get(cen[a], axis) < get(cen[bq], axis); });
    }
    int l = buildRec(s, tb, cen, first, mid - first);
    int r = buildRec(s, tb, cen, mid, first + count - mid);
    s.nodes[me].left = l; s.nodes[me].right = r; s.nodes[me].count = 0;
    return me;
}
void buildBvh(Scene& s) {
    uint32_t T = uint32_t(s.idx.size() / 3);
    std::vector<Box> tb(T); std::vector<V3> cen(T);
    s.bvhTris = s.activeTris;
    for (uint32_t t : s.activeTris) {
        for (int k = 0; k < 3; k++) tb[t].grow(s.wtri[size_t(t) * 3 + k]);
        cen[t] = (tb[t].lo + tb[t].hi) * 0.5f;
    }
    s.nodes.clear();
    if (!s.bvhTris.empty()) buildRec(s, tb, cen, 0, uint32_t(s.bvhTris.size()));
}

static inline bool slab(const BvhNode& n, const RayPrep& r, float tmax, float& tn) {
    float tx1 = (n.lo.x - r.o.x) * r.idir.x, tx2 = (n.hi.x - r.o.x) * r.idir.x;
    float ty1 = (n.lo.y - r.o.y) * r.idir.y, ty2 = (n.hi.y - r.o.y) * r.idir.y;
    float tz1 = (n.lo.z - r.o.z) * r.idir.z, tz2 = (n.hi.z - r.o.z) * r.idir.z;
    float tnear = std::max(std::max(std::min(tx1, tx2), std::min(ty1, ty2)), std::min(tz1, tz2));
    float tfar  = std::min(std::min(std::max(tx1, tx2), std::max(ty1, ty2)), std::max(tz1, tz2));
    // conservative widening (oracle: generous)
    tnear = tnear - std::fabs(tnear) * 1e-5f - 1e-6f;
    tfar  = tfar + std::fabs(tfar) * 1e-5f + 1e-6f;
    tn = tnear;
    return tfar >= std::max(tnear, r.tmin - 1e-6f) && tnear <= tmax;
}

// Two-level spec (shared with ohb_traverse.h travEnterInstance): the ray is mapped into each instance's object space with
// o' = ((m0 x + m1 y) + m2 z) + m3, d' = (m0 x + m1 y) + m2 z over the rows of world->object (one rounding per operation, no
// renormalisation: t is the same parameter in both spaces), the watertight test runs on the OBJECT-space triangles, and the
// closest candidate over all instances wins, equal t resolved toward the lower global id.
static void toObject(const Instance& in, V3 o, V3 d, V3& oo, V3& od) {
    const float* m = in.inv;
    oo = {((m[0] * o.x + m[1] * o.y) + m[2] * o.z) + m[3], ((m[4] * o.x + m[5] * o.y) + m[6] * o.z) + m[7], ((m[8] * o.x + m[9] * o.y) + m[10] * o.z) + m[11]};
    od = {(m[0] * d.x + m[1] * d.y) + m[2] * d.z, (m[4] * d.x + m[5] * d.y) + m[6] * d.z, (m[8] * d.x + m[9] * d.y) + m[10] * d.z};
}
void buildTwoLevel(Scene& s) {
    s.blas.clear();
    for (uint32_t i = 0; i < s.inst.size(); i++) {
        const Instance& in = s.inst[i];
        if ((in.mask & 0xFFu) == 0u || in.triCount == 0u) continue;
        Scene::Blas bl; bl.inst = i; bl.firstTri = in.firstTri; bl.sub = std::make_unique<Scene>();
        Scene& sub = *bl.sub;
        sub.idx.resize(size_t(in.triCount) * 3); sub.wtri.resize(size_t(in.triCount) * 3);
        for (uint32_t t = 0; t < in.triCount; t++) {
            for (int k = 0; k < 3; k++) sub.wtri[size_t(t) * 3 + k] = s.pos[s.idx[size_t(in.firstTri + t) * 3 + k]];      // object space, untransformed
            sub.activeTris.push_back(t);
        }
        buildBvh(sub);
        s.blas.push_back(std::move(bl));
    }
}
static ohb_hit traceClosestFlat(const Scene& s, V3 o, V3 d, float tmin, float tmax);
static bool traceAnyFlat(const Scene& s, V3 o, V3 d, float tmin, float tmax);
ohb_hit traceClosest(const Scene& s, V3 o, V3 d, float tmin, float tmax) {
    if (!s.twoLevel) return traceClosestFlat(s, o, d, tmin, tmax);
    ohb_hit best{-1.0f, 0.0f, 0.0f, OHB_MISS};
    for (const Scene::Blas& bl : s.blas) {
        V3 oo, od; toObject(s.inst[bl.inst], o, d, oo, od);
        ohb_hit h = traceClosestFlat(*bl.sub, oo, od, tmin, tmax);
        if (h.prim == OHB_MISS) continue;
        uint32_t gid = bl.firstTri + h.prim;
        if (best.prim == OHB_MISS || h.t < best.t || (h.t == best.t && gid < best.prim)) best = {h.t, h.u, h.v, gid};
    }
    return best;
}
bool traceAny(const Scene& s, V3 o, V3 d, float tmin, float tmax) {
    if (!s.twoLevel) return traceAnyFlat(s, o, d, tmin, tmax);
    for (const Scene::Blas& bl : s.blas) {
        V3 oo, od; toObject(s.inst[bl.inst], o, d, oo, od);
        if (traceAnyFlat(*bl.sub, oo, od, tmin, tmax)) return true;
    }
    return false;
}
static ohb_hit traceClosestFlat(const Scene& s, V3 o, V3 d, float tmin, float tmax) {
    ohb_hit best{tmax, 0.0f, 0.0f, OHB_MISS};
    if (s.nodes.empty()) { best.t = -1.0f; return best; }
    RayPrep r = prepRay(o, d, tmin);
    int stack[128]; int sp = 0; stack[sp++] = 0;
    float dummy;
    while (sp) {
        const BvhNode& n = s.nodes[stack[--sp]];
        if (!slab(n, r, best.t, dummy)) continue;
        if (n.count) {
            for (uint32_t i = n.first; i < n.first + n.count; i++) {
                uint32_t t = s.bvhTris[i];
                float tt, bu, bv;
                // test against tmax (not best.t) so that equal-t ties can be resolved by id
                if (!intersectTri(r, s.wtri[size_t(t) * 3], s.wtri[size_t(t) * 3 + 1], s.wtri[size_t(t) * 3 + 2], tmax, tt, bu, bv)) continue;
                if (best.prim == OHB_MISS || tt < best.t || (tt == best.t && t < best.prim)) best = {tt, bu, bv, t};
            }
        } else {
            float tl, tr;
            bool hl = slab(s.nodes[n.left], r, best.t, tl), hr = slab(s.nodes[n.right], r, best.t, tr);
            if (hl && hr) { if (tl < tr) { stack[sp++] = n.right; stack[sp++] = n.left; } else { stack[sp++] = n.left; stack[sp++] = n.right; } }
            else if (hl) stack[sp++] = n.left;
            else if (hr) stack[sp++] = n.right;
        }
    }
    if (best.prim == OHB_MISS) best.t = -1.0f;
    return best;
}
ohb_hit traceClosestBrute(const Scene& s, V3 o, V3 d, float tmin, float tmax) {
    ohb_hit best{tmax, 0.0f, 0.0f, OHB_MISS};
    RayPrep r = prepRay(o, d, tmin);
    for (uint32_t t : s.activeTris) {
        float tt, bu, bv;
        if (!intersectTri(r, s.wtri[size_t(t) * 3], s.wtri[size_t(t) * 3 + 1], s.wtri[size_t(t) * 3 + 2], tmax, tt, bu, bv)) continue;
        if (best.prim == OHB_MISS || tt < best.t || (tt == best.t && t < best.prim)) best = {tt, bu, bv, t};
    }
    if (best.prim == OHB_MISS) best.t = -1.0f;
    return best;
}
static bool traceAnyFlat(const Scene& s, V3 o, V3 d, float tmin, float tmax) {
    if (s.nodes.empty()) return false;
    RayPrep r = prepRay(o, d, tmin);
    int stack[128]; int sp = 0; stack[sp++] = 0;
    float dummy;
    while (sp) {
        const BvhNode& n = s.nodes[stack[--sp]];
        if (!slab(n, r, tmax, dummy)) continue;
        if (n.count) {
            for (uint32_t i = n.first; i < n.first + n.count; i++) {
                uint32_t t = s.bvhTris[i];
                float tt, bu, bv;
                if (intersectTri(r, s.wtri[size_t(t) * 3], s.wtri[size_t(t) * 3 + 1], s.wtri[size_t(t) * 3 + 2], tmax, tt, bu, bv)) return true;
            }
        } else { stack[sp++] = n.left; stack[sp++] = n.right; }
    }
    return false;
}

}  // namespace orc

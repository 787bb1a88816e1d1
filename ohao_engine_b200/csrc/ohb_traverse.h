// ohb_traverse.h — stack-based BVH2 traversal + watertight ray/triangle test (per-thread code).
//
// Replaces traceRayEXT (pt_raygen_offline.rgen:198-199 closest, :394-400 any-hit), whose
// arithmetic lives in the Vulkan driver.  The intersection arithmetic is the spec shared with the
// oracle (oracle/oracle_scene.h): Woop-Benthin-Wald watertight test, one fp32 rounding per
// operation (x* helpers never contract to FMA), fp64 fallback when an edge function is exactly
// zero, two-sided, accept tmin < t < tmax, equal-t ties resolved toward the lower global id.
//
// Node fetch = 4 x LDG.128 (64 B), triangle fetch = 3 x LDG.128 (48 B), both 16-B aligned.
#pragma once
#include "ohb_scene.h"

namespace ohb {

#define OHB_STACK_SIZE 64
#define OHB_MAX_LEAF 4

// leaf reference: ~((first << 2) | (count - 1)), always negative
OHB_HD int32_t makeLeafRef(uint32_t first, uint32_t count) { return ~int32_t((first << 2) | (count - 1u)); }
OHB_HD uint32_t leafFirst(int32_t ref) { return uint32_t(~ref) >> 2; }
OHB_HD uint32_t leafCount(int32_t ref) { return (uint32_t(~ref) & 3u) + 1u; }

struct RayPrep {
    f3 o, d; float tmin;
    int kx, ky, kz; float Sx, Sy, Sz;
    f3 idir;
};
OHB_HD RayPrep prepRay(f3 o, f3 d, float tmin) {
    RayPrep r; r.o = o; r.d = d; r.tmin = tmin;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = 0; float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; m = az; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    if (comp(d, kz) < 0.0f) { int t = kx; kx = ky; ky = t; }
    r.kx = kx; r.ky = ky; r.kz = kz;
    float dz = comp(d, kz);
    r.Sx = xdiv(comp(d, kx), dz); r.Sy = xdiv(comp(d, ky), dz); r.Sz = xdiv(1.0f, dz);
    float sx = fabsf(d.x) > 1e-20f ? d.x : copysignf(1e-20f, d.x);
    float sy = fabsf(d.y) > 1e-20f ? d.y : copysignf(1e-20f, d.y);
    float sz = fabsf(d.z) > 1e-20f ? d.z : copysignf(1e-20f, d.z);
    r.idir = mk3(1.0f / sx, 1.0f / sy, 1.0f / sz);
    return r;
}

OHB_HD bool intersectTri(const RayPrep& r, f3 p0, f3 p1, f3 p2, float tmax, float& t, float& bu, float& bv) {
    f3 A = mk3(xsub(p0.x, r.o.x), xsub(p0.y, r.o.y), xsub(p0.z, r.o.z));
    f3 B = mk3(xsub(p1.x, r.o.x), xsub(p1.y, r.o.y), xsub(p1.z, r.o.z));
    f3 C = mk3(xsub(p2.x, r.o.x), xsub(p2.y, r.o.y), xsub(p2.z, r.o.z));
    float Akz = comp(A, r.kz), Bkz = comp(B, r.kz), Ckz = comp(C, r.kz);
    float Ax = xsub(comp(A, r.kx), xmul(r.Sx, Akz)), Ay = xsub(comp(A, r.ky), xmul(r.Sy, Akz));
    float Bx = xsub(comp(B, r.kx), xmul(r.Sx, Bkz)), By = xsub(comp(B, r.ky), xmul(r.Sy, Bkz));
    float Cx = xsub(comp(C, r.kx), xmul(r.Sx, Ckz)), Cy = xsub(comp(C, r.ky), xmul(r.Sy, Ckz));
    float U = xsub(xmul(Cx, By), xmul(Cy, Bx));
    float V = xsub(xmul(Ax, Cy), xmul(Ay, Cx));
    float W = xsub(xmul(Bx, Ay), xmul(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = float(double(Cx) * double(By) - double(Cy) * double(Bx));
        V = float(double(Ax) * double(Cy) - double(Ay) * double(Cx));
        W = float(double(Bx) * double(Ay) - double(By) * double(Ax));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = xadd(xadd(U, V), W);
    if (det == 0.0f) return false;
    float Az = xmul(r.Sz, Akz), Bz = xmul(r.Sz, Bkz), Cz = xmul(r.Sz, Ckz);
    float T = xadd(xadd(xmul(U, Az), xmul(V, Bz)), xmul(W, Cz));
    float tt = xdiv(T, det);
    if (!(tt > r.tmin && tt < tmax)) return false;
    t = tt; bu = xdiv(V, det); bv = xdiv(W, det);
    return true;
}

// Conservative slab test on [tlo, thi]; (lo - o) * idir keeps the relative error of each plane
// distance at 2^-23, and the interval is widened by ~2.5 ulp so a box is never culled wrongly.
OHB_HD bool slab(float lox, float hix, float loy, float hiy, float loz, float hiz, const RayPrep& r, float tlo, float thi, float& tn) {
    float tx1 = (lox - r.o.x) * r.idir.x, tx2 = (hix - r.o.x) * r.idir.x;
    float ty1 = (loy - r.o.y) * r.idir.y, ty2 = (hiy - r.o.y) * r.idir.y;
    float tz1 = (loz - r.o.z) * r.idir.z, tz2 = (hiz - r.o.z) * r.idir.z;
    float tnear = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fminf(tz1, tz2));
    float tfar  = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fmaxf(tz1, tz2));
    tnear -= fabsf(tnear) * 3e-7f;
    tfar  += fabsf(tfar) * 3e-7f;
    tn = tnear;
    return fmaxf(tnear, tlo) <= fminf(tfar, thi);
}

OHB_HD void leafClosest(const SceneDev& s, const RayPrep& r, int32_t ref, float tmax, ohb_hit& best) {
    uint32_t first = leafFirst(ref), cnt = leafCount(ref);
    for (uint32_t i = 0; i < cnt; i++) {
        const f4* tp = s.tris + size_t(first + i) * 3u;
        f4 v0 = ld4(tp), v1 = ld4(tp + 1), v2 = ld4(tp + 2);
        float tt, bu, bv;
        // tested against the caller's tmax (not best.t) so equal-t ties resolve by id
        if (!intersectTri(r, xyz(v0), xyz(v1), xyz(v2), tmax, tt, bu, bv)) continue;
        uint32_t id = f2u(v0.w);
        if (best.prim == OHB_MISS || tt < best.t || (tt == best.t && id < best.prim)) { best.t = tt; best.u = bu; best.v = bv; best.prim = id; }
    }
}
OHB_HD bool leafAny(const SceneDev& s, const RayPrep& r, int32_t ref, float tmax) {
    uint32_t first = leafFirst(ref), cnt = leafCount(ref);
    for (uint32_t i = 0; i < cnt; i++) {
        const f4* tp = s.tris + size_t(first + i) * 3u;
        f4 v0 = ld4(tp), v1 = ld4(tp + 1), v2 = ld4(tp + 2);
        float tt, bu, bv;
        if (intersectTri(r, xyz(v0), xyz(v1), xyz(v2), tmax, tt, bu, bv)) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// Resumable "while-while" traversal with postponed leaves (Aila & Laine 2009, "Understanding the
// efficiency of ray traversal on GPUs"): a lane walks inner nodes until it holds a leaf, keeps walking
// speculatively until every lane of the warp holds one, and only then are triangles tested — so the
// long watertight test runs with most lanes active instead of 1.6 of 32 (ncu r01).
// The state lives in registers + a local-memory stack and survives a pause, which lets the persistent
// kernels refill idle lanes with new rays when too few lanes of a warp are still traversing.
// ---------------------------------------------------------------------------------------------
#define OHB_TRAV_SENTINEL 0x7FFFFFFF
#if OHB_DEVICE_CODE
#define OHB_WARP_ANY(pred) (__ballot_sync(__activemask(), (pred)) != 0u)
#define OHB_WARP_ACTIVE() (__popc(__activemask()))
#else
#define OHB_WARP_ANY(pred) (pred)
#define OHB_WARP_ACTIVE() 32
#endif

struct Trav {
    RayPrep r; float tmax; ohb_hit best;
    int32_t node, leaf; int sp; bool anyHit;
};
// The stack is a separate local array (a struct holding a dynamically indexed array is demoted to local memory whole).
OHB_HD int32_t travPop(Trav& t, const int32_t* stack) { return t.sp ? stack[--t.sp] : OHB_TRAV_SENTINEL; }
OHB_HD void travInit(Trav& t, const SceneDev& s, f3 o, f3 d, float tmin, float tmax) {
    t.r = prepRay(o, d, tmin); t.tmax = tmax;
    t.best.t = tmax; t.best.u = 0.0f; t.best.v = 0.0f; t.best.prim = OHB_MISS;
    t.sp = 0; t.leaf = 0; t.anyHit = false;
    t.node = s.numTris ? s.rootRef : OHB_TRAV_SENTINEL;
    if (t.node < 0) { t.leaf = t.node; t.node = OHB_TRAV_SENTINEL; }   // the whole scene is one leaf
}
// Runs until the query is finished (returns true) or fewer than `minActive` lanes of the warp are still
// traversing (returns false; call again later).  ANY = TerminateOnFirstHit.
template <bool ANY>
OHB_HD bool travRun(Trav& t, int32_t* stack, const SceneDev& s, int minActive) {
    for (;;) {
        while (t.node >= 0 && t.node != OHB_TRAV_SENTINEL) {
            const f4* np = s.nodes + size_t(t.node) * 4u;
            f4 n0 = ld4(np), n1 = ld4(np + 1), n2 = ld4(np + 2), n3 = ld4(np + 3);
            float t0, t1;
            // <= best.t (not <): a box touching at exactly best.t may hold an equal-t, lower-id triangle
            bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, t.r, t.r.tmin, t.best.t, t0);
            bool h1 = slab(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, t.r, t.r.tmin, t.best.t, t1);
            int32_t c0 = int32_t(f2u(n3.x)), c1 = int32_t(f2u(n3.y));
            if (h0 && h1) {
                if (t1 < t0) { int32_t x = c0; c0 = c1; c1 = x; }
                if (t.sp < OHB_STACK_SIZE) stack[t.sp++] = c1;
                t.node = c0;
            } else if (h0) t.node = c0;
            else if (h1) t.node = c1;
            else t.node = travPop(t, stack);
            if (t.node < 0 && t.leaf == 0) { t.leaf = t.node; t.node = travPop(t, stack); }   // postpone the first leaf, keep walking
            if (!OHB_WARP_ANY(t.leaf == 0)) break;                                       // every lane holds a leaf
        }
        while (t.leaf < 0) {
            if (ANY) { if (leafAny(s, t.r, t.leaf, t.tmax)) { t.anyHit = true; t.node = OHB_TRAV_SENTINEL; t.leaf = 0; t.sp = 0; return true; } }
            else leafClosest(s, t.r, t.leaf, t.tmax, t.best);
            t.leaf = 0;
            if (t.node < 0) { t.leaf = t.node; t.node = travPop(t, stack); }
        }
        if (t.node == OHB_TRAV_SENTINEL) return true;
        if (OHB_WARP_ACTIVE() < minActive) return false;
    }
}

// Closest hit.  Returns prim == OHB_MISS and t = -1 on miss.
OHB_HD ohb_hit traceClosest(const SceneDev& s, f3 o, f3 d, float tmin, float tmax) {
    Trav t; int32_t stack[OHB_STACK_SIZE]; travInit(t, s, o, d, tmin, tmax);
    while (!travRun<false>(t, stack, s, 0)) {}
    if (t.best.prim == OHB_MISS) t.best.t = -1.0f;
    return t.best;
}
// Any hit in (tmin, tmax): TerminateOnFirstHit | SkipClosestHit.
OHB_HD bool traceAny(const SceneDev& s, f3 o, f3 d, float tmin, float tmax) {
    Trav t; int32_t stack[OHB_STACK_SIZE]; travInit(t, s, o, d, tmin, tmax);
    while (!travRun<true>(t, stack, s, 0)) {}
    return t.anyHit;
}

}  // namespace ohb

// ohb_traverse.h — compressed 8-wide BVH traversal + watertight ray/triangle test (per-thread code).
//
// Replaces traceRayEXT (pt_raygen_offline.rgen:198-199 closest, :394-400 any-hit), whose
// arithmetic lives in the Vulkan driver.  The intersection arithmetic is the spec shared with the
// oracle (oracle/oracle_scene.h): Woop-Benthin-Wald watertight test, one fp32 rounding per
// operation (x* helpers never contract to FMA), fp64 fallback when an edge function is exactly
// zero, two-sided, accept tmin < t < tmax, equal-t ties resolved toward the lower global id.
//
// Acceleration structure: 8-wide BVH with child boxes quantised to 8 bits on a per-node grid
// (Ylitie, Karras, Laine 2017, "Efficient incoherent ray traversal on GPUs through compressed wide
// BVHs").  One node = 80 B = 5 x LDG.128:
//   w0 : p.x, p.y, p.z (fp32 grid origin), {ex, ey, ez, imask} bytes (biased exponents of the grid step,
//        bit s of imask = child slot s is an inner node)
//   w1 : index of the first inner child, index of the first leaf triangle, meta[0..3], meta[4..7]
//        meta = 0 empty | 001_11sss inner child in slot s | uuu_ooooo leaf: unary triangle count, offset
//   w2 : qlo.x[0..7], qlo.y[0..7]     w3 : qlo.z[0..7], qhi.x[0..7]     w4 : qhi.y[0..7], qhi.z[0..7]
// Children sit in slots so that (slot ^ ray octant) is a front-to-back order; a node visit yields ONE
// bit mask of hit children (top byte: inner children in visit order, low 24 bits: leaf triangles), so the
// stack holds at most one entry per tree level (plus postponed triangle groups) and no distances.
// Triangle fetch = 3 x LDG.128 (48 B), 16-B aligned.
//
// Conservative boxes: the builder rounds the quantised planes outward (checked in fp64) around boxes
// already padded by 2^-20 (ohb_bvh.h padBox), and every slab interval is widened by more than the rounding
// error of its plane distances plus the placement error of the triangle test (intersectWideNode).  A box
// is therefore never culled wrongly and the hit (t, u, v, id) is independent of the tree: this 8-wide
// LBVH and the oracle's binary SAH tree give bit-identical results.
#pragma once
#include "ohb_scene.h"

namespace ohb {

#define OHB_STACK_SIZE 40       // >= OHB_MAX_LEVELS + OHB_POSTPONE_SLOTS
#define OHB_MAX_LEVELS 32       // deepest 8-wide tree the traversal stack is sized for (ohb_build_accel checks)
#define OHB_POSTPONE_SLOTS 8    // triangle groups may be postponed only while sp < this
#define OHB_MAX_LEAF 3          // triangles per leaf child (3-bit unary count)
#define OHB_WNODE_VECS 5        // 16-B words per node

struct alignas(8) u2 { uint32_t x, y; };

OHB_HD uint32_t popc32(uint32_t v) {
#if OHB_DEVICE_CODE
    return uint32_t(__popc(v));
#else
    return uint32_t(__builtin_popcount(v));
#endif
}
OHB_HD uint32_t bfind32(uint32_t v) { return 31u - uint32_t(clz32(v)); }   // v != 0
// per byte: 0xFF where bit 7 is set, else 0x00
OHB_HD uint32_t signExtendBytes(uint32_t v) {
#if OHB_DEVICE_CODE
    uint32_t r; asm("prmt.b32 %0, %1, 0, 0xba98;" : "=r"(r) : "r"(v)); return r;
#else
    return ((v >> 7) & 0x01010101u) * 0xFFu;
#endif
}
// 65536 + 2 * (byte j of v) as fp32, built by ONE byte permute: the byte lands in mantissa bits 8..15 of 2^16
// (I2F.U8 is a quarter-rate XU instruction; 48 of them per node visit made the XU pipe the limiter).
OHB_HD float biasedByte(uint32_t v, int j) {
#if OHB_DEVICE_CODE
    return __uint_as_float(__byte_perm(v, 0x47800000u, 0x7604u | (uint32_t(j) << 4)));
#else
    return u2f(0x47800000u | (((v >> (8 * j)) & 0xFFu) << 8));
#endif
}
// byte j (0..7) of the 64-bit value hi:lo
OHB_HD uint32_t byteOf64(uint32_t lo, uint32_t hi, uint32_t j) {
#if OHB_DEVICE_CODE
    return __byte_perm(lo, hi, j) & 0xFFu;
#else
    return uint32_t(((uint64_t(hi) << 32) | lo) >> (8u * j)) & 0xFFu;
#endif
}

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

// ---------------------------------------------------------------------------------------------
// Resumable traversal.  State = current node group G (x: index of the first inner child, y: hit bits of
// the inner children in the top byte | imask), current triangle group Gt (x: first triangle, y: bits),
// a local-memory stack of groups.  It survives a pause, which lets the persistent kernels refill idle
// lanes with new rays when too few lanes of a warp are still traversing.
// ---------------------------------------------------------------------------------------------
#if OHB_DEVICE_CODE
#define OHB_WARP_ACTIVE() (__popc(__activemask()))
#elif !defined(OHB_WARP_ACTIVE)
#define OHB_WARP_ACTIVE() 32       // host build (tests/emul may supply its own stand-in)
#endif

// tests/emul counts node visits / triangle tests per query through these hooks; they compile to nothing in the product
#ifndef OHB_STAT_NODE
#define OHB_STAT_NODE()
#define OHB_STAT_TRI()
#endif

struct Trav {
    RayPrep r; float tmax; ohb_hit best;
    u2 G, Gt; int sp; uint32_t octinv; bool anyHit;
};
typedef u2 TravStackEntry;

OHB_HD void travInit(Trav& t, const SceneDev& s, f3 o, f3 d, float tmin, float tmax) {
    t.r = prepRay(o, d, tmin); t.tmax = tmax;
    t.best.t = tmax; t.best.u = 0.0f; t.best.v = 0.0f; t.best.prim = OHB_MISS;
    t.sp = 0; t.anyHit = false;
    // bit 2/1/0 set = the ray travels toward +x/+y/+z: slot ^ octinv is then the visit priority (7 first)
    uint32_t octinv = (t.r.idir.x < 0.0f ? 0u : 4u) | (t.r.idir.y < 0.0f ? 0u : 2u) | (t.r.idir.z < 0.0f ? 0u : 1u);
    t.octinv = octinv;
    t.G.x = 0u; t.G.y = s.numTris ? 0x80000000u : 0u;      // the root is "slot 7 ^ octinv" of a virtual group
    t.Gt.x = 0u; t.Gt.y = 0u;
}

// One node visit: tests the 8 quantised child boxes against [tmin, thi].  Returns the inner children that were
// hit as a byte in bits 24..31, bit 24 + (slot ^ octinv) (= visit priority), and the hit leaf triangles in
// triMask (bit k = triangle triBase + k).
// Plane distance = fma(65536 + 2q, s/2, a - 32768 s) = q s + a with one rounding; arithmetic error per plane
// (s = 2^e/d, a = (p-o)/d): 1.8e-7|a| (offset) + 1.5e-5|s| (1/d in q s) + 6e-8|a| + 2e-3|s| (folded bias) +
// 6e-8|a| + 1.5e-5|s| (final rounding) + 6e-8|a| (a -+ err)  <  4e-7|a| + 4e-3|s|.
// The watertight test itself places a hit only to within a few ulp of the DISTANCE to the triangle's vertices
// (its t is a barycentric mean of vertex depths whose weights carry the cancellation error of the sheared
// coordinates), on every axis — a flat, axis-aligned triangle through the origin has a zero-width slab that
// this error exceeds.  Each slab is therefore widened by 1e-6 x the largest per-axis distance from the ray
// origin to the far side of the node, converted to ray-parameter units (>= the 4e-7|a| above).
OHB_HD uint32_t intersectWideNode(const Trav& t, const u4* np, float thi, uint32_t& childBase, uint32_t& triBase, uint32_t& imask, uint32_t& triMask) {
    const u4 w0 = ldu4(np), w1 = ldu4(np + 1), w2 = ldu4(np + 2), w3 = ldu4(np + 3), w4 = ldu4(np + 4);
    OHB_STAT_NODE();
    childBase = w1.x; triBase = w1.y; imask = w0.w >> 24;
    const RayPrep& r = t.r;
    // grid step and origin in ray-parameter units
    const float gx = u2f((w0.w & 0xFFu) << 23), gy = u2f(((w0.w >> 8) & 0xFFu) << 23), gz = u2f(((w0.w >> 16) & 0xFFu) << 23);
    const float px = u2f(w0.x) - r.o.x, py = u2f(w0.y) - r.o.y, pz = u2f(w0.z) - r.o.z;
    const float sx = gx * r.idir.x, sy = gy * r.idir.y, sz = gz * r.idir.z;
    const float ax = px * r.idir.x, ay = py * r.idir.y, az = pz * r.idir.z;
    // widening: 1e-6 of the farthest the node can be from the origin (any axis), see the header comment
    const float dist = 1e-6f * fmaxf(fmaxf(fmaf(256.0f, gx, fabsf(px)), fmaf(256.0f, gy, fabsf(py))), fmaf(256.0f, gz, fabsf(pz)));
    const float ex = fmaf(fabsf(sx), 4e-3f, dist * fabsf(r.idir.x)), ey = fmaf(fabsf(sy), 4e-3f, dist * fabsf(r.idir.y)), ez = fmaf(fabsf(sz), 4e-3f, dist * fabsf(r.idir.z));
    const float axn = fmaf(-32768.0f, sx, ax - ex), axf = fmaf(-32768.0f, sx, ax + ex);
    const float ayn = fmaf(-32768.0f, sy, ay - ey), ayf = fmaf(-32768.0f, sy, ay + ey);
    const float azn = fmaf(-32768.0f, sz, az - ez), azf = fmaf(-32768.0f, sz, az + ez);
    const float hx = sx * 0.5f, hy = sy * 0.5f, hz = sz * 0.5f;
    const bool nx = r.idir.x < 0.0f, ny = r.idir.y < 0.0f, nz = r.idir.z < 0.0f;
    uint32_t h = 0u;                                     // bit s = the box in slot s was hit
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int g = 0; g < 2; g++) {
        const uint32_t qlox = g ? w2.y : w2.x, qloy = g ? w2.w : w2.z, qloz = g ? w3.y : w3.x;
        const uint32_t qhix = g ? w3.w : w3.z, qhiy = g ? w4.y : w4.x, qhiz = g ? w4.w : w4.z;
        const uint32_t qnx = nx ? qhix : qlox, qfx = nx ? qlox : qhix;
        const uint32_t qny = ny ? qhiy : qloy, qfy = ny ? qloy : qhiy;
        const uint32_t qnz = nz ? qhiz : qloz, qfz = nz ? qloz : qhiz;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int j = 0; j < 4; j++) {
            const float tnx = fmaf(biasedByte(qnx, j), hx, axn), tfx = fmaf(biasedByte(qfx, j), hx, axf);
            const float tny = fmaf(biasedByte(qny, j), hy, ayn), tfy = fmaf(biasedByte(qfy, j), hy, ayf);
            const float tnz = fmaf(biasedByte(qnz, j), hz, azn), tfz = fmaf(biasedByte(qfz, j), hz, azf);
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, r.tmin));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, thi));
            // <= (not <): a box touching at exactly best.t may hold an equal-t, lower-id triangle
            if (tn <= tf) h |= 1u << (4 * g + j);
        }
    }
    // leaf children: expand each hit slot into its triangle bits (few per visit, so a loop over the set bits)
    triMask = 0u;
    for (uint32_t lh = h & ~imask; lh; lh &= lh - 1u) {
        const uint32_t m = byteOf64(w1.z, w1.w, bfind32(lh & (0u - lh)));
        triMask |= (m >> 5) << (m & 31u);
    }
    // inner children: move bit s to bit s ^ octinv (three conditional block swaps)
    uint32_t x = h & imask;
    const uint32_t o = t.octinv;
    uint32_t u;
    u = ((x >> 4) ^ x) & ((o & 4u) ? 0x0Fu : 0u); x ^= u | (u << 4);
    u = ((x >> 2) ^ x) & ((o & 2u) ? 0x33u : 0u); x ^= u | (u << 2);
    u = ((x >> 1) ^ x) & ((o & 1u) ? 0x55u : 0u); x ^= u | (u << 1);
    return x << 24;
}

// Runs until the query is finished (returns true) or fewer than `minActive` lanes of the warp are still
// traversing (returns false; call again later).  ANY = TerminateOnFirstHit.  Lanes postpone their remaining leaf
// triangles when fewer than 1/postponeDen of the lanes that entered the triangle loop are still in it (>= 2; 0 = never).
// (A variant that parked triangle groups and tested them in warp-voted steps to raise the SIMT efficiency of the
// triangle test was measured 17 % SLOWER on both workloads, profiles/r1e_sweep.txt: a closest hit that is found
// late stops culling the nodes behind it.  Triangles are therefore tested right after the node visit.)
#define OHB_POSTPONE_DEN_DEFAULT 5
template <bool ANY>
OHB_HD bool travRun(Trav& t, TravStackEntry* stack, const SceneDev& s, int minActive, int postponeDen) {
    for (;;) {
        if (t.G.y & 0xFF000000u) {
            // next inner child of the current group, front to back
            const uint32_t bit = bfind32(t.G.y);
            const uint32_t slot = (bit - 24u) ^ t.octinv;
            t.G.y &= ~(1u << bit);
            const uint32_t idx = t.G.x + popc32(t.G.y & 0xFFu & ((1u << slot) - 1u));
            if (t.G.y & 0xFF000000u) stack[t.sp++] = t.G;
            uint32_t childBase, triBase, imask, triMask;
            const uint32_t hits = intersectWideNode(t, s.wnodes + size_t(idx) * OHB_WNODE_VECS, ANY ? t.tmax : t.best.t, childBase, triBase, imask, triMask);
            t.G.x = childBase; t.G.y = hits | imask;
            t.Gt.x = triBase; t.Gt.y = triMask;
        } else {
            t.Gt = t.G; t.G.x = 0u; t.G.y = 0u;            // a postponed triangle group came off the stack
        }
        // leaf triangles; when most lanes of the warp have none left, the rest postpone theirs (Ylitie et al. §4.3)
        const int entered = postponeDen ? OHB_WARP_ACTIVE() : 0;
        while (t.Gt.y) {
            if (postponeDen && OHB_WARP_ACTIVE() * postponeDen < entered && t.sp < OHB_POSTPONE_SLOTS) { stack[t.sp++] = t.Gt; t.Gt.y = 0u; break; }
            const uint32_t k = bfind32(t.Gt.y);
            t.Gt.y &= ~(1u << k);
            const f4* tp = s.tris + size_t(t.Gt.x + k) * 3u;
            f4 v0 = ld4(tp), v1 = ld4(tp + 1), v2 = ld4(tp + 2);
            float tt, bu, bv;
            OHB_STAT_TRI();
            // tested against the caller's tmax (not best.t) so equal-t ties resolve by id
            if (!intersectTri(t.r, xyz(v0), xyz(v1), xyz(v2), t.tmax, tt, bu, bv)) continue;
            if (ANY) { t.anyHit = true; t.G.y = 0u; t.Gt.y = 0u; t.sp = 0; return true; }
            uint32_t id = f2u(v0.w);
            if (t.best.prim == OHB_MISS || tt < t.best.t || (tt == t.best.t && id < t.best.prim)) { t.best.t = tt; t.best.u = bu; t.best.v = bv; t.best.prim = id; }
        }
        if (!(t.G.y & 0xFF000000u)) {
            if (t.sp == 0) return true;
            t.G = stack[--t.sp];
        }
        if (OHB_WARP_ACTIVE() < minActive) return false;
    }
}

// Closest hit.  Returns prim == OHB_MISS and t = -1 on miss.
OHB_HD ohb_hit traceClosest(const SceneDev& s, f3 o, f3 d, float tmin, float tmax) {
    Trav t; TravStackEntry stack[OHB_STACK_SIZE]; travInit(t, s, o, d, tmin, tmax);
    if (t.G.y) while (!travRun<false>(t, stack, s, 0, OHB_POSTPONE_DEN_DEFAULT)) {}
    if (t.best.prim == OHB_MISS) t.best.t = -1.0f;
    return t.best;
}
// Any hit in (tmin, tmax): TerminateOnFirstHit | SkipClosestHit.
OHB_HD bool traceAny(const SceneDev& s, f3 o, f3 d, float tmin, float tmax) {
    Trav t; TravStackEntry stack[OHB_STACK_SIZE]; travInit(t, s, o, d, tmin, tmax);
    if (t.G.y) while (!travRun<true>(t, stack, s, 0, OHB_POSTPONE_DEN_DEFAULT)) {}
    return t.anyHit;
}

}  // namespace ohb

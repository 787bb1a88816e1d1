// ohb_traverse.h — 8-wide BVH traversal + watertight ray/triangle test (per-thread code).
//
// Replaces traceRayEXT (pt_raygen_offline.rgen:198-199 closest, :394-400 any-hit), whose
// arithmetic lives in the Vulkan driver.  The intersection arithmetic is the spec shared with the
// oracle (oracle/oracle_scene.h): Woop-Benthin-Wald watertight test in fp32 with a FIXED operation
// order (vertex - origin, shear by one IEEE fma, edge functions as two rounded products and one
// subtraction — never contracted, which is what keeps shared edges watertight —, fp64 fallback when
// an edge function is exactly zero), two-sided, accept tmin < t < tmax, equal-t ties resolved
// toward the lower global id.
//
// Everything here is shaped by what the r1 profiles and tools/ubench/pipes.cu measured on the B200:
// the traversal kernels are bound by the ALU pipe (LOP3/SHF/PRMT/SEL/FMNMX/FSETP issue at HALF rate, FFMA/
// FMUL/FADD at full rate, IMAD at half rate on its own pipe, POPC/FLO at 1/8), not by bytes.  So:
//   * node planes are bf16 OFFSETS from the node's fp32 min corner.  The odd child of a 32-bit word is used AS
//     the fp32 operand of the plane fma (the even child's bits below it are < 1 bf16 ulp of junk that the
//     builder folds into its outward rounding), the even child costs one shift — which the compiler places on
//     the IMAD pipe.  No byte->float conversion (r1: 48 PRMT + 48 register moves per visit), no per-node
//     exponent decode;
//   * the triangle record is stored transposed — (x0 x1 x2 id)(y0 y1 y2 id)(z0 z1 z2 id) — so the ray's axis
//     permutation (kx, ky, kz) is three row addresses instead of ~18 runtime selects per test;
//   * the hit is kept as (t, V, W, det, id); the two barycentric divisions happen once per ray.
//
// One node = 128 B = 8 x LDG.128 = exactly one L1 line:
//   w0 : p.x, p.y, p.z (fp32 min corner of the node box), { imask : 8 | 0 : 8 | ext : 16 }
//        bit s of imask = child slot s is an inner node; ext = bf16 (rounded up) of the node's largest extent
//   w1 : index of the first inner child, index of the first leaf triangle, meta[0..3], meta[4..7]
//        meta = 0 empty | 001_11sss inner child in slot s | uuu_ooooo leaf: unary triangle count, offset
//   w2, w3 : lo.x, hi.x   w4, w5 : lo.y, hi.y   w6, w7 : lo.z, hi.z — 8 bf16 each, word k = (slot 2k+1) << 16 | slot 2k
// fetched as 4 x LDG.E.256 (header, x planes, y planes, z planes); piece k of node i is stored at position k ^ (i & 3).
// Children sit in slots so that (slot ^ ray octant) is a front-to-back order; a node visit yields ONE
// bit mask of hit children (top byte: inner children in visit order, low 24 bits: leaf triangles), so the
// stack holds at most one entry per tree level (plus postponed triangle groups) and no distances
// (Ylitie, Karras, Laine 2017, "Efficient incoherent ray traversal on GPUs through compressed wide BVHs").
//
// Conservative boxes: the builder rounds the bf16 planes outward (checked in fp64, junk bits included) around
// boxes already padded by 2^-20 (ohb_bvh.h padBox), and every slab interval is widened by more than the
// rounding error of its plane distances plus the placement error of the triangle test (intersectWideNode).
// A box is therefore never culled wrongly and the hit (t, u, v, id) is independent of the tree: this 8-wide
// LBVH and the oracle's binary SAH tree give bit-identical results.
#pragma once
#include "ohb_scene.h"

namespace ohb {

#define OHB_STACK_SIZE 40       // >= OHB_MAX_LEVELS + OHB_POSTPONE_SLOTS
#define OHB_MAX_LEVELS 32       // deepest 8-wide tree the traversal stack is sized for (ohb_build_accel checks)
#define OHB_POSTPONE_SLOTS 8    // triangle groups may be postponed only while sp < this
#ifndef OHB_MAX_LEAF
#define OHB_MAX_LEAF 3          // triangles per leaf child (3-bit unary count)
#endif
#define OHB_WNODE_VECS 8        // 16-B words per node
#define OHB_EMPTY_LO 0x7F7Fu    // bf16 planes of an empty slot: lo = 3.39e38, hi = 0 -> never hit

struct alignas(8) u2 { uint32_t x, y; };

OHB_HD uint32_t popc32(uint32_t v) {
#if OHB_DEVICE_CODE
    return uint32_t(__popc(v));
#else
    return uint32_t(__builtin_popcount(v));
#endif
}
OHB_HD uint32_t bfind32(uint32_t v) { return 31u - uint32_t(clz32(v)); }   // v != 0
// (v << (s & 31)): SHF.L.W takes the low 5 bits of the shift register, so no masking instruction is needed
OHB_HD uint32_t shlWrap(uint32_t v, uint32_t s) {
#if OHB_DEVICE_CODE
    return __funnelshift_l(0u, v, s);
#else
    return v << (s & 31u);
#endif
}
// byte j (0..3) of v, zero-extended
OHB_HD uint32_t byteOf(uint32_t v, int j) {
#if OHB_DEVICE_CODE
    return __byte_perm(v, 0u, 0x4440u | uint32_t(j));
#else
    return (v >> (8 * j)) & 0xFFu;
#endif
}
OHB_HD float xfma(float a, float b, float c) {       // one IEEE rounding on both sides
#if OHB_DEVICE_CODE
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

// tests/emul counts node visits / triangle tests per query through these hooks; they compile to nothing in the product
#ifndef OHB_STAT_NODE
#define OHB_STAT_NODE()
#define OHB_STAT_TRI()
#endif

struct RayPrep {
    f3 o, idir; float tmin;
    f3 ok;                                   // origin in the permuted axes (kx, ky, kz)
    float Sx, Sy, Sz;
    uint32_t rowX, rowY, rowZ;               // rows (0..2) of the transposed triangle record that hold axes kx, ky, kz
};
// 256-bit load (LDG.E.256, sm_100+): two adjacent 16-B words of a node with ONE L1 tag lookup.  tools/ubench/l1.cu: a
// gather costs the L1 one clock per distinct 128-B line per instruction whatever its width (1.0 lane-loads/clk/SM from
// 4 to 32 B per lane), and the traversal kernels sit at 92 % of that limit — so bytes per load is what counts.
struct u8v { u4 lo, hi; };
OHB_HD u8v ldu8(const u4* p) {
    u8v r;
#if OHB_DEVICE_CODE
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w) : "l"(p));
#else
    r.lo = p[0]; r.hi = p[1];
#endif
    return r;
}
OHB_HD RayPrep prepRay(f3 o, f3 d, float tmin) {
    RayPrep r; r.o = o; r.tmin = tmin;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = 0; float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; m = az; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    if (comp(d, kz) < 0.0f) { int t = kx; kx = ky; ky = t; }
    r.rowX = uint32_t(kx); r.rowY = uint32_t(ky); r.rowZ = uint32_t(kz);
    r.ok = mk3(comp(o, kx), comp(o, ky), comp(o, kz));
    float dz = comp(d, kz);
    r.Sx = xdiv(comp(d, kx), dz); r.Sy = xdiv(comp(d, ky), dz); r.Sz = xdiv(1.0f, dz);
    float sx = fabsf(d.x) > 1e-20f ? d.x : copysignf(1e-20f, d.x);
    float sy = fabsf(d.y) > 1e-20f ? d.y : copysignf(1e-20f, d.y);
    float sz = fabsf(d.z) > 1e-20f ? d.z : copysignf(1e-20f, d.z);
    r.idir = mk3(1.0f / sx, 1.0f / sy, 1.0f / sz);
    return r;
}

// Candidate of the watertight test: t = T / det, u = V / det, v = W / det (divisions deferred, same values).
struct TriHit { float t, V, W, det; };
struct BestHit { float V, W, det; };
// fp64 re-evaluation of the three edge functions (taken only when one is exactly zero in fp32); out of line to keep
// the DMUL/DFMA/F2F block out of the hot loop's instruction stream
#if defined(__CUDACC__)
static __device__ __host__ __noinline__
#else
static inline
#endif
f3 edgeFallback(float Ax, float Ay, float Bx, float By, float Cx, float Cy) {     // by value: reference parameters put U, V, W into local memory on EVERY test
    f3 r;
    r.x = float(double(Cx) * double(By) - double(Cy) * double(Bx));
    r.y = float(double(Ax) * double(Cy) - double(Ay) * double(Cx));
    r.z = float(double(Bx) * double(Ay) - double(By) * double(Ax));
    return r;
}
// tp = the 3 rows of one triangle.  Returns true with the candidate when tmin < t <= tlim (the caller rejects
// t == tlim where that is not a tie to break); id = global triangle id.
struct TriRows { f4 rx, ry, rz; };
OHB_HD TriRows loadTri(const RayPrep& r, const f4* tp) { TriRows q; q.rx = ld4(tp + r.rowX); q.ry = ld4(tp + r.rowY); q.rz = ld4(tp + r.rowZ); return q; }
OHB_HD bool intersectTri(const RayPrep& r, const TriRows& q, float tlim, TriHit& h, uint32_t& id) {
    const f4 rx = q.rx, ry = q.ry, rz = q.rz;
    OHB_STAT_TRI();
    id = f2u(rz.w);
    const float Akx = xsub(rx.x, r.ok.x), Bkx = xsub(rx.y, r.ok.x), Ckx = xsub(rx.z, r.ok.x);
    const float Aky = xsub(ry.x, r.ok.y), Bky = xsub(ry.y, r.ok.y), Cky = xsub(ry.z, r.ok.y);
    const float Akz = xsub(rz.x, r.ok.z), Bkz = xsub(rz.y, r.ok.z), Ckz = xsub(rz.z, r.ok.z);
    const float Ax = xfma(-r.Sx, Akz, Akx), Ay = xfma(-r.Sy, Akz, Aky);
    const float Bx = xfma(-r.Sx, Bkz, Bkx), By = xfma(-r.Sy, Bkz, Bky);
    const float Cx = xfma(-r.Sx, Ckz, Ckx), Cy = xfma(-r.Sy, Ckz, Cky);
    float U = xsub(xmul(Cx, By), xmul(Cy, Bx));
    float V = xsub(xmul(Ax, Cy), xmul(Ay, Cx));
    float W = xsub(xmul(Bx, Ay), xmul(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) { const f3 e = edgeFallback(Ax, Ay, Bx, By, Cx, Cy); U = e.x; V = e.y; W = e.z; }
    if (fminf(fminf(U, V), W) < 0.0f && fmaxf(fmaxf(U, V), W) > 0.0f) return false;
    const float det = xadd(xadd(U, V), W);
    if (det == 0.0f) return false;
    const float T = xmul(xfma(W, Ckz, xfma(V, Bkz, xmul(U, Akz))), r.Sz);
    const float tt = xdiv(T, det);
    if (!(tt > r.tmin && tt <= tlim)) return false;
    h.t = tt; h.V = V; h.W = W; h.det = det;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Resumable traversal.  State = current node group G (x: index of the first inner child, y: hit bits of
// the inner children in the top byte | imask), current triangle group Gt (x: first triangle, y: bits),
// a stack of groups.  It survives a pause, which lets the persistent kernels refill idle lanes with new
// rays when too few lanes of a warp are still traversing.
// ---------------------------------------------------------------------------------------------
#if OHB_DEVICE_CODE
#define OHB_WARP_ACTIVE() (__popc(__activemask()))
#elif !defined(OHB_WARP_ACTIVE)
#define OHB_WARP_ACTIVE() 32       // host build (tests/emul may supply its own stand-in)
#endif

// tlim: closest-hit queries — the caller's tmax until a hit is found, then the closest t so far; any-hit queries — tmax.
// A closest-hit candidate at exactly tlim is a tie (resolved toward the lower id) once a hit exists, and is outside
// the open interval (tmin, tmax) before: one register serves as both bounds.
struct Trav {
    RayPrep r; float tlim; BestHit best; uint32_t prim;
    u2 G, Gt; int sp; uint32_t octinv; bool anyHit;
    // two-level traversal only: the world-space ray while inside an instance, and the tree the current groups refer to
    f3 wo, wd; const u4* curNodes; const f4* curTris; bool inBlas;
};
typedef u2 TravStackEntry;
// Traversal stack.  ArrayStack: a per-thread array (local memory on the device).  SharedStack<SH, STRIDE>: the first SH
// entries of every thread live in shared memory (entry k of thread t at sh[k * STRIDE + t]: a warp's accesses to one
// level are 32 consecutive 8-B words, conflict-free), deeper entries in local memory.  The 8-wide tree keeps at most
// one entry per level, so 8 shared entries hold the whole stack of almost every ray of a 2 M-triangle scene.
struct ArrayStack {
    TravStackEntry* p;
    OHB_HD void push(int sp, TravStackEntry v) { p[sp] = v; }
    OHB_HD TravStackEntry pop(int sp) const { return p[sp]; }
};
template <int SH, int STRIDE>
struct SharedStack {
    TravStackEntry* sh; TravStackEntry loc[OHB_STACK_SIZE - SH];
    OHB_HD void push(int sp, TravStackEntry v) { if (sp < SH) sh[sp * STRIDE] = v; else loc[sp - SH] = v; }
    OHB_HD TravStackEntry pop(int sp) const { return sp < SH ? sh[sp * STRIDE] : loc[sp - SH]; }
};
// Top of the tree staged in shared memory (north_star: "shared-memory staging of the top tree levels"): the builder
// allocates wide nodes level by level, so nodes [0, count) are the top levels.  top == nullptr: everything from global.
struct TopNodes { const u4* top; uint32_t count; };

OHB_HD uint32_t octantOf(const RayPrep& r) { return (r.idir.x < 0.0f ? 0u : 4u) | (r.idir.y < 0.0f ? 0u : 2u) | (r.idir.z < 0.0f ? 0u : 1u); }
OHB_HD void travInit(Trav& t, const SceneDev& s, f3 o, f3 d, float tmin, float tmax) {
    t.r = prepRay(o, d, tmin); t.tlim = tmax;
    t.best.V = 0.0f; t.best.W = 0.0f; t.best.det = 1.0f; t.prim = OHB_MISS;
    t.sp = 0; t.anyHit = false;
    // bit 2/1/0 set = the ray travels toward +x/+y/+z: slot ^ octinv is then the visit priority (7 first)
    t.octinv = octantOf(t.r);
    t.G.x = 0u; t.G.y = s.numTris ? 0x80000000u : 0u;      // the root is "slot 7 ^ octinv" of a virtual group
    t.Gt.x = 0u; t.Gt.y = 0u;
    t.wo = o; t.wd = d; t.inBlas = false;
    t.curNodes = s.twoLevel ? s.tlasNodes : s.wnodes; t.curTris = s.twoLevel ? s.tlasLeaves : s.tris;
}
// ---- two-level traversal (RTAccelerationStructure's BLAS/TLAS, rt_acceleration_structure.cpp:205-535) ----------------------
// A TLAS leaf entry is an instance.  Entering it parks what is left of the TLAS-level groups on the stack under a sentinel,
// maps the ray into the instance's object space — o' = M^-1 o, d' = M^-1 d, NOT renormalised, so t is the same parameter in
// both spaces and the closest-hit bound carries over — and continues in the instance's BLAS; popping the sentinel restores
// the world-space ray.  The transform arithmetic is part of the spec shared with the oracle: ((m0 x + m1 y) + m2 z) + m3,
// one rounding per operation (xformPoint of the builder).
#define OHB_TL_SENTINEL 0xFFFFFFFFu
OHB_HD void travEnterInstance(Trav& t, const SceneDev& s, uint32_t inst) {
    const f4* iv = s.instInv + size_t(inst) * 3u;
    const f4 r0 = iv[0], r1 = iv[1], r2 = iv[2];
    const f3 o = t.wo, d = t.wd;
    const f3 oo = mk3(xadd(xadd(xadd(xmul(r0.x, o.x), xmul(r0.y, o.y)), xmul(r0.z, o.z)), r0.w),
                      xadd(xadd(xadd(xmul(r1.x, o.x), xmul(r1.y, o.y)), xmul(r1.z, o.z)), r1.w),
                      xadd(xadd(xadd(xmul(r2.x, o.x), xmul(r2.y, o.y)), xmul(r2.z, o.z)), r2.w));
    const f3 od = mk3(xadd(xadd(xmul(r0.x, d.x), xmul(r0.y, d.y)), xmul(r0.z, d.z)),
                      xadd(xadd(xmul(r1.x, d.x), xmul(r1.y, d.y)), xmul(r1.z, d.z)),
                      xadd(xadd(xmul(r2.x, d.x), xmul(r2.y, d.y)), xmul(r2.z, d.z)));
    const float tmin = t.r.tmin;
    t.r = prepRay(oo, od, tmin); t.octinv = octantOf(t.r);
    const u4 bi = s.blasInfo[inst];
    t.curNodes = s.wnodes + size_t(bi.x) * OHB_WNODE_VECS; t.curTris = s.tris + size_t(bi.y) * 3u;
    t.inBlas = true;
    t.G.x = 0u; t.G.y = bi.z ? 0x80000000u : 0u; t.Gt.x = 0u; t.Gt.y = 0u;
}
OHB_HD void travLeaveInstance(Trav& t, const SceneDev& s) {
    const float tmin = t.r.tmin;
    t.r = prepRay(t.wo, t.wd, tmin); t.octinv = octantOf(t.r);
    t.curNodes = s.tlasNodes; t.curTris = s.tlasLeaves; t.inBlas = false;
}
// the result of a finished closest-hit query in ABI form (prim == OHB_MISS and t = -1 on miss)
OHB_HD ohb_hit travResult(const Trav& t) {
    ohb_hit h;
    if (t.prim == OHB_MISS) { h.t = -1.0f; h.u = 0.0f; h.v = 0.0f; h.prim = OHB_MISS; return h; }
    h.t = t.tlim; h.u = xdiv(t.best.V, t.best.det); h.v = xdiv(t.best.W, t.best.det); h.prim = t.prim;
    return h;
}

// One node visit: tests the 8 child boxes against [tmin, thi].  Returns the hit children as ONE mask: inner
// children in bits 24..31 at bit 24 + (slot ^ octinv) (= visit priority), leaf triangles in bits 0..23
// (bit k = triangle triBase + k).
// Plane distance t = fma(v, 1/d, (p - o) * (1/d)) with v = the plane's bf16 offset from p.  Arithmetic error per
// plane: 6e-8|p - o| in the difference, 6e-8|a| in the product, 6e-8 relative in 1/d (both terms), 6e-8|t| in the
// fma  <  2.5e-7 (|a| + |v / d|).
// The watertight test itself places a hit only to within a few ulp of the DISTANCE to the triangle's vertices
// (its t is a barycentric mean of vertex depths whose weights carry the cancellation error of the sheared
// coordinates), on every axis — a flat, axis-aligned triangle through the origin has a zero-width slab that
// this error exceeds.  Each slab is therefore widened by 1.2e-6 x the largest per-axis distance from the ray
// origin to the far side of the node (|p - o| + ext on the worst axis), converted to ray-parameter units: that
// covers the 5e-7 of arithmetic error above and leaves 7e-7 (a dozen ulp) for the placement error.
#define OHB_NODE_SWZ(i) ((i) & 3u)
OHB_HD u8v ldsu8(const u4* p) { u8v r; r.lo = p[0]; r.hi = p[1]; return r; }
// ORDERED = false (any-hit queries): the inner children keep their slot order — an occlusion test does not care which occluder it
// finds first, and the octant permutation of the meta bytes is 12 instructions per visit
template <bool ORDERED = true>
OHB_HD uint32_t intersectWideNode(const Trav& t, const u4* gnodes, const TopNodes& tn8, uint32_t idx, float thi, uint32_t& childBase, uint32_t& triBase, uint32_t& imask) {
    const RayPrep& r = t.r;
    const uint32_t swz = OHB_NODE_SWZ(idx);
    // the four 32-B pieces of node i sit at positions k ^ (i & 3) of its 128-B line (OHB_NODE_SWZ): all lanes of a warp
    // read "piece k of my node" in the same instruction, and with every node laid out alike they would all hit the same
    // 8 of the 32 L1 data banks — measured 1.0 lane-loads/clk/SM against 1.8-2.9 with rotated offsets (tools/ubench/l1.cu)
    u8v hdr, px8, py8, pz8;
    if (tn8.top && idx < tn8.count) {
        const u4* np = tn8.top + size_t(idx) * OHB_WNODE_VECS;
        hdr = ldsu8(np + 2u * (0u ^ swz)); px8 = ldsu8(np + 2u * (1u ^ swz)); py8 = ldsu8(np + 2u * (2u ^ swz)); pz8 = ldsu8(np + 2u * (3u ^ swz));
    } else {
#if OHB_DEVICE_CODE
        // the node array is 128-B aligned (cudaMalloc; BLAS sub-arrays start at whole nodes): the piece offset (k ^ swz) << 5 is
        // OR-ed into the low address word — one LOP3 per piece instead of a 64-bit add with carry
        const uintptr_t nb = reinterpret_cast<uintptr_t>(gnodes + size_t(idx) * OHB_WNODE_VECS);
        hdr = ldu8(reinterpret_cast<const u4*>(nb | uintptr_t((0u ^ swz) << 5))); px8 = ldu8(reinterpret_cast<const u4*>(nb | uintptr_t((1u ^ swz) << 5)));
        py8 = ldu8(reinterpret_cast<const u4*>(nb | uintptr_t((2u ^ swz) << 5))); pz8 = ldu8(reinterpret_cast<const u4*>(nb | uintptr_t((3u ^ swz) << 5)));
#else
        const u4* np = gnodes + size_t(idx) * OHB_WNODE_VECS;
        hdr = ldu8(np + 2u * (0u ^ swz)); px8 = ldu8(np + 2u * (1u ^ swz)); py8 = ldu8(np + 2u * (2u ^ swz)); pz8 = ldu8(np + 2u * (3u ^ swz));
#endif
    }
    OHB_STAT_NODE();
    const u4 w0 = hdr.lo, w1 = hdr.hi;
    // near / far planes by ray sign: 24 word selects (the 16-B vectors would be selectable by address, but only as
    // six 128-bit loads — two more L1 lookups per visit than the selects cost in ALU time)
    const bool nx = r.idir.x < 0.0f, ny = r.idir.y < 0.0f, nz = r.idir.z < 0.0f;
    u4 vnx, vfx, vny, vfy, vnz, vfz;
    vnx.x = nx ? px8.hi.x : px8.lo.x; vnx.y = nx ? px8.hi.y : px8.lo.y; vnx.z = nx ? px8.hi.z : px8.lo.z; vnx.w = nx ? px8.hi.w : px8.lo.w;
    vfx.x = nx ? px8.lo.x : px8.hi.x; vfx.y = nx ? px8.lo.y : px8.hi.y; vfx.z = nx ? px8.lo.z : px8.hi.z; vfx.w = nx ? px8.lo.w : px8.hi.w;
    vny.x = ny ? py8.hi.x : py8.lo.x; vny.y = ny ? py8.hi.y : py8.lo.y; vny.z = ny ? py8.hi.z : py8.lo.z; vny.w = ny ? py8.hi.w : py8.lo.w;
    vfy.x = ny ? py8.lo.x : py8.hi.x; vfy.y = ny ? py8.lo.y : py8.hi.y; vfy.z = ny ? py8.lo.z : py8.hi.z; vfy.w = ny ? py8.lo.w : py8.hi.w;
    vnz.x = nz ? pz8.hi.x : pz8.lo.x; vnz.y = nz ? pz8.hi.y : pz8.lo.y; vnz.z = nz ? pz8.hi.z : pz8.lo.z; vnz.w = nz ? pz8.hi.w : pz8.lo.w;
    vfz.x = nz ? pz8.lo.x : pz8.hi.x; vfz.y = nz ? pz8.lo.y : pz8.hi.y; vfz.z = nz ? pz8.lo.z : pz8.hi.z; vfz.w = nz ? pz8.lo.w : pz8.hi.w;
    childBase = w1.x; triBase = w1.y; imask = w0.w & 0xFFu;
    const float px = u2f(w0.x) - r.o.x, py = u2f(w0.y) - r.o.y, pz = u2f(w0.z) - r.o.z;
    const float ext = u2f(w0.w & 0xFFFF0000u);
    const float dist = 1.2e-6f * fmaxf(fmaxf(fabsf(px) + ext, fabsf(py) + ext), fabsf(pz) + ext);
    const float ex = dist * fabsf(r.idir.x), ey = dist * fabsf(r.idir.y), ez = dist * fabsf(r.idir.z);
    const float ax = px * r.idir.x, ay = py * r.idir.y, az = pz * r.idir.z;
    const float axn = ax - ex, axf = ax + ex, ayn = ay - ey, ayf = ay + ey, azn = az - ez, azf = az + ez;
    uint32_t hits = 0u;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int g = 0; g < 2; g++) {
        // meta bytes: inner children get their slot index XORed with the ray octant (= visit priority)
        uint32_t m4 = g ? w1.w : w1.z;
        const uint32_t inner4 = ORDERED ? (m4 & (m4 << 1) & 0x10101010u) : 0u;  // 0x10 in the bytes of inner children
        if (ORDERED) m4 ^= (inner4 >> 4) * t.octinv;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int j = 0; j < 4; j++) {
            const int k = 2 * g + (j >> 1);                                    // word that holds slot 4g + j
            const uint32_t wnx = k == 0 ? vnx.x : (k == 1 ? vnx.y : (k == 2 ? vnx.z : vnx.w));
            const uint32_t wny = k == 0 ? vny.x : (k == 1 ? vny.y : (k == 2 ? vny.z : vny.w));
            const uint32_t wnz = k == 0 ? vnz.x : (k == 1 ? vnz.y : (k == 2 ? vnz.z : vnz.w));
            const uint32_t wfx = k == 0 ? vfx.x : (k == 1 ? vfx.y : (k == 2 ? vfx.z : vfx.w));
            const uint32_t wfy = k == 0 ? vfy.x : (k == 1 ? vfy.y : (k == 2 ? vfy.z : vfy.w));
            const uint32_t wfz = k == 0 ? vfz.x : (k == 1 ? vfz.y : (k == 2 ? vfz.z : vfz.w));
            const bool odd = j & 1;                                            // odd slot: the word itself is the operand
            const float tnx = fmaf(u2f(odd ? wnx : wnx << 16), r.idir.x, axn), tfx = fmaf(u2f(odd ? wfx : wfx << 16), r.idir.x, axf);
            const float tny = fmaf(u2f(odd ? wny : wny << 16), r.idir.y, ayn), tfy = fmaf(u2f(odd ? wfy : wfy << 16), r.idir.y, ayf);
            const float tnz = fmaf(u2f(odd ? wnz : wnz << 16), r.idir.z, azn), tfz = fmaf(u2f(odd ? wfz : wfz << 16), r.idir.z, azf);
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, r.tmin));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, thi));
            // <= (not <): a box touching at exactly best.t may hold an equal-t, lower-id triangle
            const uint32_t b = byteOf(m4, j);
            if (tn <= tf) hits |= shlWrap(b >> 5, b);
        }
    }
    return hits;
}

// Runs until the query is finished (returns true) or fewer than `minActive` lanes of the warp are still
// traversing (returns false; call again later).  ANY = TerminateOnFirstHit.  Lanes postpone their remaining leaf
// triangles when fewer than 1/postponeDen of the lanes that entered the triangle loop are still in it (>= 2; 0 = never).
// (A variant that parked triangle groups and tested them in warp-voted steps to raise the SIMT efficiency of the
// triangle test was measured 17 % SLOWER on both workloads, profiles/r1e_sweep.txt: a closest hit that is found
// late stops culling the nodes behind it.  Triangles are therefore tested right after the node visit.)
#define OHB_POSTPONE_DEN_DEFAULT 5
#ifndef OHB_ANY_UNORDERED
#define OHB_ANY_UNORDERED 1      // any-hit queries visit inner children in slot order (no octant permutation)
#endif
#ifndef OHB_TRI_UNROLL
#define OHB_TRI_UNROLL 1          // 0 = never, 1 = any-hit kernels only, 2 = always
#endif
#ifndef OHB_TRAV_PREFETCH
#define OHB_TRAV_PREFETCH 0          // 1 = keep the runtime-selected node prefetch variants (OHB_PREFETCH knob)
#endif
// software prefetch of a node's 128-B line (no register, no scoreboard): mode bit 0 = into L2, bit 1 = into L1, bit 4 = one
// prefetch per 32-B piece instead of one per line
OHB_HD void prefetchNode(const u4* np, int mode) {
#if OHB_DEVICE_CODE
    const int n = (mode & 16) ? 4 : 1;
    for (int k = 0; k < n; k++) {
        if (mode & 2) asm volatile("prefetch.global.L1 [%0];" :: "l"(np + 2 * k));
        else asm volatile("prefetch.global.L2 [%0];" :: "l"(np + 2 * k));
    }
#endif
}
// index of the child a group would hand out next (the same arithmetic as the pop in travRun)
OHB_HD uint32_t nextChildOf(const u2& G, uint32_t octinv) {
    const uint32_t bit = bfind32(G.y);
    const uint32_t slot = (bit - 24u) ^ octinv;
    return G.x + popc32(G.y & 0xFFu & ((1u << slot) - 1u));
}
// postponeDen carries the prefetch mode in bits 8.. : bits 8-9 = prefetch the NEXT SIBLING when a group with children left
// is pushed (the node a later stack pop will want; 1 = L2, 2 = L1), bits 10-11 = prefetch the FIRST HIT CHILD right after
// the box tests, before the leaf triangles of this node are tested, bit 12 = four prefetches per node
template <bool ANY, class Stack, bool TL = false>
OHB_HD bool travRun(Trav& t, Stack& stack, const SceneDev& s, const TopNodes& top, int minActive, int postponeDenAndPf) {
#if OHB_TRAV_PREFETCH
    const int postponeDen = postponeDenAndPf & 0xFF, pfSib = (postponeDenAndPf >> 8) & 3, pfKid = (postponeDenAndPf >> 10) & 3, pf4 = (postponeDenAndPf >> 8) & 16;
#else
    // the prefetch variants lost 2-12 % (profiles/r2h_sweep_prefetch.txt): compiled out, their per-iteration mode tests with them
    const int postponeDen = postponeDenAndPf & 0xFF; constexpr int pfSib = 0, pfKid = 0, pf4 = 0;
#endif
    for (;;) {
        if (t.G.y & 0xFF000000u) {
            // next inner child of the current group, front to back
            const uint32_t bit = bfind32(t.G.y);
            const uint32_t slot = (OHB_ANY_UNORDERED && ANY) ? (bit - 24u) : ((bit - 24u) ^ t.octinv);
            t.G.y &= ~(1u << bit);
            const uint32_t idx = t.G.x + popc32(t.G.y & 0xFFu & ((1u << slot) - 1u));
            if (t.G.y & 0xFF000000u) {
                stack.push(t.sp++, t.G);
                if (pfSib) prefetchNode(s.wnodes + size_t(nextChildOf(t.G, t.octinv)) * OHB_WNODE_VECS, pfSib | pf4);
            }
            uint32_t childBase, triBase, imask;
            const uint32_t hits = intersectWideNode<!(OHB_ANY_UNORDERED && ANY)>(t, TL ? t.curNodes : s.wnodes, top, idx, t.tlim, childBase, triBase, imask);
            t.G.x = childBase; t.G.y = (hits & 0xFF000000u) | imask;
            t.Gt.x = triBase; t.Gt.y = hits & 0x00FFFFFFu;
            if (pfKid && (t.G.y & 0xFF000000u) && t.Gt.y) prefetchNode(s.wnodes + size_t(nextChildOf(t.G, t.octinv)) * OHB_WNODE_VECS, pfKid | pf4);
        } else {
            t.Gt = t.G; t.G.x = 0u; t.G.y = 0u;            // a postponed triangle group came off the stack
        }
        // leaf triangles; when most lanes of the warp have none left, the rest postpone theirs (Ylitie et al. §4.3)
        if (TL && !t.inBlas && t.Gt.y) {
            // TLAS level: the group's entries are instances.  Take the first one, park the rest, descend into its BLAS.
            const uint32_t k = bfind32(t.Gt.y);
            t.Gt.y &= ~(1u << k);
            const uint32_t inst = f2u(ld4(t.curTris + size_t(t.Gt.x + k) * 3u).w);
            if (t.G.y & 0xFF000000u) stack.push(t.sp++, t.G);
            if (t.Gt.y) stack.push(t.sp++, t.Gt);
            TravStackEntry sentinel; sentinel.x = OHB_TL_SENTINEL; sentinel.y = 0u; stack.push(t.sp++, sentinel);
            travEnterInstance(t, s, inst);
            continue;
        }
        const int entered = postponeDen ? OHB_WARP_ACTIVE() : 0;
        while (t.Gt.y) {
            if (postponeDen && OHB_WARP_ACTIVE() * postponeDen < entered && t.sp < OHB_POSTPONE_SLOTS) { stack.push(t.sp++, t.Gt); t.Gt.y = 0u; break; }
            // TWO triangles per round, both fetched before either is tested: the kernels wait on memory, not on issue
            // slots (r2j: 22 % of all samples sat on the triangle loads), and the register file is half empty here —
            // the node phase, not this loop, sets the kernel's register count
            const uint32_t k0 = bfind32(t.Gt.y);
            t.Gt.y &= ~(1u << k0);
            const bool two = t.Gt.y != 0u;
            const uint32_t k1 = two ? bfind32(t.Gt.y) : k0;
            t.Gt.y &= ~(1u << k1);
            const f4* triBasePtr = TL ? t.curTris : s.tris;
            const TriRows q0 = loadTri(t.r, triBasePtr + size_t(t.Gt.x + k0) * 3u);
            const TriRows q1 = loadTri(t.r, triBasePtr + size_t(t.Gt.x + k1) * 3u);
            // Any-hit queries run two copies of the test body; closest-hit queries one body that selects its triangle's twelve
            // values per pass (12 FSEL + the loop's predicate bookkeeping).  Measured (profiles/r2ar_sweep_tri_test.txt): unrolling
            // takes 1.5-2 % off k_trace_shadow on both scenes, but adds 0.7 % to k_trace_closest on the 2 M scene (-0.9 % on helmet).
#ifdef __CUDA_ARCH__
            constexpr int triUnroll = ((OHB_TRI_UNROLL == 2) || (OHB_TRI_UNROLL == 1 && ANY)) ? 2 : 1;
#pragma unroll (triUnroll)
#endif
            for (int j = 0; j < 2; j++) {
                if (j == 1 && !two) break;
                TriHit h; uint32_t id;
                if (!intersectTri(t.r, j ? q1 : q0, t.tlim, h, id)) continue;
                if (ANY) {
                    if (h.t == t.tlim) continue;                                  // open interval
                    t.anyHit = true; t.G.y = 0u; t.Gt.y = 0u; t.sp = 0; return true;
                }
                // t == tlim: a tie with the current hit (lower id wins), or t == tmax with no hit yet (outside the interval)
                if (h.t < t.tlim || (t.prim != OHB_MISS && id < t.prim)) { t.tlim = h.t; t.best.V = h.V; t.best.W = h.W; t.best.det = h.det; t.prim = id; }
            }
        }
        if (!(t.G.y & 0xFF000000u)) {
            if (t.sp == 0) return true;
            t.G = stack.pop(--t.sp);
            if (TL) {
                while (t.G.x == OHB_TL_SENTINEL && t.G.y == 0u) {        // the instance's BLAS is exhausted: back to the TLAS level
                    travLeaveInstance(t, s);
                    if (t.sp == 0) return true;
                    t.G = stack.pop(--t.sp);
                }
            }
        }
        if (OHB_WARP_ACTIVE() < minActive) return false;
    }
}

// Closest hit.  Returns prim == OHB_MISS and t = -1 on miss.
OHB_HD ohb_hit traceClosest(const SceneDev& s, f3 o, f3 d, float tmin, float tmax) {
    Trav t; TravStackEntry mem[OHB_STACK_SIZE]; ArrayStack stack{mem}; const TopNodes top{nullptr, 0u}; travInit(t, s, o, d, tmin, tmax);
    if (t.G.y) {
        if (s.twoLevel) { while (!travRun<false, ArrayStack, true>(t, stack, s, top, 0, OHB_POSTPONE_DEN_DEFAULT)) {} }
        else { while (!travRun<false>(t, stack, s, top, 0, OHB_POSTPONE_DEN_DEFAULT)) {} }
    }
    return travResult(t);
}
// Any hit in (tmin, tmax): TerminateOnFirstHit | SkipClosestHit.
OHB_HD bool traceAny(const SceneDev& s, f3 o, f3 d, float tmin, float tmax) {
    Trav t; TravStackEntry mem[OHB_STACK_SIZE]; ArrayStack stack{mem}; const TopNodes top{nullptr, 0u}; travInit(t, s, o, d, tmin, tmax);
    if (t.G.y) {
        if (s.twoLevel) { while (!travRun<true, ArrayStack, true>(t, stack, s, top, 0, OHB_POSTPONE_DEN_DEFAULT)) {} }
        else { while (!travRun<true>(t, stack, s, top, 0, OHB_POSTPONE_DEN_DEFAULT)) {} }
    }
    return t.anyHit;
}

}  // namespace ohb

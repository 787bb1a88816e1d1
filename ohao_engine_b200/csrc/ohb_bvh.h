// ohb_bvh.h — LBVH construction, per-thread stages (Karras 2012 "Maximizing parallelism in the
// construction of BVHs, octrees and k-d trees") + SAH treelets + collapse into the compressed 8-wide BVH.
//
// Replaces RTAccelerationStructure::{createBLAS, buildTLAS} (rt_acceleration_structure.cpp:205-535),
// i.e. vkCmdBuildAccelerationStructuresKHR, whose algorithm is inside the Vulkan driver.  The
// engine's two-level interface is kept at the API (ohb_set_instances); all reference meshes are
// static (rt_build.cpp:853,865), so instances are flattened into ONE world-space BVH — a TLAS
// "build" is a re-transform + rebuild, there is no per-ray instance transform.
//
// Pipeline (one kernel each, ohb_kernels.cu):
//   worldTriangles -> sceneBounds -> morton63 -> radix sort (key u64, val u32) -> karras hierarchy
//   -> bottom-up sweep (atomic visit counters): AABB refit + triangle counts + SAH cost
//   -> 3 passes of SAH treelet restructuring (7-leaf treelets, gamma = 7, 14, 28; Karras & Aila 2013)
//   -> top-down collapse, one launch per level: greedy 8-wide nodes (subtrees of <= 3 triangles become leaf
//      children), octant slot assignment, bf16 outward-rounded plane offsets -> 128-B nodes + 48-B transposed
//      triangles, every subtree's triangles contiguous.
#pragma once
#include "ohb_traverse.h"

namespace ohb {

struct BuildArrays {
    // inputs
    const uint8_t* positions; uint64_t posStride;     // object-space positions, strided
    const uint32_t* indices;                          // global
    const uint32_t* triInst;                          // per global tri
    const f4* instXform;                              // 3 rows per instance (object->world)
    const uint32_t* activeTris; uint32_t n;           // triangles to build over
    uint32_t objectSpace;                             // 1: no instance transform (a BLAS of the two-level structure is built in object space)
    // intermediates (length n unless noted)
    f4* wtri;            // 3n : world-space triangles in INPUT order, v0.w = global id
    f4* primLo; f4* primHi;
    uint32_t* boundsBits;   // 6 : ordered-uint scene bounds (min xyz, max xyz)
    uint64_t* keys; uint32_t* vals;          // morton + index (sorted in place by the radix sort)
    // hierarchy over n-1 internal nodes
    int32_t* left; int32_t* right;           // child: >=0 internal, <0 => ~leafIndex (sorted position)
    int32_t* parentInner; int32_t* parentLeaf;
    f4* nodeLo; f4* nodeHi;                  // internal nodes: (lo, triangle count bits) (hi, SAH cost of the subtree)
    uint32_t* visit;                         // n-1 atomic flags
    uint32_t* wideCounters;                  // 4 : {wide nodes allocated, items of level A, items of level B, -}
    float* sah;                              // 2 : {wide-node area sum, leaf area*count sum}
    // outputs
    u4* wnodes; f4* tris;                    // 8 x u4 per wide node, 3 x f4 (transposed) per triangle (ohb_traverse.h)
};

OHB_HD uint32_t floatOrdered(float f) { uint32_t u = f2u(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
OHB_HD float orderedFloat(uint32_t u) { return u2f((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

OHB_HD void atomicMinU32(uint32_t* p, uint32_t v) {
#if OHB_DEVICE_CODE
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
OHB_HD void atomicMaxU32(uint32_t* p, uint32_t v) {
#if OHB_DEVICE_CODE
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
OHB_HD uint32_t atomicIncU32(uint32_t* p) {
#if OHB_DEVICE_CODE
    return atomicAdd(p, 1u);
#else
    return (*p)++;
#endif
}
OHB_HD uint32_t atomicIncU32N(uint32_t* p, uint32_t n) {
#if OHB_DEVICE_CODE
    return atomicAdd(p, n);
#else
    uint32_t r = *p; *p += n; return r;
#endif
}
OHB_HD void atomicAddF32(float* p, float v) {
#if OHB_DEVICE_CODE
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
OHB_HD void fenceDevice() {
#if OHB_DEVICE_CODE
    __threadfence();
#endif
}

// Stage 1: object -> world (spec arithmetic: ((m0*x + m1*y) + m2*z) + m3, one rounding per op).
OHB_HD f3 xformPoint(const f4* rows, f3 p) {
    f4 r0 = rows[0], r1 = rows[1], r2 = rows[2];
    return mk3(xadd(xadd(xadd(xmul(r0.x, p.x), xmul(r0.y, p.y)), xmul(r0.z, p.z)), r0.w),
               xadd(xadd(xadd(xmul(r1.x, p.x), xmul(r1.y, p.y)), xmul(r1.z, p.z)), r1.w),
               xadd(xadd(xadd(xmul(r2.x, p.x), xmul(r2.y, p.y)), xmul(r2.z, p.z)), r2.w));
}
OHB_HD void buildWorldTri(const BuildArrays& b, uint32_t a) {
    uint32_t g = b.activeTris[a];
    const f4* rows = b.instXform + size_t(b.triInst[g]) * 3u;
    f3 v[3];
    for (int k = 0; k < 3; k++) {
        const float* pp = reinterpret_cast<const float*>(b.positions + size_t(b.indices[size_t(g) * 3 + k]) * b.posStride);
        v[k] = b.objectSpace ? mk3(pp[0], pp[1], pp[2]) : xformPoint(rows, mk3(pp[0], pp[1], pp[2]));
    }
    b.wtri[size_t(a) * 3 + 0] = mk4(v[0], u2f(g));
    b.wtri[size_t(a) * 3 + 1] = mk4(v[1], 0.0f);
    b.wtri[size_t(a) * 3 + 2] = mk4(v[2], 0.0f);
    f3 lo = vmin(v[0], vmin(v[1], v[2])), hi = vmax(v[0], vmax(v[1], v[2]));
    b.primLo[a] = mk4(lo, 0.0f); b.primHi[a] = mk4(hi, 0.0f);
    atomicMinU32(b.boundsBits + 0, floatOrdered(lo.x)); atomicMinU32(b.boundsBits + 1, floatOrdered(lo.y)); atomicMinU32(b.boundsBits + 2, floatOrdered(lo.z));
    atomicMaxU32(b.boundsBits + 3, floatOrdered(hi.x)); atomicMaxU32(b.boundsBits + 4, floatOrdered(hi.y)); atomicMaxU32(b.boundsBits + 5, floatOrdered(hi.z));
}

// Two-level structure (RTAccelerationStructure::buildTLAS, rt_acceleration_structure.cpp:419-535): the TLAS is the same
// 8-wide tree built over ONE primitive per instance.  Primitive a = instance a: its box is the world-space bounding box of
// the 8 transformed corners of the instance's object-space BLAS root box, padded by more than the rounding error of the
// transform (a hit is found in object space; its world position agrees with the transformed box only to a few ulp).
// The "triangle" record of the primitive carries the instance index where a triangle carries its global id.
OHB_HD void buildTlasPrim(const BuildArrays& b, const f4* blasLo, const f4* blasHi, const uint32_t* instOfPrim, uint32_t a) {
    const uint32_t inst = instOfPrim[a];
    const f4* rows = b.instXform + size_t(inst) * 3u;
    const f3 l = xyz(blasLo[inst]), h = xyz(blasHi[inst]);
    f3 lo = mk3(3.0e38f), hi = mk3(-3.0e38f);
    for (int c = 0; c < 8; c++) {
        f3 p = xformPoint(rows, mk3((c & 1) ? h.x : l.x, (c & 2) ? h.y : l.y, (c & 4) ? h.z : l.z));
        lo = vmin(lo, p); hi = vmax(hi, p);
    }
    const f3 e = hi - lo; const float ext = fmaxf(e.x, fmaxf(e.y, e.z));
    const float k = 4.0e-6f;
    lo = mk3(lo.x - (fabsf(lo.x) + ext) * k - 1e-30f, lo.y - (fabsf(lo.y) + ext) * k - 1e-30f, lo.z - (fabsf(lo.z) + ext) * k - 1e-30f);
    hi = mk3(hi.x + (fabsf(hi.x) + ext) * k + 1e-30f, hi.y + (fabsf(hi.y) + ext) * k + 1e-30f, hi.z + (fabsf(hi.z) + ext) * k + 1e-30f);
    b.wtri[size_t(a) * 3 + 0] = mk4(0.0f, 0.0f, 0.0f, u2f(inst));
    b.wtri[size_t(a) * 3 + 1] = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    b.wtri[size_t(a) * 3 + 2] = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    b.primLo[a] = mk4(lo, 0.0f); b.primHi[a] = mk4(hi, 0.0f);
    atomicMinU32(b.boundsBits + 0, floatOrdered(lo.x)); atomicMinU32(b.boundsBits + 1, floatOrdered(lo.y)); atomicMinU32(b.boundsBits + 2, floatOrdered(lo.z));
    atomicMaxU32(b.boundsBits + 3, floatOrdered(hi.x)); atomicMaxU32(b.boundsBits + 4, floatOrdered(hi.y)); atomicMaxU32(b.boundsBits + 5, floatOrdered(hi.z));
}
// object-space root box of the BLAS just built (the binary root's box, or the single triangle's)
OHB_HD void storeBlasRootBox(const BuildArrays& b, f4* blasLo, f4* blasHi, uint32_t inst, uint32_t* maxLevels) {
    if (b.n >= 2u) { blasLo[inst] = b.nodeLo[0]; blasHi[inst] = b.nodeHi[0]; } else { blasLo[inst] = b.primLo[0]; blasHi[inst] = b.primHi[0]; }
    atomicMaxU32(maxLevels, b.wideCounters[3]);          // BLAS builds run concurrently on several streams
}

// Stage 2: 63-bit Morton code of the AABB centre (21 bits per axis).
OHB_HD uint64_t expandBits21(uint64_t v) {
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
OHB_HD void buildMorton(const BuildArrays& b, uint32_t a) {
    f3 lo = mk3(orderedFloat(b.boundsBits[0]), orderedFloat(b.boundsBits[1]), orderedFloat(b.boundsBits[2]));
    f3 hi = mk3(orderedFloat(b.boundsBits[3]), orderedFloat(b.boundsBits[4]), orderedFloat(b.boundsBits[5]));
    f3 ext = hi - lo;
    f3 c = (xyz(b.primLo[a]) + xyz(b.primHi[a])) * 0.5f;
    float sx = ext.x > 0.0f ? (c.x - lo.x) / ext.x : 0.0f, sy = ext.y > 0.0f ? (c.y - lo.y) / ext.y : 0.0f, sz = ext.z > 0.0f ? (c.z - lo.z) / ext.z : 0.0f;
    const float S = 2097152.0f;   // 2^21
    uint64_t ix = uint64_t(fminf(fmaxf(sx * S, 0.0f), S - 1.0f)), iy = uint64_t(fminf(fmaxf(sy * S, 0.0f), S - 1.0f)), iz = uint64_t(fminf(fmaxf(sz * S, 0.0f), S - 1.0f));
    b.keys[a] = (expandBits21(ix) << 2) | (expandBits21(iy) << 1) | expandBits21(iz);
    b.vals[a] = a;
}

// Stage 4: Karras hierarchy.  delta(i,j) = common prefix length of the (key, index) pairs.
OHB_HD int karrasDelta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], c = keys[j];
    if (a == c) return 64 + clz32(uint32_t(i) ^ uint32_t(j));
    return clz64(a ^ c);
}
OHB_HD void buildHierarchyNode(const BuildArrays& b, int i) {
    const int n = int(b.n);
    const uint64_t* keys = b.keys;
    int d = (karrasDelta(keys, n, i, i + 1) - karrasDelta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = karrasDelta(keys, n, i, i - d);
    int lmax = 2;
    while (karrasDelta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (karrasDelta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = karrasDelta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {   // t = ceil(l / 2^k), k = 1, 2, ... down to 1
        if (karrasDelta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + (d < 0 ? -1 : 0);
    int lo = i < j ? i : j, hi = i < j ? j : i;
    int lc, rc;
    if (lo == gamma) { lc = ~gamma; b.parentLeaf[gamma] = i; } else { lc = gamma; b.parentInner[gamma] = i; }
    if (hi == gamma + 1) { rc = ~(gamma + 1); b.parentLeaf[gamma + 1] = i; } else { rc = gamma + 1; b.parentInner[gamma + 1] = i; }
    b.left[i] = lc; b.right[i] = rc;
    if (i == 0) b.parentInner[0] = -1;
}

// ---- coherent (L2) accesses for arrays that other threads of the same kernel write --------------------------------
OHB_HD f4 ldcg4(const f4* p) {
#if OHB_DEVICE_CODE
    float4 v = __ldcg(reinterpret_cast<const float4*>(p)); return mk4(v.x, v.y, v.z, v.w);
#else
    return *p;
#endif
}
OHB_HD int32_t ldcgi(const int32_t* p) {
#if OHB_DEVICE_CODE
    return __ldcg(p);
#else
    return *p;
#endif
}
OHB_HD float boxArea(f3 lo, f3 hi) { f3 e = hi - lo; return 2.0f * (e.x * e.y + e.y * e.z + e.z * e.x); }

// SAH constants of the optimiser (Karras & Aila 2013): Ci = 1.2 per node visit, Ct = 1 per triangle test.
#define OHB_SAH_CI 1.2f
#define OHB_SAH_CT 1.0f
#define OHB_TREELET_LEAVES 7

// Per-node record kept in nodeLo/nodeHi: (lo, triangle count as bits) and (hi, SAH cost of the subtree).
struct NodeInfo { f3 lo, hi; uint32_t cnt; float cost; };
OHB_HD NodeInfo nodeInfo(const BuildArrays& b, int c) {
    NodeInfo r;
    if (c < 0) {
        uint32_t a = b.vals[~c]; r.lo = xyz(b.primLo[a]); r.hi = xyz(b.primHi[a]); r.cnt = 1u; r.cost = OHB_SAH_CT * boxArea(r.lo, r.hi);
    } else {
        f4 l = ldcg4(b.nodeLo + c), h = ldcg4(b.nodeHi + c); r.lo = xyz(l); r.hi = xyz(h); r.cnt = f2u(l.w); r.cost = h.w;
    }
    return r;
}
OHB_HD void writeNode(const BuildArrays& b, int n, f3 lo, f3 hi, uint32_t cnt, float childCost) {
    float A = boxArea(lo, hi);
    float cost = (cnt <= OHB_MAX_LEAF) ? OHB_SAH_CT * A * float(cnt) : OHB_SAH_CI * A + childCost;   // <= OHB_MAX_LEAF triangles collapse to one leaf
    b.nodeLo[n] = mk4(lo, u2f(cnt)); b.nodeHi[n] = mk4(hi, cost);
}

// Treelet restructuring (Karras & Aila 2013, "Fast parallel construction of high-quality bounding volume
// hierarchies", §3): the treelet of up to 7 leaves with the largest surface areas below `root` is rebuilt into the
// SAH-optimal topology by dynamic programming over its 2^7 leaf subsets; the treelet's own internal nodes are
// reused, subtrees hanging below the treelet leaves are untouched.  One thread per treelet (the DP tables live in
// local memory); sibling subtrees are processed concurrently by the bottom-up sweep.
OHB_HD void optimizeTreelet(const BuildArrays& b, int root) {
    const int NL = OHB_TREELET_LEAVES;
    int leaves[NL], internal[NL - 1]; int nl = 2, ni = 1;
    internal[0] = root; leaves[0] = ldcgi(b.left + root); leaves[1] = ldcgi(b.right + root);
    NodeInfo info[NL];
    info[0] = nodeInfo(b, leaves[0]); info[1] = nodeInfo(b, leaves[1]);
    while (nl < NL) {
        int best = -1; float bestA = -1.0f;
        for (int k = 0; k < nl; k++) if (leaves[k] >= 0 && info[k].cnt > OHB_MAX_LEAF) { float a = boxArea(info[k].lo, info[k].hi); if (a > bestA) { bestA = a; best = k; } }
        if (best < 0) break;
        int c = leaves[best]; internal[ni++] = c;
        leaves[best] = ldcgi(b.left + c); leaves[nl] = ldcgi(b.right + c);
        info[best] = nodeInfo(b, leaves[best]); info[nl] = nodeInfo(b, leaves[nl]); nl++;
    }
    if (nl < 3) return;
    const int full = (1 << nl) - 1;
    float area[1 << NL], copt[1 << NL]; uint8_t part[1 << NL];
    for (int s = 1; s <= full; s++) {
        f3 lo = mk3(3.0e38f), hi = mk3(-3.0e38f);
        for (int k = 0; k < nl; k++) if (s & (1 << k)) { lo = vmin(lo, info[k].lo); hi = vmax(hi, info[k].hi); }
        area[s] = boxArea(lo, hi);
    }
    for (int k = 0; k < nl; k++) copt[1 << k] = info[k].cost;
    for (int s = 3; s <= full; s++) {
        if ((s & (s - 1)) == 0) continue;                 // singletons
        float best = 3.0e38f; int bp = 0;
        int delta = (s - 1) & s;                          // s without its lowest bit
        int p = (-delta) & s;
        do { float c = copt[p] + copt[s ^ p]; if (c < best) { best = c; bp = p; } p = (p - delta) & s; } while (p != 0);
        copt[s] = OHB_SAH_CI * area[s] + best; part[s] = uint8_t(bp);
    }
    float oldCost = ldcg4(b.nodeHi + root).w;
    if (!(copt[full] < oldCost * 0.99999f)) return;       // keep the current topology unless strictly better
    // rebuild: the treelet's internal nodes are re-assigned top-down from the full set
    int stackN[NL], stackS[NL]; int sp = 0, next = 1;
    stackN[0] = root; stackS[0] = full; sp = 1;
    while (sp) {
        sp--; int n = stackN[sp], s = stackS[sp];
        int sides[2] = {int(part[s]), s ^ int(part[s])}; int child[2];
        for (int h = 0; h < 2; h++) {
            int q = sides[h];
            if ((q & (q - 1)) == 0) {
                int k = 0; while (!(q & (1 << k))) k++;
                child[h] = leaves[k];
                if (child[h] >= 0) b.parentInner[child[h]] = n; else b.parentLeaf[~child[h]] = n;
            } else {
                child[h] = internal[next++]; b.parentInner[child[h]] = n;
                stackN[sp] = child[h]; stackS[sp] = q; sp++;
            }
        }
        b.left[n] = child[0]; b.right[n] = child[1];
        f3 lo = mk3(3.0e38f), hi = mk3(-3.0e38f); uint32_t cnt = 0u;
        for (int k = 0; k < nl; k++) if (s & (1 << k)) { lo = vmin(lo, info[k].lo); hi = vmax(hi, info[k].hi); cnt += info[k].cnt; }
        b.nodeLo[n] = mk4(lo, u2f(cnt)); b.nodeHi[n] = mk4(hi, copt[s]);
    }
}

// Stage 5: bottom-up sweep.  Thread per sorted leaf; the second arrival at a node merges its children (box, triangle
// count, SAH cost) and, when `gamma` > 0 and the subtree holds at least gamma triangles, restructures the treelet
// rooted there before climbing on.  gamma = 0 is the plain refit.
OHB_HD void sweepFromLeaf(const BuildArrays& b, uint32_t leaf, uint32_t gamma) {
    int node = b.parentLeaf[leaf];
    while (node >= 0) {
        fenceDevice();
        if (atomicIncU32(b.visit + node) == 0u) return;    // first arrival: sibling not ready
        fenceDevice();
        NodeInfo l = nodeInfo(b, ldcgi(b.left + node)), r = nodeInfo(b, ldcgi(b.right + node));
        uint32_t cnt = l.cnt + r.cnt;
        writeNode(b, node, vmin(l.lo, r.lo), vmax(l.hi, r.hi), cnt, l.cost + r.cost);
        if (gamma && cnt >= gamma) optimizeTreelet(b, node);
        node = ldcgi(b.parentInner + node);
    }
}

OHB_HD uint32_t nodeCount(const BuildArrays& b, int i) { return f2u(b.nodeLo[i].w); }
// Conservative outward padding of stored boxes (DESIGN.md "Robustness").
OHB_HD void padBox(f3& lo, f3& hi) {
    const float k = 9.5367431640625e-7f;   // 2^-20
    lo = mk3(lo.x - fabsf(lo.x) * k - 1e-30f, lo.y - fabsf(lo.y) * k - 1e-30f, lo.z - fabsf(lo.z) * k - 1e-30f);
    hi = mk3(hi.x + fabsf(hi.x) * k + 1e-30f, hi.y + fabsf(hi.y) * k + 1e-30f, hi.z + fabsf(hi.z) * k + 1e-30f);
}

// ---------------------------------------------------------------------------------------------
// Stage 6: collapse the optimised binary tree into the compressed 8-wide BVH (node format: ohb_traverse.h).
// Top-down, one level per launch, one thread per wide node (WideItem).  A wide node is grown greedily from
// a binary node: the child with the largest surface area that still holds more than OHB_MAX_LEAF triangles
// is replaced by its two children until 8 children exist or only leaf-sized subtrees remain.  Children are
// placed in the 8 slots so that (slot ^ ray octant) approximates a front-to-back order (greedy assignment on
// dot(child centre - node centre, octant sign), Ylitie et al. 2017 §3.2).  Child planes are stored as bf16 offsets
// from the node's min corner, rounded outward; the rounding is verified in fp64.
// Triangle layout: a node's leaf triangles first (<= 24, slot order), then the block of each inner child in
// slot order — every subtree is one contiguous range, computed top-down from the subtree triangle counts.
// ---------------------------------------------------------------------------------------------
struct WideItem { int32_t bvh2; uint32_t wide, triStart, pad; };

// append the triangles of the leaf-sized binary subtree `c` to the triangle array, returns how many
OHB_HD uint32_t emitLeafTris(const BuildArrays& b, int c, uint32_t dst) {
    int st[4]; int sp = 0; st[sp++] = c; uint32_t k = 0;
    while (sp) {
        int x = st[--sp];
        if (x < 0) {
            uint32_t a = b.vals[~x];
            // transposed record (x0 x1 x2 id)(y0 y1 y2 id)(z0 z1 z2 id): the ray's axis permutation becomes a row address
            const f4 v0 = b.wtri[size_t(a) * 3 + 0], v1 = b.wtri[size_t(a) * 3 + 1], v2 = b.wtri[size_t(a) * 3 + 2];
            b.tris[size_t(dst + k) * 3 + 0] = mk4(v0.x, v1.x, v2.x, v0.w);
            b.tris[size_t(dst + k) * 3 + 1] = mk4(v0.y, v1.y, v2.y, v0.w);
            b.tris[size_t(dst + k) * 3 + 2] = mk4(v0.z, v1.z, v2.z, v0.w);
            k++;
        } else { st[sp++] = b.right[x]; st[sp++] = b.left[x]; }
    }
    return k;
}
OHB_HD void emitWideNode(const BuildArrays& b, const WideItem& it, WideItem* outItems, uint32_t* outCount) {
    int child[8]; f3 lo[8], hi[8]; uint32_t cnt[8]; int nc;
    auto fetch = [&](int k, int c) { NodeInfo r = nodeInfo(b, c); child[k] = c; lo[k] = r.lo; hi[k] = r.hi; cnt[k] = r.cnt; };
    if (b.n == 1u) { fetch(0, ~0); nc = 1; }
    else { fetch(0, b.left[it.bvh2]); fetch(1, b.right[it.bvh2]); nc = 2; }
    while (nc < 8) {
        int best = -1; float bestA = -1.0f;
        for (int k = 0; k < nc; k++) if (cnt[k] > OHB_MAX_LEAF) { float a = boxArea(lo[k], hi[k]); if (a > bestA) { bestA = a; best = k; } }
        if (best < 0) break;
        int c = child[best];
        fetch(best, b.left[c]); fetch(nc, b.right[c]); nc++;
    }
    // Free slots left and only leaf-sized subtrees: split the multi-triangle leaves too (largest area x count first).
    // The slab test of a visit costs the same for 8 slots as for 4, so every extra box is free culling of triangle tests.
    while (nc < 8) {
        int best = -1; float bestA = -1.0f;
        for (int k = 0; k < nc; k++) if (cnt[k] >= 2u && cnt[k] <= OHB_MAX_LEAF) { float a = boxArea(lo[k], hi[k]) * float(cnt[k]); if (a > bestA) { bestA = a; best = k; } }
        if (best < 0) break;
        int c = child[best];
        fetch(best, b.left[c]); fetch(nc, b.right[c]); nc++;
    }
    // node box = union of the padded child boxes
    f3 nlo = mk3(3.0e38f), nhi = mk3(-3.0e38f);
    float leafArea = 0.0f;
    for (int k = 0; k < nc; k++) {
        if (cnt[k] <= OHB_MAX_LEAF) leafArea += boxArea(lo[k], hi[k]) * float(cnt[k]);
        padBox(lo[k], hi[k]); nlo = vmin(nlo, lo[k]); nhi = vmax(nhi, hi[k]);
    }
    atomicAddF32(b.sah + 0, boxArea(nlo, nhi)); atomicAddF32(b.sah + 1, leafArea);
    // slot assignment
    int slotOf[8], childAt[8];
    for (int k = 0; k < 8; k++) { slotOf[k] = -1; childAt[k] = -1; }
    f3 nc2 = nlo + nhi;                                    // 2 x node centre
    for (int round = 0; round < nc; round++) {
        float bestC = 3.0e38f; int bk = -1, bs = -1;
        for (int k = 0; k < nc; k++) {
            if (slotOf[k] >= 0) continue;
            f3 dc = (lo[k] + hi[k]) - nc2;
            for (int s = 0; s < 8; s++) {
                if (childAt[s] >= 0) continue;
                float c = ((s & 4) ? -dc.x : dc.x) + ((s & 2) ? -dc.y : dc.y) + ((s & 1) ? -dc.z : dc.z);
                if (c < bestC) { bestC = c; bk = k; bs = s; }
            }
        }
        slotOf[bk] = bs; childAt[bs] = bk;
    }
    // bf16 plane offsets from the node's min corner p, rounded outward and verified in fp64 against the padded child
    // boxes.  p sits a second pad below the node box, so every lo offset is >= 1e-30: far above the < 9.2e-41 of junk
    // the even slot's bits add below an odd slot's bf16 when the traversal uses the whole word as the fp32 operand.
    float pmin[3]; uint32_t qlo[3][8], qhi[3][8]; float extMax = 0.0f;
    for (int ax = 0; ax < 3; ax++) {
        const float nl = comp(nlo, ax);
        const float pf = nl - (fabsf(nl) * 9.5367431640625e-7f + 1e-30f);
        pmin[ax] = pf;
        const double p = double(pf);
        extMax = fmaxf(extMax, float((double(comp(nhi, ax)) - p) * 1.000001));
        for (int s = 0; s < 8; s++) {
            int k = childAt[s];
            if (k < 0) { qlo[ax][s] = OHB_EMPTY_LO; qhi[ax][s] = 0u; continue; }
            const double l = double(comp(lo[k], ax)), h = double(comp(hi[k], ax));
            uint32_t ql = f2u(float(l - p)) >> 16;                       // truncation = round toward zero (the offset is positive)
            while (ql > 0u && p + double(u2f(ql << 16)) > l) ql--;
            uint32_t fh = f2u(float(h - p)), qh = (fh >> 16) + ((fh & 0xFFFFu) ? 1u : 0u);
            while (p + double(u2f(qh << 16)) < h) qh++;
            qlo[ax][s] = ql; qhi[ax][s] = qh;
        }
        // odd slots of the lo planes are read as the whole word: bf16 << 16 | (even slot's bf16)
        for (int s = 1; s < 8; s += 2) {
            int k = childAt[s];
            if (k < 0) continue;
            const double l = double(comp(lo[k], ax));
            while (qlo[ax][s] > 0u && p + double(u2f((qlo[ax][s] << 16) | qlo[ax][s - 1])) > l) qlo[ax][s]--;
        }
    }
    uint32_t ext16 = (f2u(extMax) >> 16) + 1u;
    // children: inner ones get consecutive wide-node indices in slot order, leaves their triangle offsets
    uint32_t imask = 0u, numInner = 0u;
    for (int s = 0; s < 8; s++) { int k = childAt[s]; if (k >= 0 && cnt[k] > OHB_MAX_LEAF) { imask |= 1u << s; numInner++; } }
    uint32_t childBase = 0u, itemBase = 0u;
    if (numInner) {
        childBase = atomicIncU32N(b.wideCounters + 0, numInner);
        itemBase = atomicIncU32N(outCount, numInner);
    }
    uint32_t meta[8]; uint32_t triOff = 0u, rank = 0u;
    for (int s = 0; s < 8; s++) {
        int k = childAt[s];
        if (k < 0) { meta[s] = 0u; continue; }
        if (cnt[k] > OHB_MAX_LEAF) continue;
        uint32_t m = emitLeafTris(b, child[k], it.triStart + triOff);
        meta[s] = (((1u << m) - 1u) << 5) | triOff;
        triOff += m;
    }
    uint32_t sub = it.triStart + triOff;
    for (int s = 0; s < 8; s++) {
        int k = childAt[s];
        if (k < 0 || cnt[k] <= OHB_MAX_LEAF) continue;
        meta[s] = (1u << 5) | (24u + uint32_t(s));
        WideItem w; w.bvh2 = child[k]; w.wide = childBase + rank; w.triStart = sub; w.pad = 0u;
        outItems[itemBase + rank] = w;
        rank++; sub += cnt[k];
    }
    auto pack4 = [](const uint32_t* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); };
    auto pack8 = [](const uint32_t* v) { u4 r; r.x = v[0] | (v[1] << 16); r.y = v[2] | (v[3] << 16); r.z = v[4] | (v[5] << 16); r.w = v[6] | (v[7] << 16); return r; };
    u4 w0, w1;
    w0.x = f2u(pmin[0]); w0.y = f2u(pmin[1]); w0.z = f2u(pmin[2]);
    w0.w = imask | (ext16 << 16);
    w1.x = childBase; w1.y = it.triStart; w1.z = pack4(meta); w1.w = pack4(meta + 4);
    u4* np = b.wnodes + size_t(it.wide) * OHB_WNODE_VECS;
    const uint32_t swz = OHB_NODE_SWZ(it.wide);                  // bank swizzle of the four 32-B pieces (ohb_traverse.h)
    np[2u * (0u ^ swz)] = w0; np[2u * (0u ^ swz) + 1u] = w1;
    for (uint32_t ax = 0; ax < 3; ax++) { np[2u * ((1u + ax) ^ swz)] = pack8(qlo[ax]); np[2u * ((1u + ax) ^ swz) + 1u] = pack8(qhi[ax]); }
}

}  // namespace ohb

// ohb_common.h — portable scalar/vector helpers shared by every kernel of the B200 path tracer.
//
// All per-thread logic of the product lives in OHB_HD inline functions so that the same source
// compiles (a) into the sm_100a kernels of libohao_b200.so and (b) into tests/emul, a host-only
// single-threaded kernel emulator used for debugging in the GPU-less build container.  The
// emulator is test tooling: libohao_b200.so has no CPU path and ohb_create() fails without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define OHB_HD __host__ __device__ __forceinline__
#define OHB_D __device__ __forceinline__
#else
#define OHB_HD inline
#define OHB_D inline
#endif

#if defined(__CUDA_ARCH__)
#define OHB_DEVICE_CODE 1
#else
#define OHB_DEVICE_CODE 0
#endif

namespace ohb {

struct alignas(8) f2 { float x, y; };
struct f3 { float x, y, z; };
struct alignas(16) f4 { float x, y, z, w; };
struct alignas(16) u4 { uint32_t x, y, z, w; };

OHB_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
OHB_HD f3 mk3(float a) { return mk3(a, a, a); }
OHB_HD f4 mk4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
OHB_HD f4 mk4(f3 a, float w) { return mk4(a.x, a.y, a.z, w); }
OHB_HD f3 xyz(f4 a) { return mk3(a.x, a.y, a.z); }
OHB_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
OHB_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
OHB_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
OHB_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
OHB_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
OHB_HD f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
// vec3 / float with one reciprocal.  Device: MUFU.RCP (1 ulp) instead of the 16-instruction IEEE sequence (code size, see normalize).
OHB_HD f3 operator/(f3 a, float s) {
#if defined(__CUDA_ARCH__)
    float r = __fdividef(1.0f, s);
#else
    float r = 1.0f / s;
#endif
    return mk3(a.x * r, a.y * r, a.z * r);
}
// "Evaluation-class" division: radiance, BSDF and pdf arithmetic only — never a value that decides where a ray goes (camera
// rays, sampled directions, shading frames keep IEEE `/`), so a 2-ulp quotient moves a sample by 1e-7 of its radiance and
// cannot feed the scene's chaotic amplification of path geometry.  Device: MUFU.RCP + multiply instead of the 9-instruction
// IEEE sequence (6 % of k_shade's instructions, profiles/r2au); with -prec-div=false on EVERYTHING the 2 M scene's per-sample
// mismatch grew 3.5x (profiles/r2av), which is why the flag is not used.  Host: plain division.
#ifndef OHB_FAST_EDIV
#define OHB_FAST_EDIV 1
#endif
OHB_HD float ediv(float a, float b) {
#if defined(__CUDA_ARCH__) && OHB_FAST_EDIV
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
OHB_HD f3 operator/(f3 a, f3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
OHB_HD f3& operator+=(f3& a, f3 b) { a = a + b; return a; }
OHB_HD f3& operator*=(f3& a, f3 b) { a = a * b; return a; }
OHB_HD f3& operator*=(f3& a, float s) { a = a * s; return a; }
OHB_HD f3& operator/=(f3& a, float s) { a = a / s; return a; }
OHB_HD float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
OHB_HD f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
OHB_HD float length(f3 a) { return sqrtf(dot(a, a)); }
// GLSL normalize().  Device: one MUFU.RSQ (<= 2 ulp) like the inversesqrt a GPU driver emits for the shader; the
// IEEE sqrt + IEEE reciprocal it replaces was 16 instructions at each of ~25 call sites, 9 % of k_bounce's code, and
// the shading kernels are instruction-fetch bound (their bodies are 2-3x the 32 KB L1.5 instruction cache).
OHB_HD f3 normalize(f3 a) {
#if OHB_DEVICE_CODE
    return a * rsqrtf(dot(a, a));
#else
    return a * (1.0f / sqrtf(dot(a, a)));
#endif
}
OHB_HD f3 vabs(f3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
OHB_HD f3 vmin(f3 a, f3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
OHB_HD f3 vmax(f3 a, f3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
OHB_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
OHB_HD float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
OHB_HD f3 mix(f3 a, f3 b, float t) { return a * (1.0f - t) + b * t; }
OHB_HD f3 mix(f3 a, f3 b, f3 t) { return a * (mk3(1.0f) - t) + b * t; }
OHB_HD f3 reflect(f3 I, f3 N) { return I - N * (2.0f * dot(N, I)); }
// Accurate libm-grade transcendentals, ONE shared copy per kernel: the slow paths of sinf/cosf/powf/atan2f
// are hundreds of instructions each, and inlining them at ~40 call sites made the shading kernel 200 KB of
// SASS that thrashed the instruction cache (ncu: stall_no_instruction 13.4 of 21 cycles per issue, r01).
#if defined(__CUDACC__)
#define OHB_SHARED_FN static __device__ __host__ __noinline__
#else
#define OHB_SHARED_FN static inline
#endif
OHB_SHARED_FN float ohb_sin(float x) { return sinf(x); }
OHB_SHARED_FN float ohb_cos(float x) { return cosf(x); }
// sin and cos of one angle.  OHB_TRIG_MODE 0: sinf + cosf, two argument reductions; 1: sincosf, one reduction, the values
// sinf / cosf return; 2 (ohb_sincos_turns only): sincospif — no 2*pi rounding of the angle, no slow path (A/B: profiles/r2af)
#ifndef OHB_TRIG_MODE
#define OHB_TRIG_MODE 1
#endif
OHB_SHARED_FN void ohb_sincos(float x, float& s, float& c) {
#if OHB_DEVICE_CODE && OHB_TRIG_MODE >= 1
    sincosf(x, &s, &c);
#else
    s = sinf(x); c = cosf(x);
#endif
}
// sin and cos of 2*pi*turns (the shader's `phi = 6.2831853 * u`)
OHB_SHARED_FN void ohb_sincos_turns(float turns, float& s, float& c) {
#if OHB_DEVICE_CODE && OHB_TRIG_MODE == 2
    sincospif(2.0f * turns, &s, &c);
#elif OHB_DEVICE_CODE && OHB_TRIG_MODE == 1
    sincosf(6.2831853f * turns, &s, &c);
#else
    const float x = 6.2831853f * turns; s = sinf(x); c = cosf(x);
#endif
}
OHB_SHARED_FN float ohb_pow(float x, float e) { return powf(x, e); }
OHB_SHARED_FN float ohb_atan2(float y, float x) { return atan2f(y, x); }
OHB_HD f3 vpow(f3 a, float e) { return mk3(ohb_pow(a.x, e), ohb_pow(a.y, e), ohb_pow(a.z, e)); }
// The closest-hit shader's sRGB decode `pow(texel, 2.2)` on x >= 0: x^2 * 2^(0.2 log2 x) — the exponent's error enters scaled
// by 0.2 instead of 2.2, so MUFU.LG2 / MUFU.EX2 accuracy (2 ulp each) gives <= ~5 ulp against powf at 7 instructions
// instead of powf's ~50 (8.5 % of the primary k_shade's instructions on the textured scene, profiles/r2ai).  Host code: powf.
#ifndef OHB_FAST_POW
#define OHB_FAST_POW 1
#endif
OHB_HD float pow22(float x) {
#if OHB_DEVICE_CODE && OHB_FAST_POW
    return (x * x) * exp2f(0.2f * __log2f(x));
#else
    return ohb_pow(x, 2.2f);
#endif
}
OHB_HD f3 vpow22(f3 a) { return mk3(pow22(a.x), pow22(a.y), pow22(a.z)); }
OHB_HD float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }
OHB_HD float maxcomp(f3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
OHB_HD float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
OHB_HD float smoothstepf(float e0, float e1, float x) { float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
OHB_HD float luminance(f3 c) { return dot(c, mk3(0.2126f, 0.7152f, 0.0722f)); }
OHB_HD float comp(f3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

OHB_HD uint32_t f2u(float f) {
#if OHB_DEVICE_CODE
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
OHB_HD float u2f(uint32_t u) {
#if OHB_DEVICE_CODE
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
OHB_HD uint32_t brev32(uint32_t v) {
#if OHB_DEVICE_CODE
    return __brev(v);
#else
    v = (v << 16) | (v >> 16);
    v = ((v & 0x00FF00FFu) << 8) | ((v & 0xFF00FF00u) >> 8);
    v = ((v & 0x0F0F0F0Fu) << 4) | ((v & 0xF0F0F0F0u) >> 4);
    v = ((v & 0x33333333u) << 2) | ((v & 0xCCCCCCCCu) >> 2);
    v = ((v & 0x55555555u) << 1) | ((v & 0xAAAAAAAAu) >> 1);
    return v;
#endif
}
OHB_HD int clz64(uint64_t v) {
#if OHB_DEVICE_CODE
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}
OHB_HD int clz32(uint32_t v) {
#if OHB_DEVICE_CODE
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

// ---- exactly-rounded fp32 primitives (never contracted into FMA) for the intersection spec ----
OHB_HD float xmul(float a, float b) {
#if OHB_DEVICE_CODE
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
OHB_HD float xadd(float a, float b) {
#if OHB_DEVICE_CODE
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
OHB_HD float xsub(float a, float b) {
#if OHB_DEVICE_CODE
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
OHB_HD float xdiv(float a, float b) {
#if OHB_DEVICE_CODE
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}

// ---- 128-bit read-only loads -------------------------------------------------------------------
OHB_HD f4 ld4(const f4* p) {
#if OHB_DEVICE_CODE
    float4 v = __ldg(reinterpret_cast<const float4*>(p)); return mk4(v.x, v.y, v.z, v.w);
#else
    return *p;
#endif
}
OHB_HD u4 ldu4(const u4* p) {
#if OHB_DEVICE_CODE
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); u4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    return *p;
#endif
}
OHB_HD uint32_t ldg(const uint32_t* p) {
#if OHB_DEVICE_CODE
    return __ldg(p);
#else
    return *p;
#endif
}
OHB_HD float ldg(const float* p) {
#if OHB_DEVICE_CODE
    return __ldg(p);
#else
    return *p;
#endif
}
OHB_HD f2 ldg(const f2* p) {
#if OHB_DEVICE_CODE
    float2 v = __ldg(reinterpret_cast<const float2*>(p)); f2 r; r.x = v.x; r.y = v.y; return r;
#else
    return *p;
#endif
}

// ---- queue slot allocation: warp-aggregated atomic on the device, plain increment in the emulator
OHB_HD uint32_t alloc_slot(uint32_t* counter) {
#if OHB_DEVICE_CODE
    unsigned m = __activemask();
    unsigned lane = threadIdx.x & 31u;
    int leader = __ffs((int)m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
#else
    return (*counter)++;
#endif
}
OHB_HD void count_add(unsigned long long* counter, unsigned long long v) {
#if OHB_DEVICE_CODE
    atomicAdd(counter, v);
#else
    *counter += v;
#endif
}

}  // namespace ohb

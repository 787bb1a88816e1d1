// oracle/oracle_math.h — TEST INFRASTRUCTURE ONLY (see oracle.cpp header).
// Small GLSL-flavoured vector helpers for the CPU oracle.  Plain fp32, compiled with
// -ffp-contract=off so every a*b+c below rounds twice unless fmaf() is spelled out.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace orc {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

static inline V3 v3(float a) { return {a, a, a}; }
static inline V3 v3(float x, float y, float z) { return {x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
static inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
static inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
static inline V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
static inline V3& operator*=(V3& a, float s) { a = a * s; return a; }
static inline V3& operator/=(V3& a, float s) { a = a / s; return a; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
static inline float length(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 normalize(V3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
static inline V3 vabs(V3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
static inline V3 vmax(V3 a, V3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
static inline V3 vmin(V3 a, V3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
static inline float clampf(float x, float lo, float hi) { return std::min(std::max(x, lo), hi); }
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline V3 mix(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
static inline V3 mix(V3 a, V3 b, V3 t) { return a * (v3(1.0f) - t) + b * t; }
static inline V3 reflect(V3 I, V3 N) { return I - N * (2.0f * dot(N, I)); }
static inline V3 vpow(V3 a, float e) { return {std::pow(a.x, e), std::pow(a.y, e), std::pow(a.z, e)}; }
static inline float maxcomp(V3 a) { return std::max(a.x, std::max(a.y, a.z)); }
static inline float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
static inline float smoothstep(float e0, float e1, float x) {
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
static inline float luminance(V3 c) { return dot(c, v3(0.2126f, 0.7152f, 0.0722f)); }
static inline float get(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// Column-major 4x4 (glm::mat4 storage): m[c*4+r].
struct M4 { float m[16]; float at(int r, int c) const { return m[c * 4 + r]; } };
static inline V4 mulv(const M4& a, V4 v) {
    return {a.m[0] * v.x + a.m[4] * v.y + a.m[8] * v.z + a.m[12] * v.w,
            a.m[1] * v.x + a.m[5] * v.y + a.m[9] * v.z + a.m[13] * v.w,
            a.m[2] * v.x + a.m[6] * v.y + a.m[10] * v.z + a.m[14] * v.w,
            a.m[3] * v.x + a.m[7] * v.y + a.m[11] * v.z + a.m[15] * v.w};
}
static inline M4 mulm(const M4& a, const M4& b) {
    M4 r;
    for (int c = 0; c < 4; c++)
        for (int rr = 0; rr < 4; rr++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += a.m[k * 4 + rr] * b.m[c * 4 + k];
            r.m[c * 4 + rr] = s;
        }
    return r;
}
// General inverse by cofactors (what glm::inverse does), fp32.
static inline M4 inverse(const M4& in) {
    const float* m = in.m;
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    float id = 1.0f / det;
    M4 out;
    for (int i = 0; i < 16; i++) out.m[i] = inv[i] * id;
    return out;
}

// 3x3 helpers (row-major r[row][col]) for the per-instance normal matrix.
struct M3 { float r[3][3]; };
static inline V3 mulv(const M3& a, V3 v) {
    return {a.r[0][0] * v.x + a.r[0][1] * v.y + a.r[0][2] * v.z,
            a.r[1][0] * v.x + a.r[1][1] * v.y + a.r[1][2] * v.z,
            a.r[2][0] * v.x + a.r[2][1] * v.y + a.r[2][2] * v.z};
}
static inline M3 inverse_transpose(const M3& a) {
    // inverse = adj/det ; transpose(inverse) = cofactor matrix / det
    M3 c;
    c.r[0][0] = a.r[1][1] * a.r[2][2] - a.r[1][2] * a.r[2][1];
    c.r[0][1] = a.r[1][2] * a.r[2][0] - a.r[1][0] * a.r[2][2];
    c.r[0][2] = a.r[1][0] * a.r[2][1] - a.r[1][1] * a.r[2][0];
    c.r[1][0] = a.r[0][2] * a.r[2][1] - a.r[0][1] * a.r[2][2];
    c.r[1][1] = a.r[0][0] * a.r[2][2] - a.r[0][2] * a.r[2][0];
    c.r[1][2] = a.r[0][1] * a.r[2][0] - a.r[0][0] * a.r[2][1];
    c.r[2][0] = a.r[0][1] * a.r[1][2] - a.r[0][2] * a.r[1][1];
    c.r[2][1] = a.r[0][2] * a.r[1][0] - a.r[0][0] * a.r[1][2];
    c.r[2][2] = a.r[0][0] * a.r[1][1] - a.r[0][1] * a.r[1][0];
    float det = a.r[0][0] * c.r[0][0] + a.r[0][1] * c.r[0][1] + a.r[0][2] * c.r[0][2];
    float id = 1.0f / det;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c.r[i][j] *= id;
    return c;
}

}  // namespace orc

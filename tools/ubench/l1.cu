// tools/ubench/l1.cu — L1/shared-memory gather throughput on B200 for the access shapes of BVH traversal.
// Each lane reads W bytes (4/8/16/32) per load from a pseudo-random, W-aligned address inside a working set that fits L1
// (or lives in shared memory).  "grp" = lanes that share one 128-B line per load (1 = fully divergent, 32 = coalesced).
// Reports lane-loads and bytes per clock per SM.  Measurement tooling only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CHK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
#define ITER 512

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int W> __device__ __forceinline__ uint32_t ldw(const char* p) {
    if (W == 4) { uint32_t v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
    if (W == 8) { uint32_t a, b; asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p)); return a ^ b; }
    if (W == 16) { uint32_t a, b, c, d; asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p)); return a ^ b ^ c ^ d; }
    if (W == 32) { uint32_t a, b, c, d, e, f, g, h; asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p)); return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h; }
    return 0;
}
template <int W> __device__ __forceinline__ uint32_t ldsw(uint32_t a) {
    if (W == 4) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
    if (W == 8) { uint32_t x, y; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(a)); return x ^ y; }
    if (W == 16) { uint32_t x, y, z, w; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a)); return x ^ y ^ z ^ w; }
    return 0;
}

// global gather: working set `bytes` (power of two), `grp` lanes share a line
template <int W>
__global__ void __launch_bounds__(256) kg(const char* buf, uint32_t bytes, int grp, int sameNode, uint32_t* out) {
    uint32_t s = (blockIdx.x * 256u + threadIdx.x) / uint32_t(grp) * 2654435761u + 12345u, acc = 0;
    const uint32_t sub = (threadIdx.x % uint32_t(grp)) * uint32_t(W) & 127u;
    for (int it = 0; it < ITER; it++) {
        uint32_t r = lcg(s);
        uint32_t line = (r * 128u) & (bytes - 1u);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // sameNode: the 4 loads walk one 128-B line (a node fetch); else 4 independent lines
            uint32_t off;
            if (sameNode == 0) off = ((line + uint32_t(k) * 8192u + uint32_t(k) * 128u) & (bytes - 1u)) + sub;
            else if (sameNode == 1) off = line + ((sub + uint32_t(k) * uint32_t(W)) & 127u);
            else if (sameNode == 2) off = line + ((threadIdx.x * uint32_t(W) + uint32_t(k) * uint32_t(W)) & 127u);        // offset rotated by lane
            else off = line + ((((r >> 12) & 7u) * 16u + uint32_t(k) * uint32_t(W)) & (128u - uint32_t(W)));             // offset rotated by a random amount
            acc ^= ldw<W>(buf + off);
        }
    }
    if (acc == 0x1234567u) out[0] = acc;
}
template <int W>
__global__ void __launch_bounds__(256) ks(uint32_t bytes, uint32_t* out) {
    extern __shared__ uint32_t sm[];
    for (uint32_t i = threadIdx.x; i < bytes / 4; i += 256) sm[i] = i * 2654435761u;
    __syncthreads();
    uint32_t s = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    for (int it = 0; it < ITER; it++) {
        uint32_t r = lcg(s);
        uint32_t line = (r * 128u) & (bytes - 1u);
#pragma unroll
        for (int k = 0; k < 4; k++) acc ^= ldsw<W>(base + line + uint32_t(k) * uint32_t(W));
    }
    if (acc == 0x1234567u) out[0] = acc;
}

static double g_mhz; static int g_sms;
template <int W> int runG(const char* buf, uint32_t bytes, int grp, int sameNode, uint32_t* out, int ctasPerSM) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kg<W><<<g_sms * ctasPerSM, 256>>>(buf, bytes, grp, sameNode, out); CHK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) { cudaEventRecord(a); kg<W><<<g_sms * ctasPerSM, 256>>>(buf, bytes, grp, sameNode, out); cudaEventRecord(b); CHK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    double laneLoads = double(g_sms) * ctasPerSM * 256.0 * ITER * 4; double cyc = best * 1e-3 * g_mhz * 1e6;
    printf("global W=%2d B  set %7u B  grp %2d  %s  ctas/SM %d : %6.3f lane-loads/clk/SM  %7.1f B/clk/SM  (%.3f ms)\n", W, bytes, grp, sameNode == 0 ? "4 lines    " : sameNode == 1 ? "one line x4" : sameNode == 2 ? "rot by lane" : "rot random ", ctasPerSM, laneLoads / cyc / g_sms, laneLoads * W / cyc / g_sms, best);
    return 0;
}
template <int W> int runS(uint32_t bytes, uint32_t* out, int ctasPerSM) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaFuncSetAttribute(ks<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    ks<W><<<g_sms * ctasPerSM, 256, bytes>>>(bytes, out); CHK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) { cudaEventRecord(a); ks<W><<<g_sms * ctasPerSM, 256, bytes>>>(bytes, out); cudaEventRecord(b); CHK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    double laneLoads = double(g_sms) * ctasPerSM * 256.0 * ITER * 4; double cyc = best * 1e-3 * g_mhz * 1e6;
    printf("shared W=%2d B  set %7u B  random line x4 words   ctas/SM %d : %6.3f lane-loads/clk/SM  %7.1f B/clk/SM  (%.3f ms)\n", W, bytes, ctasPerSM, laneLoads / cyc / g_sms, laneLoads * W / cyc / g_sms, best);
    return 0;
}
int main() {
    cudaDeviceProp p; CHK(cudaGetDeviceProperties(&p, 0)); g_sms = p.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); g_mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz\n", p.name, g_sms, g_mhz);
    char* buf; CHK(cudaMalloc(&buf, 64u << 20)); CHK(cudaMemset(buf, 1, 64u << 20)); uint32_t* out; CHK(cudaMalloc(&out, 4));
    for (uint32_t bytes : {65536u, 8u << 20}) for (int sameNode : {0, 1}) {
        runG<4>(buf, bytes, 1, sameNode, out, 4); runG<8>(buf, bytes, 1, sameNode, out, 4); runG<16>(buf, bytes, 1, sameNode, out, 4); runG<32>(buf, bytes, 1, sameNode, out, 4);
    }
    for (uint32_t bytes : {65536u, 8u << 20}) for (int mode : {2, 3}) { runG<4>(buf, bytes, 1, mode, out, 4); runG<16>(buf, bytes, 1, mode, out, 4); runG<32>(buf, bytes, 1, mode, out, 4); }
    runG<16>(buf, 65536u, 2, 0, out, 4); runG<16>(buf, 65536u, 4, 0, out, 4); runG<16>(buf, 65536u, 8, 0, out, 4); runG<16>(buf, 65536u, 32, 0, out, 4);
    runG<32>(buf, 65536u, 4, 0, out, 4);
    runG<16>(buf, 65536u, 1, 1, out, 8); runG<32>(buf, 65536u, 1, 1, out, 8);
    runS<4>(65536u, out, 3); runS<8>(65536u, out, 3); runS<16>(65536u, out, 3);
    return 0;
}

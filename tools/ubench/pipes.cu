// tools/ubench/pipes.cu — issue/pipe throughput of the instruction classes the traversal kernels are made of (B200, sm_100a).
// Each kernel runs ITER iterations of 8 independent chains x UNR ops of one class (or a 1:1 mix) on 148 x 8 CTAs of 128
// threads and reports warp-instructions per clock per SM (4 = one per scheduler per clock).  Measurement tooling only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
#define CHK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int OP>
__global__ void __launch_bounds__(128) k(uint32_t* out, uint32_t seed, float fa, float fb) {
    uint32_t r[8]; float f[8]; unsigned long long d[4];
#pragma unroll
    for (int i = 0; i < 8; i++) { r[i] = seed * (threadIdx.x + 1 + i); f[i] = float(threadIdx.x + i) * fa + 1.0f; }
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = (unsigned long long)(r[i]) << 32 | r[i + 4];
    uint32_t c = seed | 0x47800000u, sel = seed & 0x7777u;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
                if (OP == 1) { if (u & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) & 7])); else asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 3) & 7])); }
                if (OP == 2) { if (u & 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "f"(fb)); else asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 3) & 7]), "f"(fa)); }
                if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
                if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 3) & 7]));
                if (OP == 5) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 3) & 7]));
                if (OP == 6) { if (u & 1) asm volatile("shl.b32 %0, %1, 16;" : "=r"(r[i]) : "r"(r[(i + 1) & 7])); else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 3) & 7])); }
                if (OP == 7) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 3) & 7]));
                if (OP == 8) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(f[i]) : "f"(fa), "f"(fb));
                if (OP == 9) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
                if (OP == 10) { if (i < 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(d[(i + 1) & 3]), "l"(d[(i + 2) & 3])); }
                if (OP == 11) { if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); else asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(r[i]) : "r"(r[(i + 2) & 7])); }
                if (OP == 12) { if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); else asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 2) & 7]), "f"(fb)); }
                if (OP == 13) { if (i & 1) asm volatile("shl.b32 %0, %1, 16;" : "=r"(r[i]) : "r"(r[(i + 2) & 7])); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); }
                if (OP == 14) { if (i < 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(d[(i + 1) & 3]), "l"(d[(i + 2) & 3])); else asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 2) & 7]), "f"(fb)); }
                if (OP == 15) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa));
                if (OP == 16) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa));
                if (OP == 17) asm volatile("{.reg .pred p; setp.le.f32 p, %1, %2; @p or.b32 %0, %0, %3;}" : "+r"(r[i]) : "f"(f[i]), "f"(fa), "r"(sel));
                if (OP == 18) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(f[i]) : "f"(fa));
                if (OP == 19) { if (i % 3 == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb)); else if (i % 3 == 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 2) & 7]), "f"(fb)); else asm volatile("shl.b32 %0, %1, 16;" : "=r"(r[i]) : "r"(r[(i + 3) & 7])); }
                if (OP == 20) asm volatile("popc.b32 %0, %0;" : "+r"(r[i]));
                if (OP == 21) asm volatile("bfind.u32 %0, %0;" : "+r"(r[i]));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i] ^ __float_as_uint(f[i]);
#pragma unroll
    for (int i = 0; i < 4; i++) acc ^= uint32_t(d[i]) ^ uint32_t(d[i] >> 32);
    if (acc == 0x12345u) out[0] = acc;
}

template <int OP> int run(const char* name, int opsPerInner, uint32_t* out, int sms, double mhz) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<sms * 8, 128>>>(out, 3u, 1.0001f, 0.5f); CHK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a); k<OP><<<sms * 8, 128>>>(out, 3u, 1.0001f, 0.5f); cudaEventRecord(b); CHK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    double warpInst = double(sms) * 8 * 4 * double(ITER) * 4 * opsPerInner;
    double cyc = best * 1e-3 * mhz * 1e6;
    printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SM (at %.0f MHz)\n", name, best, warpInst / cyc / sms, mhz);
    return 0;
}

int main() {
    cudaDeviceProp p; CHK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); double mhz = khz / 1000.0;
    printf("%s, %d SMs, clockRate attr %.0f MHz (rates assume the GPU runs at this clock)\n", p.name, sms, mhz);
    uint32_t* out; CHK(cudaMalloc(&out, 4));
    run<0>("FFMA (3 reg)", 8, out, sms, mhz);
    run<18>("FFMA (imm)", 8, out, sms, mhz);
    run<15>("FMUL", 8, out, sms, mhz);
    run<16>("FADD", 8, out, sms, mhz);
    run<10>("FFMA2", 4, out, sms, mhz);
    run<1>("FMNMX", 8, out, sms, mhz);
    run<2>("FMNMX3", 8, out, sms, mhz);
    run<3>("PRMT", 8, out, sms, mhz);
    run<4>("LOP3", 8, out, sms, mhz);
    run<5>("SHF", 8, out, sms, mhz);
    run<6>("SHL:LOP3 1:1", 8, out, sms, mhz);
    run<7>("IMAD", 8, out, sms, mhz);
    run<9>("IADD", 8, out, sms, mhz);
    run<8>("FSETP+FSEL (2 inst)", 16, out, sms, mhz);
    run<17>("FSETP+@P LOP (2 inst)", 16, out, sms, mhz);
    run<20>("POPC", 8, out, sms, mhz);
    run<21>("FLO (bfind)", 8, out, sms, mhz);
    run<11>("FFMA:PRMT 1:1", 8, out, sms, mhz);
    run<12>("FFMA:FMNMX3 1:1", 8, out, sms, mhz);
    run<13>("FFMA:IMAD.SHL 1:1", 8, out, sms, mhz);
    run<14>("FFMA2:FMNMX3 1:1", 8, out, sms, mhz);
    run<19>("FFMA:FMNMX3:SHL 3:3:2", 8, out, sms, mhz);
    return 0;
}

// pipe_probe.cu -- measures issue/pipe rates that bound the CD kernels on this GPU (build: nvcc -arch=sm_100a).
// Prints lane-ops per clock per SM for: FFMA, FFMA2 (packed f32x2), FMNMX (ALU pipe), MUFU.RCP, and mixes.
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 2048
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
template <int MODE> __global__ void __launch_bounds__(256) probe(float* out) {
    float a[8]; unsigned long long p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); }
    const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
    const unsigned long long pm = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m), pc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = fmaf(a[i], m, c);
                if (MODE == 1) p[i] = ffma2(p[i], pm, pc);
                if (MODE == 2) a[i] = fmaxf(a[i] * 1.0f, m + i);          // FMUL? keep ALU: see below
                if (MODE == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                if (MODE == 4) { a[i] = fmaf(a[i], m, c); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[(i + 4) & 7]) : "f"(c)); }
                if (MODE == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
                if (MODE == 6) { p[i] = ffma2(p[i], pm, pc); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c)); }
                if (MODE == 7) { p[i] = ffma2(p[i], pm, pc); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c)); if ((i & 3) == 0) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
                if (MODE == 8) { asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %1; selp.f32 %0, %0, %1, q;}" : "+f"(a[i]) : "f"(c)); }
            }
        }
    }
    float r = 0; for (int i = 0; i < 8; ++i) r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    if (r == 123.456f) out[0] = r;
}
template <int MODE> void run(const char* name, double ops_per_inner, int sms, double mhz) {
    float* d; cudaMalloc(&d, 64); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 8; probe<MODE><<<blocks, 256>>>(d); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); probe<MODE><<<blocks, 256>>>(d); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    double inner = (double)ITERS * 4 * 8 * 256.0 * blocks;       // inner-statement executions (lane granularity)
    double per_clk_sm = inner * ops_per_inner / (best * 1e-3) / (mhz * 1e6) / sms;
    printf("%-44s %8.3f ms  %7.1f lane-instr/clk/SM (at %.0f MHz)  %6.2f G lane-instr/s\n", name, best, per_clk_sm, mhz, inner * ops_per_inner / (best * 1e-3) / 1e9);
    cudaFree(d);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount; double mhz = p.clockRate / 1e3;
    printf("%s, %d SMs, clockRate %.0f MHz\n", p.name, sms, mhz);
    run<0>("FFMA (3-reg)", 1, sms, mhz);
    run<1>("FFMA2 (f32x2, counts 1 instr)", 1, sms, mhz);
    run<5>("FMNMX (max.f32)", 1, sms, mhz);
    run<8>("FSETP+FSEL (2 instr)", 2, sms, mhz);
    run<3>("MUFU.RCP", 1, sms, mhz);
    run<4>("FFMA + FMNMX (2 instr)", 2, sms, mhz);
    run<6>("FFMA2 + FMNMX (2 instr)", 2, sms, mhz);
    run<7>("FFMA2 + FMNMX + 1/4 MUFU (2.25 instr)", 2.25, sms, mhz);
    return 0;
}

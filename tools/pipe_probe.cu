// Throw-away-ish probe of the sm_100a FP32 / SFU pipes: which packed forms buy what.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a),"f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a),"=f"(b) : "l"(v)); }

template<int MODE>
__global__ void __launch_bounds__(256) probe(float* out, float seed, long long* cyc)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = seed + threadIdx.x * 1e-3f + i;
    float a = seed * 0.999f, b = seed * 1e-3f;
    u64 p[8], pa = pk(a, a), pb = pk(b, b);
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = pk(x[2*i], x[2*i+1]);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {           // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 1) {    // 8 FFMA2 (=16 lane FMAs)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        } else if (MODE == 2) {    // 8 FADD2
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
        } else if (MODE == 3) {    // 16 MUFU.EX2
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        } else if (MODE == 4) {    // 8 FFMA2 + 2 MUFU (d=8-like ratio 16 lane-fma : 1 mufu per pair => 8 FFMA2 : 1 MUFU); here 8:2
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[0]));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[1]));
        } else if (MODE == 5) {    // 16 scalar FFMA + 2 MUFU
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
        } else if (MODE == 6) {    // 4 FFMA2 + 4 FADD2 + 2 MUFU + 8 FMNMX (ALU pipe co-issue test)
#pragma unroll
            for (int i = 0; i < 4; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
#pragma unroll
            for (int i = 4; i < 8; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[0]));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[1]));
#pragma unroll
            for (int i = 2; i < 10; i++) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
        } else if (MODE == 7) {    // 16 scalar FFMA + 16 FMNMX (is ALU pipe free?)
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        } else if (MODE == 8) {    // 8 FFMA2 + 8 FMNMX
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        } else if (MODE == 9) {    // 8 FFMA2 + 8 scalar FFMA (do packed and scalar share the pipe?)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
        }
    }
    long long t1 = clock64();
    float s = a + b;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { float u, v; upk(p[i], u, v); s += u + v; }
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int MODE> void run(const char* name, double lane_fma, double lane_mufu, double lane_alu, int warps_note)
{
    float* out; long long* cyc; int nb = 148 * 8;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, nb * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<nb, 256>>>(out, 1.0f, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) probe<MODE><<<nb, 256>>>(out, 1.0f, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    long long h[148*8]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double mc = 0; for (int i = 0; i < nb; i++) mc += h[i]; mc /= nb;
    double thr = (double)nb * 256 * ITERS;
    // per-SM per-clk rates from cycle counts: 8 CTAs resident/SM (2048 thr) all concurrently
    double per_sm_clk = 2048.0 * ITERS / mc;
    printf("%-34s ms=%8.4f  cyc/CTA=%9.0f  lane-fma/s=%.3e (%.1f/clk/SM)  mufu/s=%.3e (%.1f/clk/SM)  alu/s=%.3e (%.1f/clk/SM)  implied MHz=%.0f\n",
           name, ms, mc, thr * lane_fma / (ms * 1e-3), per_sm_clk * lane_fma,
           thr * lane_mufu / (ms * 1e-3), per_sm_clk * lane_mufu, thr * lane_alu / (ms * 1e-3), per_sm_clk * lane_alu,
           mc / (ms * 1e-3) / 1e6);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("device %s SMs=%d clock=%d kHz\n", pr.name, pr.multiProcessorCount, pr.clockRate);
    run<0>("16xFFMA", 16, 0, 0, 0);
    run<1>("8xFFMA2", 16, 0, 0, 0);
    run<2>("8xFADD2", 16, 0, 0, 0);
    run<3>("16xMUFU.EX2", 0, 16, 0, 0);
    run<4>("8xFFMA2+2xMUFU", 16, 2, 0, 0);
    run<5>("16xFFMA+2xMUFU", 16, 2, 0, 0);
    run<6>("4xFFMA2+4xFADD2+2MUFU+8FMNMX", 16, 2, 8, 0);
    run<7>("16xFFMA+16xFMNMX", 16, 0, 16, 0);
    run<8>("8xFFMA2+8xFMNMX", 16, 0, 8, 0);
    run<9>("8xFFMA2+8xFFMA", 24, 0, 0, 0);
    return 0;
}

// Probe: the score kernel's inner loop with operands in registers only (no LDS/TMA), to see what
// limits FFMA2 + MUFU co-issue: d FADD2 + d FFMA2 chain -> 2 MUFU -> 1 FFMA2 per packed pair.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a),"f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a),"=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b){ u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a),"l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a),"l"(b),"l"(c)); return r; }
__device__ __forceinline__ float ex2n(float a){ float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-a)); return r; }

template<int D, int Q, int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) loopk(float* out, const float* in, int iters)
{
    u64 q2[Q][D], nm2[Q], sum2[Q];
    for (int q = 0; q < Q; q++) { nm2[q] = pk(-3.f, -3.f); sum2[q] = pk(0.f, 0.f);
        for (int k = 0; k < D; k++) { float v = in[(threadIdx.x + q * 7 + k) & 63]; q2[q][k] = pk(v, v); } }
    u64 xa[D], xb[D];
    for (int k = 0; k < D; k++) { xa[k] = pk(in[k], in[k + 1]); xb[k] = pk(in[k + 2], in[k + 3]); }
    u64 wa = pk(in[5], in[6]), wb = pk(in[7], in[8]);
    const u64 bump = pk(in[9] * 1e-6f, in[10] * 1e-6f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < D; k++) { xa[k] = sub2(xa[k], bump); xb[k] = sub2(xb[k], bump); }   // stands in for the LDS
#pragma unroll
        for (int q = 0; q < Q; q++) {
            u64 aa = nm2[q], ab = nm2[q];
#pragma unroll
            for (int k = 0; k < D; k++) {
                u64 da = sub2(q2[q][k], xa[k]); u64 db = sub2(q2[q][k], xb[k]);
                aa = fma2(da, da, aa); ab = fma2(db, db, ab);
            }
            float a0, a1, b0, b1; upk(aa, a0, a1); upk(ab, b0, b1);
            if (MODE == 0) {            // as in the kernel
                u64 ea = pk(ex2n(a0), ex2n(a1)), eb = pk(ex2n(b0), ex2n(b1));
                sum2[q] = fma2(wa, ea, sum2[q]); sum2[q] = fma2(wb, eb, sum2[q]);
            } else if (MODE == 1) {     // no MUFU: terms = acc (FMA work only)
                sum2[q] = fma2(wa, aa, sum2[q]); sum2[q] = fma2(wb, ab, sum2[q]);
            } else if (MODE == 2) {     // MUFU, two independent accumulators per query (shorter dep chain on sum)
                u64 ea = pk(ex2n(a0), ex2n(a1)), eb = pk(ex2n(b0), ex2n(b1));
                sum2[q] = fma2(wa, ea, sum2[q]); nm2[q] = fma2(wb, eb, nm2[q]);
            }
        }
    }
    float s = 0; for (int q = 0; q < Q; q++) { float a, b; upk(sum2[q], a, b); s += a + b; upk(nm2[q], a, b); s += a + b; }
    if (s == 1.2345f) out[0] = s;
}
template<int D, int Q, int MODE, int MINB> void run(const char* name, float* out, float* in)
{
    int iters = 4096; int nb = 148 * 8;
    loopk<D,Q,MODE,MINB><<<nb,128>>>(out, in, 16); cudaDeviceSynchronize();
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); for (int r=0;r<3;r++) loopk<D,Q,MODE,MINB><<<nb,128>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=3;
    double pairs = (double)nb*128*iters*Q*4;
    double lane = pairs*(2*D+1) + (double)nb*128*iters*D*4;  // + the stand-in bumps (2D packed = 4D lane ops per iter)
    int occ=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, loopk<D,Q,MODE,MINB>, 128, 0);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, loopk<D,Q,MODE,MINB>);
    printf("%-28s regs=%3d occ=%d  %.3f ms  pairs/s=%.3e  fma-lane/s=%.3e (%.1f%% of 3.674e13)  ex2/s=%.3e (%.1f%% of 4.63e12)\n", name, fa.numRegs, occ, ms,
           pairs/ms*1e3, lane/ms*1e3, 100*lane/ms*1e3/3.674e13, (MODE==1?0:pairs)/ms*1e3, 100*(MODE==1?0:pairs)/ms*1e3/4.63e12);
}
int main(){ float *out,*in; cudaMalloc(&out,4); cudaMalloc(&in,256); float h[64]; for(int i=0;i<64;i++) h[i]=0.01f*i+0.3f; cudaMemcpy(in,h,256,cudaMemcpyHostToDevice);
  run<4,8,0,2>("d4 Q8 mufu mb2", out,in); run<4,8,1,2>("d4 Q8 nomufu mb2", out,in); run<4,8,2,2>("d4 Q8 mufu 2acc mb2", out,in);
  run<4,4,0,3>("d4 Q4 mufu mb3", out,in); run<4,4,0,4>("d4 Q4 mufu mb4", out,in); run<4,8,0,3>("d4 Q8 mufu mb3", out,in);
  run<8,4,0,3>("d8 Q4 mufu mb3", out,in); run<8,4,1,3>("d8 Q4 nomufu mb3", out,in); run<8,4,2,3>("d8 Q4 mufu 2acc mb3", out,in);
  run<2,8,0,4>("d2 Q8 mufu mb4", out,in); run<2,8,1,4>("d2 Q8 nomufu mb4", out,in);
  run<16,2,0,4>("d16 Q2 mufu mb4", out,in); run<16,2,1,4>("d16 Q2 nomufu mb4", out,in);
  return 0; }

// tcgen05 probe for the d = 16 "-2 X Y^T" distance term (BASELINE config 4 / north_star N1).  NOT part of libskb: a
// measurement tool.  It answers two questions on hardware: how fast is an all-pairs Gaussian-KDE evaluation when the
// 16-term dot product runs on the 5th-generation tensor cores, and how accurate is it.
//
//   dist^2(q, x) = |q - c|^2 + (|u|^2 + 2 c.u) - 2 q.u          x = c + u,  c = centre of x's tile (kd leaf of 64 points)
//
//   * q.u is a [128 queries] x [64 points] x [K] product on tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate in
//     TMEM).  fp32-level accuracy comes from a bf16 x 3 split of both operands, q = h + m + l, u = H + M + L, laid out
//     along K as  A = [h h m h l m],  B = [H M H L H M]  (K = 6 x 16 = 96: hH + hM + mH + hL + lH + mM; the dropped
//     products are <= 2^-24 relative);
//   * operands sit in shared memory in the canonical K-major, no-swizzle layout (core matrix = 8 rows x 16 bytes);
//     A is built once per CTA from its 128 queries, the B tiles are pre-split on the host and arrive by 1-D TMA;
//   * |q - c|^2 is formed per (query, tile) in fp64 and carried as a two-float (hi, lo), e_j = |u_j|^2 + 2 c.u_j as a
//     two-float from the host: the three terms of dist^2 are each ~10^3 while their sum is ~10^1..10^2;
//   * epilogue per pair: FFMA (2 dot - e_hi), FADD (+ G_hi), FADD (+ lo parts), MUFU.EX2, FFMA (w_j x term).
//
// One CTA = 128 threads = 128 queries (one TMEM lane each) against ALL tiles; MMA(t + 1) runs while the CTA does the
// epilogue of tile t (two accumulators in TMEM, two tile buffers in shared memory).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace {

constexpr int D = 16;
constexpr int TN = 64;                  // points per tile
constexpr int KK = 96;                  // K extent (bf16 elements): 6 split products x 16 dims
constexpr int A_BYTES = 128 * KK * 2;   // 24576
constexpr int B_BYTES = TN * KK * 2;    // 12288
// record layout in global memory (16-byte multiples): [B 12288][e_hi 256][e_lo 256][w 256][c 64] = 13120 bytes
constexpr int REC = B_BYTES + 256 + 256 + 256 + 64;
static_assert(REC % 16 == 0, "bulk copy size");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\nW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\nD_%=:\n\t}"
                 :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c = F32, a = b = BF16, both K-major, N = 64, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint16_t bf16_rn(float f) {
    uint32_t u = __float_as_uint(f);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
__device__ __forceinline__ float bf16_f(uint16_t h) { return __uint_as_float((uint32_t)h << 16); }

#define LD32(r, taddr)                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"          \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),           \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),     \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),   \
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])    \
                 : "r"(taddr))

__global__ void __launch_bounds__(128)
tc_probe_kernel(const float* __restrict__ q, const float* __restrict__ negm, int M,
                const unsigned char* __restrict__ recs, int ntiles, double* __restrict__ outS)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;                                   // 24576
    unsigned char* sB = smem + A_BYTES;                         // 2 x REC (tile record: B + constants)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + A_BYTES + 2 * REC);     // [2]
    uint64_t* mmad = full + 2;                                  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mmad + 2);

    const int tid = threadIdx.x, wid = tid >> 5;
    const long long row = (long long)blockIdx.x * 128 + tid;
    const long long rr = row < M ? row : M - 1;

    if (tid == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        mbar_init(&mmad[0], 1); mbar_init(&mmad[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (wid == 0) {                                             // one warp allocates 128 TMEM columns (two 64-column accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // ---- this thread's query: registers + its row of A (bf16 x 3 split, [h h m h l m] along K) ----
    float qv[D];
#pragma unroll
    for (int k = 0; k < D; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(q + rr * D + k);
        qv[k] = v.x; qv[k + 1] = v.y; qv[k + 2] = v.z; qv[k + 3] = v.w;
    }
    const float nm = negm[rr];
    {
        uint16_t h[D], m[D], l[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            h[k] = bf16_rn(qv[k]);
            const float r1 = qv[k] - bf16_f(h[k]);
            m[k] = bf16_rn(r1);
            l[k] = bf16_rn(r1 - bf16_f(m[k]));
        }
        const uint16_t* parts[6] = {h, h, m, h, l, m};
        unsigned char* rowbase = sA + (tid >> 3) * 128 + (tid & 7) * 16;
#pragma unroll
        for (int p = 0; p < 6; p++)
#pragma unroll
            for (int c = 0; c < 2; c++) {                       // two 16-byte chunks (8 bf16) per 16-dim part
                const uint16_t* s = parts[p] + c * 8;
                uint4 v;
                v.x = s[0] | ((uint32_t)s[1] << 16); v.y = s[2] | ((uint32_t)s[3] << 16);
                v.z = s[4] | ((uint32_t)s[5] << 16); v.w = s[6] | ((uint32_t)s[7] << 16);
                *reinterpret_cast<uint4*>(rowbase + (p * 2 + c) * 2048) = v;
            }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // A was written with generic stores, the MMA reads it through the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    auto issue_tma = [&](int t) {
        const int b = t & 1;
        mbar_expect_tx(&full[b], REC);
        tma_bulk_g2s(sB + b * REC, recs + (size_t)t * REC, REC, &full[b]);
    };
    auto issue_mma = [&](int t) {                               // thread 0 only
        const int b = t & 1;
        mbar_wait(&full[b], (t >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + b * REC);
#pragma unroll
        for (int s = 0; s < 6; s++)
            mma_bf16(tmem + (uint32_t)b * 64u, smem_desc(a0 + s * 4096, 2048, 128), smem_desc(b0 + s * 2048, 1024, 128), s > 0);
        mma_commit(&mmad[b]);
    };
    if (tid == 0) {
        issue_tma(0);
        if (ntiles > 1) issue_tma(1);
        issue_mma(0);
    }
    double S = 0.0;
#pragma unroll 1
    for (int t = 0; t < ntiles; t++) {
        const int b = t & 1;
        if (tid == 0 && t + 1 < ntiles) issue_mma(t + 1);       // runs on the tensor core during this tile's epilogue
        __syncwarp();
        // every thread needs the tile's constants (TMA) and its accumulator row (MMA)
        mbar_wait(&full[b], (t >> 1) & 1);
        const float* cst = reinterpret_cast<const float*>(sB + b * REC + B_BYTES);   // e_hi[64] e_lo[64] w[64] c[16]
        double g = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) {
            const double dz = (double)qv[k] - (double)cst[192 + k];
            g = fma(dz, dz, g);
        }
        const double G = -(double)nm - g;                        // m - |q - c|^2
        const float Ghi = (float)G, Glo = (float)(G - (double)Ghi);
        mbar_wait(&mmad[b], (t >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float sum = 0.f;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            uint32_t r[32];
            const uint32_t taddr = tmem + (((uint32_t)wid * 32u) << 16) + (uint32_t)b * 64u + (uint32_t)half * 32u;
            LD32(r, taddr);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 eh = *reinterpret_cast<const float4*>(cst + half * 32 + j);
                const float4 el = *reinterpret_cast<const float4*>(cst + 64 + half * 32 + j);
                const float4 w4 = *reinterpret_cast<const float4*>(cst + 128 + half * 32 + j);
                const float ehh[4] = {eh.x, eh.y, eh.z, eh.w}, ell[4] = {el.x, el.y, el.z, el.w}, ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float dot = __uint_as_float(r[j + i]);
                    float arg = __fadd_rn(fmaf(2.f, dot, -ehh[i]), Ghi);       // (2 q.u - e_hi) + G_hi: the large parts cancel here
                    arg = __fadd_rn(arg, Glo - ell[i]);
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(arg));
                    sum = fmaf(ww[i], e, sum);
                }
            }
        }
        S += (double)sum;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                         // buffer b (shared memory + accumulator) is free again
        if (tid == 0 && t + 2 < ntiles) issue_tma(t + 2);
    }
    if (row < M) outS[row] = S;
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u));
}

}  // namespace

// all pointers are device pointers; returns the average kernel time over `reps` launches in *ms
extern "C" int tc_probe_run(const float* q, const float* negm, int M, const unsigned char* recs, int ntiles,
                            double* outS, int reps, float* ms)
{
    const size_t smem = A_BYTES + 2 * REC + 64;
    cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { fprintf(stderr, "attr: %s\n", cudaGetErrorString(e)); return 1; }
    const int grid = (M + 127) / 128;
    tc_probe_kernel<<<grid, 128, smem>>>(q, negm, M, recs, ntiles, outS);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) { fprintf(stderr, "run: %s\n", cudaGetErrorString(e)); return 2; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) tc_probe_kernel<<<grid, 128, smem>>>(q, negm, M, recs, ntiles, outS);
    cudaEventRecord(e1);
    if ((e = cudaEventSynchronize(e1)) != cudaSuccess) { fprintf(stderr, "time: %s\n", cudaGetErrorString(e)); return 3; }
    cudaEventElapsedTime(ms, e0, e1);
    *ms /= reps;
    return 0;
}

extern "C" int tc_probe_record_bytes(void) { return REC; }

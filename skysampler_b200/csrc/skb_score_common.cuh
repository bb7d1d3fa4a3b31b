// Device helpers shared by the score kernels (skb_score.cu: 1-D far-field, lanes = queries and per-query-culling kernels;
// skb_sparse.cu: the sorted-incidence pipeline): packed fp32x2 PTX, mbarrier / TMA wrappers, the pair-term inner step, query
// loading with the fused fp64 whitening, box loads, the exponent-reference search (per lane and warp cooperative) and the
// finalize epilogue.
#pragma once
#include "skb_internal.h"
#include <math.h>
#include <type_traits>

namespace skb {

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ float ex2_neg(float a) {   // 2^(-a), MUFU.EX2 with the negate modifier
    float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-a)); return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n"
        "D_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// diagnostics, only when a launch carries a counter block (ScoreJob::stats; skb_set_option("stats", 1)) — warp level:
// [0] (query, point) pairs per lane re-run after an exponent-window overflow, [1] (query, point) pairs per lane
// evaluated (x 32 lanes = pairs evaluated), [2] tiles culled by the box tests, [3] tiles in range
constexpr float OVERFLOW_LIMIT = 1.329228e36f;    // 2^120: a block partial sum at/above this is re-run
constexpr float LW_CAP = 64.f;                    // a weight lighter than 2^-64 w_max moves the reference by at most this
constexpr int SUB_MAX = 64;                       // points per overflow-checked block (at most one tile)
constexpr float CULL_MARGIN = 128.f;              // box lower bound this far past the reference => all terms flush to 0

// ---------------------------------------------------------------------------------------------
// One pass of a thread's Q queries over (part of) a shared-memory tile.
//   MIN_ONLY = false:  sum2[q] += w_j * 2^(m_q - acc_j)      (packed halves = even / odd points)
//   MIN_ONLY = true :  amin[q]  = min(amin[q], acc_j - m_q)  (no SFU work)
// ---------------------------------------------------------------------------------------------
#ifndef SKB_POLY_EXP
#define SKB_POLY_EXP 1     // 1: for D == 1 a quarter of the exponentials run as a polynomial on the FMA pipe
#endif
#ifndef SKB_STEP8
#define SKB_STEP8 1       // 1: 8 training points per inner step when a thread holds <= 2 queries
#endif
#ifndef SKB_HOME_SUPER
#define SKB_HOME_SUPER 1   // 1: for D >= 5 the exponent reference comes from an exact scan of the whole home super-box
#endif
#ifndef SKB_PREFETCH
#define SKB_PREFETCH 0     // 1: double-buffer the tile words of the next 4 points in registers
#endif

// the training-side operands of one step: 4 points as two packed pairs per dimension (+ weights)
template <int D>
struct StepOperands {
    static constexpr int TN = tile_points(D);
    u64 xa[D], xb[D], wa, wb;
    template <bool WITH_W>
    __device__ __forceinline__ void load(const float* __restrict__ tile, int j)
    {
#pragma unroll
        for (int k = 0; k < D; k++) {
            const float4 v = *reinterpret_cast<const float4*>(tile + k * TN + j);
            xa[k] = pk(v.x, v.y);
            xb[k] = pk(v.z, v.w);
        }
        {
            const float4 w4 = *reinterpret_cast<const float4*>(tile + D * TN + j);
            wa = pk(w4.x, w4.y);
            wb = pk(w4.z, w4.w);
        }
    }
};

// {2^-a.lo, 2^-a.hi} on the FMA pipe: Cody-Waite split with the 1.5*2^23 magic constant and a degree-5
// minimax polynomial of 2^f on [-0.5, 0.5] (max relative error 2.3e-7 in fp32 Horner form, the same
// order as MUFU.EX2's), scaled by adding n << 23 to the exponent field.  8 packed FMA-pipe ops per two
// terms (+ 2 IMAD, 4 FMNMX).  Used for a fixed quarter of the terms when D == 1, where the SFU is the
// binding pipe and the FMA pipe is 3/8 busy: measured +4.5 % (4.45e12 -> 4.66e12 pairs/s on B200); at D == 2
// the same offload LOSES 13 % (the FMA pipe becomes the limiter), so it is not applied there.  The argument is clamped to [-126, 250] and the
// result's bit pattern to >= 0, so that far points (a >= 127) give exactly 0 and an exponent-window overflow still reads as >= 2^120.
// Q independent polynomial exponentials, written stage-major so that the Q dependent chains interleave
// (each stage is one packed FMA-pipe op per query; a single chain is 8 ops deep).
template <int Q>
__device__ __forceinline__ void exp2_neg_poly2_batch(u64 (&v)[Q])
{
    const u64 magic = pk(12582912.f, 12582912.f);
    u64 t[Q], g[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        float a0, a1;
        upk(v[q], a0, a1);
        a0 = fminf(fmaxf(a0, -126.f), 250.f);
        a1 = fminf(fmaxf(a1, -126.f), 250.f);
        v[q] = pk(a0, a1);
    }
#pragma unroll
    for (int q = 0; q < Q; q++) t[q] = sub2(magic, v[q]);
#pragma unroll
    for (int q = 0; q < Q; q++) g[q] = sub2(t[q], magic);
#pragma unroll
    for (int q = 0; q < Q; q++) asm("add.rn.f32x2 %0, %1, %2;" : "=l"(g[q]) : "l"(v[q]), "l"(g[q]));
#pragma unroll
    for (int q = 0; q < Q; q++) v[q] = fma2(pk(-1.3276470126584172e-3f, -1.3276470126584172e-3f), g[q],
                                            pk(9.675540961325169e-3f, 9.675540961325169e-3f));
#pragma unroll
    for (int q = 0; q < Q; q++) v[q] = fma2(v[q], g[q], pk(-5.550713464617729e-2f, -5.550713464617729e-2f));
#pragma unroll
    for (int q = 0; q < Q; q++) v[q] = fma2(v[q], g[q], pk(2.4022120237350464e-1f, 2.4022120237350464e-1f));
#pragma unroll
    for (int q = 0; q < Q; q++) v[q] = fma2(v[q], g[q], pk(-6.931469440460205e-1f, -6.931469440460205e-1f));
#pragma unroll
    for (int q = 0; q < Q; q++) v[q] = fma2(v[q], g[q], pk(1.0000001192092896f, 1.0000001192092896f));
#pragma unroll
    for (int q = 0; q < Q; q++) {
        float p0, p1, t0, t1;
        upk(v[q], p0, p1);
        upk(t[q], t0, t1);
        // a >= 127 drives the exponent field to <= 0: the sum goes negative as an integer and is clamped to
        // +0.0 — far terms must be EXACTLY zero, like the flushed MUFU terms, or a query's sum would depend on
        // which culled tiles its warp companions keep alive
        v[q] = pk(__int_as_float(max(__float_as_int(p0) + (__float_as_int(t0) << 23), 0)),
                  __int_as_float(max(__float_as_int(p1) + (__float_as_int(t1) << 23), 0)));
    }
}

template <int D, int Q, bool MIN_ONLY, bool POLY_B = false>
__device__ __forceinline__ void step_compute(const StepOperands<D>& op, const float (&q2)[Q][D], const float (&nm2)[Q],
                                             u64 (&sum2)[Q], float (&amin)[Q])
{
    u64 pb[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const u64 seed = pk(nm2[q], nm2[q]);
        u64 aa = seed, ab = seed;
#pragma unroll
        for (int k = 0; k < D; k++) {
            const u64 qk = pk(q2[q][k], q2[q][k]);     // folds into the scalar-broadcast operand form (R.F32)
            const u64 da = sub2(qk, op.xa[k]);
            const u64 db = sub2(qk, op.xb[k]);
            aa = fma2(da, da, aa);
            ab = fma2(db, db, ab);
        }
        float a0, a1, b0, b1;
        upk(aa, a0, a1);
        upk(ab, b0, b1);
        if (MIN_ONLY) {
            // weighted exponents a' = acc - log2 w (rare path: re-centring after an exponent-window overflow)
            float w0, w1, w2, w3;
            upk(op.wa, w0, w1);
            upk(op.wb, w2, w3);
            amin[q] = min3(amin[q], a0 + fminf(-__log2f(w0), LW_CAP), a1 + fminf(-__log2f(w1), LW_CAP));
            amin[q] = min3(amin[q], b0 + fminf(-__log2f(w2), LW_CAP), b1 + fminf(-__log2f(w3), LW_CAP));
        } else {
            const u64 ea = pk(ex2_neg(a0), ex2_neg(a1));
            sum2[q] = fma2(op.wa, ea, sum2[q]);
            if (POLY_B) {
                pb[q] = ab;
            } else {
                const u64 eb = pk(ex2_neg(b0), ex2_neg(b1));
                sum2[q] = fma2(op.wb, eb, sum2[q]);
            }
        }
    }
    if (POLY_B && !MIN_ONLY) {
        exp2_neg_poly2_batch<Q>(pb);
#pragma unroll
        for (int q = 0; q < Q; q++) sum2[q] = fma2(op.wb, pb[q], sum2[q]);
    }
}

// 8 points per step (four packed pairs per dimension): with Q <= 2 queries per thread a 4-point step has
// only 2Q independent FMA chains in flight, too few to cover the FFMA2 latency; this doubles them.
template <int D, int Q, bool MIN_ONLY>
__device__ __forceinline__ void step_compute8(const float* __restrict__ tile, int j, const float (&q2)[Q][D],
                                              const float (&nm2)[Q], u64 (&sum2)[Q], float (&amin)[Q])
{
    constexpr int TN = tile_points(D);
    u64 xa[D], xb[D], xc[D], xd[D];
#pragma unroll
    for (int k = 0; k < D; k++) {
        const float4 v = *reinterpret_cast<const float4*>(tile + k * TN + j);
        const float4 u = *reinterpret_cast<const float4*>(tile + k * TN + j + 4);
        xa[k] = pk(v.x, v.y); xb[k] = pk(v.z, v.w); xc[k] = pk(u.x, u.y); xd[k] = pk(u.z, u.w);
    }
    u64 wa = 0, wb = 0, wc = 0, wd = 0;
    {
        const float4 w4 = *reinterpret_cast<const float4*>(tile + D * TN + j);
        const float4 w5 = *reinterpret_cast<const float4*>(tile + D * TN + j + 4);
        wa = pk(w4.x, w4.y); wb = pk(w4.z, w4.w); wc = pk(w5.x, w5.y); wd = pk(w5.z, w5.w);
    }
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const u64 seed = pk(nm2[q], nm2[q]);
        u64 aa = seed, ab = seed, ac = seed, ad = seed;
#pragma unroll
        for (int k = 0; k < D; k++) {
            const u64 qk = pk(q2[q][k], q2[q][k]);
            const u64 da = sub2(qk, xa[k]);
            const u64 db = sub2(qk, xb[k]);
            const u64 dc = sub2(qk, xc[k]);
            const u64 dd = sub2(qk, xd[k]);
            aa = fma2(da, da, aa);
            ab = fma2(db, db, ab);
            ac = fma2(dc, dc, ac);
            ad = fma2(dd, dd, ad);
        }
        float a0, a1, b0, b1, c0, c1, d0, d1;
        upk(aa, a0, a1); upk(ab, b0, b1); upk(ac, c0, c1); upk(ad, d0, d1);
        if (MIN_ONLY) {
            float w[8];
            upk(wa, w[0], w[1]); upk(wb, w[2], w[3]); upk(wc, w[4], w[5]); upk(wd, w[6], w[7]);
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = fminf(-__log2f(w[i]), LW_CAP);
            amin[q] = min3(amin[q], a0 + w[0], a1 + w[1]);
            amin[q] = min3(amin[q], b0 + w[2], b1 + w[3]);
            amin[q] = min3(amin[q], c0 + w[4], c1 + w[5]);
            amin[q] = min3(amin[q], d0 + w[6], d1 + w[7]);
        } else {
            // same accumulation order as two 4-point steps: (a, b) then (c, d)
            sum2[q] = fma2(wa, pk(ex2_neg(a0), ex2_neg(a1)), sum2[q]);
            sum2[q] = fma2(wb, pk(ex2_neg(b0), ex2_neg(b1)), sum2[q]);
            sum2[q] = fma2(wc, pk(ex2_neg(c0), ex2_neg(c1)), sum2[q]);
            sum2[q] = fma2(wd, pk(ex2_neg(d0), ex2_neg(d1)), sum2[q]);
        }
    }
}

// npts must be a multiple of 8.
template <int D, int Q, bool MIN_ONLY>
__device__ __forceinline__ void tile_pass(const float* __restrict__ tile, int npts,
                                          const float (&q2)[Q][D], const float (&nm2)[Q],
                                          u64 (&sum2)[Q], float (&amin)[Q])
{
#if SKB_PREFETCH
    StepOperands<D> A, B;
    A.template load<!MIN_ONLY>(tile, 0);
#pragma unroll 1
    for (int j = 0; j < npts; j += 8) {
        B.template load<!MIN_ONLY>(tile, j + 4);
        step_compute<D, Q, MIN_ONLY>(A, q2, nm2, sum2, amin);
        const int jn = (j + 8 < npts) ? j + 8 : j;          // last iteration: harmless re-load
        A.template load<!MIN_ONLY>(tile, jn);
        step_compute<D, Q, MIN_ONLY>(B, q2, nm2, sum2, amin);
    }
#else
    // which exponentials go to the FMA pipe depends on the point index only (points 8i+6, 8i+7 of a
    // tile), never on the query slot, so a query's arithmetic stays independent of its batch position
    constexpr bool kPoly = SKB_POLY_EXP && (D == 1) && !MIN_ONLY;
    if (kPoly) {
#pragma unroll 1
        for (int j = 0; j < npts; j += 8) {
            StepOperands<D> A;
            A.template load<!MIN_ONLY>(tile, j);
            step_compute<D, Q, MIN_ONLY, false>(A, q2, nm2, sum2, amin);
            A.template load<!MIN_ONLY>(tile, j + 4);
            step_compute<D, Q, MIN_ONLY, kPoly>(A, q2, nm2, sum2, amin);
        }
    } else if (SKB_STEP8 && Q * D <= 16) {
#pragma unroll 1
        for (int j = 0; j < npts; j += 8) step_compute8<D, Q, MIN_ONLY>(tile, j, q2, nm2, sum2, amin);
    } else {
#pragma unroll 1
        for (int j = 0; j < npts; j += 4) {
            StepOperands<D> A;
            A.template load<!MIN_ONLY>(tile, j);
            step_compute<D, Q, MIN_ONLY, false>(A, q2, nm2, sum2, amin);
        }
    }
#endif
}

// log p (natural) of query r from the per-split partials of one job.
__device__ __forceinline__ void merge_partials(const double2* __restrict__ partial, int nsplit,
                                               long long M, long long r, double& mstar, double& S)
{
    mstar = INFINITY;
    bool nan_in = false;
    for (int s = 0; s < nsplit; s++) {
        const double2 p = partial[(long long)s * M + r];
        if (p.y > 0.0) mstar = fmin(mstar, p.x);          // weightless splits (S == 0) carry no reference
        nan_in |= (p.y != p.y);
    }
    S = 0.0;
    if (nan_in) { S = NAN; mstar = 0.0; return; }
    if (!(mstar < INFINITY)) { mstar = 0.0; return; }
    for (int s = 0; s < nsplit; s++) {
        const double2 p = partial[(long long)s * M + r];
        if (!(p.y > 0.0)) continue;
        double e = mstar - p.x;                 // <= 0, integer valued
        if (!(e > -4000.0)) e = -4000.0;
        S += ldexp(p.y, (int)e);
    }
}

constexpr double LN2 = 0.693147180559945309417232121458;

// one rounding sequence for every path that turns (m, S) into log p (no FMA contraction), so the register
// epilogue and the partial-state epilogue give the same bits
__device__ __forceinline__ double final_logp(double log_norm, double S, double m)
{
    return __dadd_rn(__dadd_rn(log_norm, log(S)), -__dmul_rn(m, LN2));
}

// get(f, v, mstar, S): the (m, S) state of job f for the query at visit position v
template <class Get>
__device__ __forceinline__ void finalize_query_with(const FinTable& fin, int first, int count, long long v, Get get)
{
    // v: visit-order position (states);  r: row of the query table (outputs, uniforms)
    const int* perm = fin.fj[first].perm;
    const long long r = perm ? perm[v] : v;
    double log_ratio = -fin.log_m;
    for (int f = first; f < first + count; f++) {
        const FinJob& fj = fin.fj[f];
        double mstar, S;
        get(f, v, mstar, S);
        if (fj.out_m) { fj.out_m[r] = fj.log_wmax - mstar * LN2; fj.out_s[r] = S; }
        const double lp = final_logp(fj.log_norm, S, mstar);
        if (fj.out_logp) fj.out_logp[r] = lp;
        log_ratio += fj.sign * lp + fj.sign_logjac;
    }
    if (fin.ratio) {
        if (fin.out_log_ratio) fin.out_log_ratio[r] = log_ratio;
        if (fin.u) {
            const double lu = log(fin.u[r]);
            if (fin.accept) fin.accept[r] = (lu < log_ratio) ? 1 : 0;
            if (fin.in_band) fin.in_band[r] = (fabs(log_ratio - lu) <= fin.band) ? 1 : 0;
        }
    }
}

__device__ __forceinline__ void finalize_query(const FinTable& fin, int first, int count,
                                               long long M, long long v)
{
    finalize_query_with(fin, first, count, v, [&](int f, long long vv, double& mstar, double& S) {
        merge_partials(fin.fj[f].partial, fin.fj[f].nsplit, M, vv, mstar, S);
    });
}

// ---- load this thread's Q query rows, whiten + prescale in fp64, pack as {z, z} ----
template <int D, int Q>
__device__ __forceinline__ void load_queries(const ScoreJob& job, const KdeDev* __restrict__ kde, long long qbase, int tid,
                                             long long M, float (&q2)[Q][D])
{
    {
        const double c = kde->c;
#pragma unroll
        for (int q = 0; q < Q; q++) {
            long long r = qbase + (long long)tid * Q + q;   // a warp owns 32*Q consecutive rows of the visit order
            if (r >= M) r = M - 1;                      // clamp: duplicates are computed, never stored
            if (job.perm) r = job.perm[r];
            double x[D];
            if (job.q_f32) {
                const float* row = reinterpret_cast<const float*>(job.q) + r * job.ldq;
                if (job.cols_identity && (D % 4 == 0)) {
#pragma unroll
                    for (int i = 0; i < D; i += 4) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(row + i));
                        x[i] = v.x; x[i + 1] = v.y; x[i + 2] = v.z; x[i + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < D; i++) x[i] = __ldg(row + job.cols[i]);
                }
            } else {
                const double* row = reinterpret_cast<const double*>(job.q) + r * job.ldq;
                if (job.cols_identity && (D % 2 == 0)) {
#pragma unroll
                    for (int i = 0; i < D; i += 2) {
                        const double2 v = __ldg(reinterpret_cast<const double2*>(row + i));
                        x[i] = v.x; x[i + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < D; i++) x[i] = __ldg(row + job.cols[i]);
                }
            }
            if (job.q_is_whitened) {
#pragma unroll
                for (int k = 0; k < D; k++) q2[q][k] = (float)(x[k] * c);
            } else {
#pragma unroll
                for (int i = 0; i < D; i++) x[i] -= kde->mu[i];
#pragma unroll
                for (int k = 0; k < D; k++) {
                    double z = 0.0;
#pragma unroll
                    for (int i = 0; i < D; i++) z = fma(x[i], kde->Wc[i * D + k], z);
                    q2[q][k] = (float)z;
                }
            }
        }
    }
}

// min over a tile's points of the WEIGHTED exponent a'_j = |c dz_j|^2 - log2(w_j / w_max)  (the term is 2^-a'), straight
// from L2.  The window must hang on the largest weighted term: referencing it on the nearest POINT lets a light
// nearest neighbour push heavy points a little farther out below the flush threshold.  A weight's shift is capped at
// LW_CAP binades so that no finite weight can move a term past the fp32 exponent range.
template <int D>
__device__ __forceinline__ float scan_tile_min(const float* __restrict__ tile, const float (&zq)[D], bool has_w, float best)
{
    constexpr int TN = tile_points(D);
#pragma unroll 1
    for (int j = 0; j < TN; j += 4) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (has_w) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(tile + D * TN + j));
            a0 = fminf(-__log2f(w.x), LW_CAP); a1 = fminf(-__log2f(w.y), LW_CAP);
            a2 = fminf(-__log2f(w.z), LW_CAP); a3 = fminf(-__log2f(w.w), LW_CAP);
        }
#pragma unroll
        for (int k = 0; k < D; k++) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(tile + k * TN + j));
            const float d0 = zq[k] - v.x, d1 = zq[k] - v.y, d2 = zq[k] - v.z, d3 = zq[k] - v.w;
            a0 = fmaf(d0, d0, a0); a1 = fmaf(d1, d1, a1); a2 = fmaf(d2, d2, a2); a3 = fmaf(d3, d3, a3);
        }
        best = fminf(best, fminf(fminf(a0, a1), fminf(a2, a3)));
    }
    return best;
}

// one bounding box (lo[D], hi[D], 2 D consecutive floats; box tables are 16-byte aligned when D is even) with vector loads
template <int D>
__device__ __forceinline__ void load_box(const float* __restrict__ b, float (&lo)[D], float (&hi)[D])
{
    if constexpr (D % 4 == 0) {
#pragma unroll
        for (int k = 0; k < D; k += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(b + k));
            const float4 c = __ldg(reinterpret_cast<const float4*>(b + D + k));
            lo[k] = a.x; lo[k + 1] = a.y; lo[k + 2] = a.z; lo[k + 3] = a.w;
            hi[k] = c.x; hi[k + 1] = c.y; hi[k + 2] = c.z; hi[k + 3] = c.w;
        }
    } else if constexpr (D % 2 == 0) {
#pragma unroll
        for (int k = 0; k < D; k += 2) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(b + k));
            const float2 c = __ldg(reinterpret_cast<const float2*>(b + D + k));
            lo[k] = a.x; lo[k + 1] = a.y; hi[k] = c.x; hi[k + 1] = c.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < D; k++) { lo[k] = __ldg(b + k); hi[k] = __ldg(b + D + k); }
    }
}

// ---- warp-cooperative helpers (lanes = boxes / points) ----
template <int D>
__device__ __forceinline__ float box_lb(const float* __restrict__ b, const float (&zq)[D])
{
    float lb = 0.f;
#pragma unroll
    for (int k = 0; k < D; k++) {
        const float lo = __ldg(b + k), hi = __ldg(b + D + k);
        const float gap = fmaxf(fmaxf(lo - zq[k], zq[k] - hi), 0.f);      // empty boxes are (inf, -inf): gap = inf
        lb = fmaf(gap, gap, lb);
    }
    return lb;
}

// min over a tile of the weighted exponent |c dz|^2 - log2 w, lanes = points (warp cooperative, warp-uniform result)
template <int D>
__device__ __forceinline__ float coop_scan_tile(const float* __restrict__ tile, const float (&zq)[D], bool has_w, int lane)
{
    constexpr int TN = tile_points(D);
    float b = INFINITY;
#pragma unroll
    for (int p = lane; p < TN; p += 32) {
        float a = has_w ? fminf(-__log2f(__ldg(tile + D * TN + p)), LW_CAP) : 0.f;
#pragma unroll
        for (int k = 0; k < D; k++) {
            const float d = zq[k] - __ldg(tile + k * TN + p);
            a = fmaf(d, d, a);
        }
        b = fminf(b, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
    return b;
}

// (value, index) argmin over the warp; ties keep the smaller index; result is warp uniform
__device__ __forceinline__ void warp_argmin(float& v, int& i)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

// Reference exponents, per query and independent of split / batch / GPU:
//  (1) nearest super-box (SG consecutive kd leaves) by box distance, then the nearest tile box inside
//      it: the query's "home" tile;
//  (2) the home tile is evaluated exactly (min-only, straight from L2).  A real point's distance is an
//      upper bound of the true minimum, and the true nearest neighbour usually lives in the home tile,
//      so the reference is tight from the first tile on — tight references are what lets the box test
//      cull.  (If a closer point turns up later the reference is tightened / recentred on the fly.)
// (1) of the reference search: the super-box nearest to each of the thread's queries, exact branch and bound over the chunk
// boxes (ties keep the first)
template <int D, int Q>
__device__ __forceinline__ void home_superbox(const KdeDev* __restrict__ kde, const float (&q2)[Q][D], int (&sbest)[Q], bool alive = true)
{
    const float* __restrict__ sbx = kde->sboxes;
    const int nsuper = kde->nsuper;
    {
        float lbbest[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) { lbbest[q] = INFINITY; sbest[q] = 0; }
        // branch and bound over the chunk boxes: a chunk whose box is no closer than the best super-box so
        // far cannot hold a closer one (ties keep the first), so the warp skips it when no lane wants it
        const float* __restrict__ cbx = kde->cboxes;
        const int nchunk = kde->nchunk;
#pragma unroll 1
        for (int cidx = 0; cidx < nchunk; cidx++) {
            bool want = false;
            {
                float lo[D], hi[D];
                load_box<D>(cbx + (size_t)cidx * 2 * D, lo, hi);
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    float lb = 0.f;
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        const float zq = q2[q][k];
                        const float gap = fmaxf(fmaxf(lo[k] - zq, zq - hi[k]), 0.f);
                        lb = fmaf(gap, gap, lb);
                    }
                    want |= (lb < lbbest[q]);
                }
                if (!(lo[0] <= hi[0]) || !alive) want = false;
            }
            if (!__any_sync(0xffffffffu, want)) continue;
            const int send = (cidx + 1) * CG < nsuper ? (cidx + 1) * CG : nsuper;
#pragma unroll 1
            for (int sidx = cidx * CG; sidx < send; sidx++) {
                float lo[D], hi[D];
                load_box<D>(sbx + (size_t)sidx * 2 * D, lo, hi);
                if (!(lo[0] <= hi[0])) continue;
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    float lb = 0.f;
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        const float zq = q2[q][k];
                        const float gap = fmaxf(fmaxf(lo[k] - zq, zq - hi[k]), 0.f);
                        lb = fmaf(gap, gap, lb);
                    }
                    if (lb < lbbest[q]) { lbbest[q] = lb; sbest[q] = sidx; }
                }
            }
        }
    }
}

template <int D, int Q>
__device__ __forceinline__ void home_reference(const KdeDev* __restrict__ kde, int ntiles, const float (&q2)[Q][D],
                                               float (&bmin)[Q], float (&mref)[Q], float (&nm2)[Q], bool alive = true)
{
    constexpr int TN = tile_points(D);
    constexpr int STAGE_FLOATS = stage_floats(D);
    {
        const float* __restrict__ sbx = kde->sboxes;
        const float* __restrict__ bx = kde->boxes;
        const int nsuper = kde->nsuper;
        const bool has_w = kde->has_w != 0;
        const float ref_bias = kde->ref_bias;
        int sbest[Q];
        home_superbox<D, Q>(kde, q2, sbest, alive);
#pragma unroll
        for (int q = 0; q < Q; q++) {
            if (!alive) { bmin[q] = ref_bias; mref[q] = 0.f; nm2[q] = 0.f; continue; }
            float zq[D];
#pragma unroll
            for (int k = 0; k < D; k++) zq[k] = q2[q][k];
            // nearest tile box inside the home super-box
            int tbest = sbest[q] * SG;
            float lbt = INFINITY;
            const int tend = (sbest[q] + 1) * SG < ntiles ? (sbest[q] + 1) * SG : ntiles;
#pragma unroll 1
            for (int t = sbest[q] * SG; t < tend; t++) {
                float lb = 0.f;
                bool valid = true;
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const float lo = __ldg(bx + (size_t)t * 2 * D + k), hi = __ldg(bx + (size_t)t * 2 * D + D + k);
                    valid &= (lo <= hi);
                    const float gap = fmaxf(fmaxf(lo - zq[k], zq[k] - hi), 0.f);
                    lb = fmaf(gap, gap, lb);
                }
                if (valid && lb < lbt) { lbt = lb; tbest = t; }
            }
            const float* __restrict__ ht = kde->tiles + (size_t)tbest * STAGE_FLOATS;
            float best = INFINITY;
            best = scan_tile_min<D>(ht, zq, has_w, best);
            // D >= 5: the other tiles of the home super-box whose box is closer than the best point so far
#pragma unroll 1
            for (int t = sbest[q] * SG; SKB_HOME_SUPER && D >= 5 && t < tend; t++) {
                float lb = 0.f;
                bool valid = (t != tbest);
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const float lo = __ldg(bx + (size_t)t * 2 * D + k), hi = __ldg(bx + (size_t)t * 2 * D + D + k);
                    valid &= (lo <= hi);
                    const float gap = fmaxf(fmaxf(lo - zq[k], zq[k] - hi), 0.f);
                    lb = fmaf(gap, gap, lb);
                }
                if (!(valid && lb < best)) continue;
                const float* __restrict__ ot = kde->tiles + (size_t)t * STAGE_FLOATS;
                best = scan_tile_min<D>(ot, zq, has_w, best);
            }
            if (!(best < 1.0e30f)) best = ref_bias;                   // inf/NaN rows (or padding only): finite reference
            bmin[q] = best;
            const float m = floorf(best) - ref_bias;
            mref[q] = m;
            nm2[q] = -m;
        }
    }
}

// ---- epilogue: finalize from registers, or publish the partial state and let the last-arriving CTA of a
// finalize block merge splits / form the ratio + accept ----
template <int D, int Q, int NTH = NT>
__device__ __forceinline__ void finish_queries(const ScoreLaunch& L, const ScoreJob& job, long long qbase, int tid, long long M,
                                               const float (&mref)[Q], const double (&S)[Q], int* s_last)
{
    constexpr int QB = NTH * Q;
    const FinJob& myfin = L.fin.fj[job.fin_slot];
    if (job.counters == nullptr) {
        // single split, ungrouped: finalize straight from registers
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const long long v = qbase + (long long)tid * Q + q;          // position in the visit order
            if (v < M) {
                const long long r = job.perm ? job.perm[v] : v;          // row of the query table
                if (myfin.out_m) { myfin.out_m[r] = myfin.log_wmax - (double)mref[q] * LN2; myfin.out_s[r] = S[q]; }
                if (myfin.out_logp) myfin.out_logp[r] = final_logp(myfin.log_norm, S[q], (double)mref[q]);
            }
        }
        return;
    }
    // partial states and arrival counters are indexed by visit-order position (shared by a group)
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const long long v = qbase + (long long)tid * Q + q;
        if (v < M) myfin.partial[(long long)blockIdx.z * M + v] = make_double2((double)mref[q], S[q]);
    }
    // last-arriving CTA of each finalize block (FB queries) merges splits / forms the ratio + accept
    __threadfence();
    __syncthreads();
    if (tid < QB / FB) {
        const long long fb = qbase / FB + tid;
        int last = 0;
        if (fb * FB < M) {
            const int prev = atomicAdd(&job.counters[fb], 1);
            last = (prev == job.expected - 1);
            if (last) job.counters[fb] = 0;             // self-reset for the next call
        }
        s_last[tid] = last;
    }
    __syncthreads();
    if constexpr (Q == 1 && NTH > FB && NTH % FB == 0) {
        // one query per thread, several finalize blocks per CTA: every thread finalizes its own row
        if (s_last[tid / FB]) {
            __threadfence();
            const long long r = qbase + tid;
            if (r < M) finalize_query(L.fin, job.fin_first, job.fin_count, M, r);
        }
        return;
    }
#pragma unroll 1
    for (int b = 0; b < QB / FB; b++) {
        if (!s_last[b]) continue;
        __threadfence();
        for (int i = tid; i < FB; i += NTH) {
            const long long r = qbase + (long long)b * FB + i;
            if (r < M) finalize_query(L.fin, job.fin_first, job.fin_count, M, r);
        }
    }
}


}  // namespace skb

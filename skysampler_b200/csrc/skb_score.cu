// Gaussian-KDE log-density scoring for sm_100a (B200): the replacement for sklearn's KD-tree
// `kernel_density` walk (sklearn neighbors/_binary_tree.pxi.tp:1393-1535, 2149-2290) that
// skysampler's KDEContainer.score_samples reaches (skysampler/emulator.py:380-385).
//
// Every (query, training point) pair is accounted for — no tolerance:
//     log p(z) = log_norm + log sum_j w_j exp(-|z - z_j|^2 / 2h^2)
// with fp32 pair terms and an fp64 log-sum-exp state per query (d <= 2: a tile's terms may be summed through its
// far-field record where a rigorous remainder bound keeps that as exact as the fp32 pair loop, see below).
//
// Arithmetic.  Per evaluated pair the FMA pipe sees exactly 2D+1 lane operations (D subtractions, D fused
// squares, one weighted accumulate) and the SFU one EX2:
//   * coordinates are prescaled by c = sqrt(log2(e)/2)/h so that exp(-d^2/2h^2) = exp2(-|c dz|^2);
//   * the per-query reference exponent m is folded into the accumulator seed (acc starts at -m), so no
//     per-pair shift is needed; m only has to lie in a ~216-binade window around the true minimum
//     distance.  It comes from an exact evaluation of the query's home tiles and is tightened from the
//     exponents of the block sums; a 64-point block whose fp32 partial sum overflows is re-run after a
//     min-only pass over that block (rare);
//   * the weight is applied as sum = fma(w_j, e, sum) — same cost as an add;
//   * two training points ride in one FADD2/FFMA2 (packed fp32x2, one issue slot for two lane ops), which
//     leaves issue slots for MUFU.EX2 and LDS — measured on B200: 8xFFMA2 + 2xMUFU per loop co-issue at
//     98% of the FMA peak, the scalar form is issue-bound at 80%.
// Culling.  Training tiles are kd-tree leaves with bounding boxes (three levels: tile, super-box of 16 tiles,
// chunk of 16 super-boxes) and a warp's queries are Morton neighbours: a tile whose box is > 128 binades
// beyond a query's reference has only terms that flush to zero in the arithmetic above and is skipped.
// Kernels of this file (the sorted-incidence pipeline for d = 3..5 lives in skb_sparse.cu):
//   skb_score_d1_kernel     d = 1: sorted intervals, one far-field record per tile (comment above the kernel).
//   skb_score_kernel<D,Q>   lanes = queries (Q whole queries per thread, tile words broadcast by LDS.128, no
//                           cross-lane reduction); a tile is evaluated when ANY query of the warp needs it; tiles
//                           arrive in shared memory through a private ring of 1-D TMA bulk copies (cp.async.bulk +
//                           mbarrier complete_tx).  d = 2 (with two levels of far-field records), KDEs of <= 512 tiles,
//                           and, with larger Q, every dimension when culling is off.
//   skb_score_pq_kernel<D>  d >= 6: lanes = two training points each, read straight from L2; box tests one tile per
//                           lane; only the queries that need a tile are evaluated (comment above the kernel).

#include "skb_score_common.cuh"

namespace skb {

// ---------------------------------------------------------------------------------------------
// Per-query culling ("PQ") kernel, instantiated for D >= SKB_PQ_MIN_D and routed for D >= 6 (skb_api.cu).  With lanes = queries a tile is evaluated for all 32
// queries of the warp as soon as ONE of them needs it; in 8-D a query needs ~1.2 % of the kd leaves but the
// union over a warp's 32 Morton neighbours is ~5 %, so 3/4 of the pairs evaluated that way are terms that
// flush to zero.  This kernel turns the warp around:
//   * the training side is walked in chunks of PQ_CH super-boxes (PQ_CH * SG tiles).  Per chunk, one query
//     per lane tests the super-boxes (ballot = which queries need it); then one TILE per lane tests its box
//     against exactly those queries (coordinates broadcast from shared memory) and the non-empty
//     (tile, query mask) pairs are compacted into a work list in shared memory;
//   * the work list is streamed: every lane reads TWO training points of the 64-point tile straight from L2
//     (coalesced LDG.64 rows; no shared-memory staging) and keeps them in registers (one packed fp32x2 operand per dimension) and the tile's query mask is walked bit by
//     bit: only the queries that need the tile are evaluated.  Lane l adds its two terms to part[query][l]
//     (fp32, shared memory); rows are folded into the owner lane's fp64 (m, S) state at fixed tile-index
//     boundaries (every 2^PQ_EPOCH_SHIFT tiles) in a fixed order.
// A query's arithmetic depends on its own coordinates and the fitted KDE only (masks, references, fold points
// are all functions of the query alone), so results stay bit-identical under any batching or sharding.
// ---------------------------------------------------------------------------------------------
#ifndef SKB_PQ
#define SKB_PQ 1
#endif
#ifndef SKB_PQ_EPOCH_SHIFT
#define SKB_PQ_EPOCH_SHIFT 8
#endif
#ifndef SKB_TEST_UNROLL
#define SKB_TEST_UNROLL 4     // tile box tests of the need-bitmap phase (lanes = queries kernel) in flight at once
#endif
#ifndef SKB_PQ_MIN_D
#define SKB_PQ_MIN_D 5
#endif
#ifndef SKB_PQ_MINB
#define SKB_PQ_MINB 24        // launch bound (CTAs per SM) of the PQ kernel for D <= 8 (80 registers; 22.6 vs 23.2 ms at 22, 23.2 at 26: spills)
#endif
#ifndef SKB_PQ_COOP_SCAN
#define SKB_PQ_COOP_SCAN 1     // 1: the home tiles are scanned by the whole warp, one query at a time (d=8 24.4 -> 23.4 ms); 0: one private walk per lane
#endif
#ifndef SKB_PQ_MINB_TOP
#define SKB_PQ_MINB_TOP 16    // D > 12 (128 registers)
#endif
#ifndef SKB_PQ_MINB_HI
#define SKB_PQ_MINB_HI 20     // same for 8 < D <= 12 (80 registers); D > 12 keeps 16 (128 registers)
#endif
constexpr int PQ_CH = CG;                         // one work list per chunk box
constexpr int PQ_CAP = PQ_CH * SG;                // work-list capacity: every tile of a chunk
constexpr int PQ_EPOCH_SHIFT = SKB_PQ_EPOCH_SHIFT;                 // fp32 row partials are folded into fp64 every 256 tile indices
constexpr int PQ_ROW = 33;                        // part[query][lane], rows padded by one word: conflict-free for both walks
#ifndef SKB_PQ_UNROLL
#define SKB_PQ_UNROLL 2       // queries of a tile's mask evaluated per step (2 or 4)
#endif
#ifndef SKB_PQ_UNROLL4_MAX_D
#define SKB_PQ_UNROLL4_MAX_D 6   // ... 4 at a time up to this dimension (short chains, registers to spare: +7 % at d=5)
#endif
#ifndef SKB_PQ_SLACK
#define SKB_PQ_SLACK 3        // a lane term >= 2^(SLACK - REF_BIAS) (a point that much closer than the reference) recentres
#endif
static_assert(PQ_CH <= 32 && PQ_CAP <= 256 && SG == 16, "PQ chunk layout");
__host__ __device__ constexpr bool pq_mode(int D, int Q) { return SKB_PQ && D >= SKB_PQ_MIN_D && Q == 1 && tile_points(D) == 64; }
__host__ __device__ constexpr int pq_qstride(int D) { return (D + 4) & ~3; }          // floats per query row: z[0..D), -m, padding to 16 B
__host__ __device__ constexpr size_t pq_smem_bytes(int D)
{
    return (size_t)32 * pq_qstride(D) * 4 + (size_t)32 * PQ_ROW * 4 + (size_t)PQ_CAP * 4 + (size_t)PQ_CAP + 16;
}

// index of the highest set bit of a non-zero mask, and the mask without it (FLO + BMSK + LOP3)
__device__ __forceinline__ int pop_top_bit(unsigned& mask)
{
    int idx;
    unsigned bit;
    asm("bfind.u32 %0, %1;" : "=r"(idx) : "r"(mask));
    asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(bit) : "r"(idx));
    mask ^= bit;
    return idx;
}

// shared memory through 32-bit shared-window addresses
__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v)
{
    asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory");
}

// one query row (z[0..D), -m, padding) from shared memory
template <int D>
__device__ __forceinline__ void pq_load_row(uint32_t qrow, float (&r)[pq_qstride(D)])
{
    constexpr int QS = pq_qstride(D);
#pragma unroll
    for (int k = 0; k < QS; k += 4) {
        if (k + 4 <= D + 1 || k < D + 1) {
            const float4 v = lds128(qrow + 4u * k);
            r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
        }
    }
}

// acc - m for this lane's two points against the query row in shared memory.  The row holds plain floats:
// ptxas folds pk(z, z) into the scalar-broadcast operand form of FADD2 / FFMA2 (`R.F32`), so duplicated
// pairs would only cost shared memory, LDS slots and registers.
template <int D>
__device__ __forceinline__ void pq_acc(uint32_t qrow, const u64 (&p2)[D], float& a0, float& a1)
{
    constexpr int QS = pq_qstride(D);
    float r[QS];
    pq_load_row<D>(qrow, r);
    u64 acc = pk(r[D], r[D]);
#pragma unroll
    for (int k = 0; k < D; k++) {
        const u64 dz = sub2(pk(r[k], r[k]), p2[k]);
        acc = fma2(dz, dz, acc);
    }
    upk(acc, a0, a1);
}

// fold row `row` of the fp32 partials into a double, fixed (lane) order, and clear it
__device__ __forceinline__ double pq_fold_row(float* __restrict__ part, int row)
{
    float* p = part + row * PQ_ROW;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int l = 0; l < 32; l += 4) {
        s0 += (double)p[l]; s1 += (double)p[l + 1]; s2 += (double)p[l + 2]; s3 += (double)p[l + 3];
        p[l] = 0.f; p[l + 1] = 0.f; p[l + 2] = 0.f; p[l + 3] = 0.f;
    }
    return __dadd_rn(__dadd_rn(s0, s1), __dadd_rn(s2, s3));
}

// One 2-D far-field record (skb_internal.h FF2_*: centre x, y, half widths rx, ry, then B_ij, i + j <= FF2_P, row i =
// B_i0 .. B_i,P-i) against the thread's Q queries: val = 2^(m - |s|^2) * sum B_ij sx^i sy^j, ok = the remainder bound holds
// (2 ln2 (|sx| rx + |sy| ry) <= 1.7).  An empty record (rx < 0) is the exact value 0.  With Q == 2 both queries ride in
// one packed fp32x2 Horner chain, the coefficient being the broadcast operand.
template <int Q>
__device__ __forceinline__ void ff2_eval(const float* __restrict__ rec, const float (&q2)[Q][2], const float (&nm2)[Q],
                                         float (&val)[Q], bool (&ok)[Q])
{
    const float4 cr = *reinterpret_cast<const float4*>(rec);
    const float* __restrict__ B = rec + 4;
    float sx[Q], sy[Q], p[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) { sx[q] = q2[q][0] - cr.x; sy[q] = q2[q][1] - cr.y; p[q] = 0.f; }
    if constexpr (Q == 2) {
        const u64 sx2 = pk(sx[0], sx[1]), sy2 = pk(sy[0], sy[1]);
        u64 p2 = pk(0.f, 0.f);
#pragma unroll
        for (int i = FF2_P; i >= 0; i--) {
            const int n = FF2_P + 1 - i;
            const int off = i * (FF2_P + 1) - (i * (i - 1)) / 2;
            u64 in2 = pk(B[off + n - 1], B[off + n - 1]);
#pragma unroll
            for (int j = n - 2; j >= 0; j--) in2 = fma2(in2, sy2, pk(B[off + j], B[off + j]));
            p2 = fma2(p2, sx2, in2);
        }
        upk(p2, p[0], p[1]);
    } else {
#pragma unroll
        for (int i = FF2_P; i >= 0; i--) {
            const int n = FF2_P + 1 - i;
            const int off = i * (FF2_P + 1) - (i * (i - 1)) / 2;
            float in[Q];
#pragma unroll
            for (int q = 0; q < Q; q++) in[q] = B[off + n - 1];
#pragma unroll
            for (int j = n - 2; j >= 0; j--) {
                const float b = B[off + j];
#pragma unroll
                for (int q = 0; q < Q; q++) in[q] = fmaf(in[q], sy[q], b);
            }
#pragma unroll
            for (int q = 0; q < Q; q++) p[q] = fmaf(p[q], sx[q], in[q]);
        }
    }
    const bool empty = cr.z < 0.f;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const float v = ex2_neg(fmaf(sy[q], sy[q], fmaf(sx[q], sx[q], nm2[q]))) * p[q];
        val[q] = empty ? 0.f : v;
        ok[q] = empty || (fabsf(sx[q]) * cr.z + fabsf(sy[q]) * cr.w <= FF2_SR_MAX);
    }
}

// ---------------------------------------------------------------------------------------------
// The score kernel.  grid = (query blocks, jobs, training splits), block = NT threads.
// ---------------------------------------------------------------------------------------------
template <int D, int Q>
__global__ void __launch_bounds__(NT, (Q == queries_per_thread(D) ? min_blocks_per_sm(D) : (SKB_DENSE_QMUL > 1 ? 2 : (D <= 4 ? 2 : (D <= 8 ? 3 : 4)))) * (128 / NT))
skb_score_kernel(const __grid_constant__ ScoreLaunch L)
{
    constexpr int QB = NT * Q;
    constexpr int TN = tile_points(D);
    constexpr int SUB = TN < SUB_MAX ? TN : SUB_MAX;
    constexpr int STAGE_FLOATS = stage_floats(D);
    constexpr uint32_t STAGE_BYTES = STAGE_FLOATS * sizeof(float);
    constexpr int NWARP = NT / 32;
    constexpr int NSTAGE = nstage(D);

    const ScoreJob& job = L.jobs[blockIdx.y];
    const long long M = job.M;
    const long long qbase = (long long)blockIdx.x * QB;
    if (qbase >= M || (int)blockIdx.z >= job.nsplit) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stages = reinterpret_cast<float*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NSTAGE * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + NSTAGE;
    // D == 2 (culled instantiation): staging buffer for the four cell records of a tile, between the barriers and the bitmap
    constexpr size_t CELL_BYTES = (D == 2 && Q == queries_per_thread(2)) ? (size_t)4 * FF2_CELL * sizeof(float) : 0;
    static_assert((2 * NSTAGE * sizeof(uint64_t)) % 16 == 0 || CELL_BYTES == 0, "cell buffer alignment");
    float* cellbuf = reinterpret_cast<float*>(empty_bar + NSTAGE);
    unsigned* need_bits = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(empty_bar + NSTAGE) + CELL_BYTES);   // one bit per tile: some warp needs it
    __shared__ int s_last[QB / FB];

    const int tid = threadIdx.x;
    const KdeDev* __restrict__ kde = job.kde;
    const int ntiles = kde->ntiles;
    const int nsplit = job.nsplit;
    const int t0 = (int)(((long long)ntiles * blockIdx.z) / nsplit);
    const int t1 = (int)(((long long)ntiles * (blockIdx.z + 1)) / nsplit);
    const int ntl = t1 - t0;
    const float* __restrict__ gtiles = kde->tiles + (size_t)t0 * STAGE_FLOATS;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NWARP); }
        mbar_fence_init();
    }
    const int nwords = (ntl + 31) >> 5;
    for (int w = tid; w < nwords; w += NT) need_bits[w] = 0u;
    __syncthreads();

    float q2[Q][D];
    load_queries<D, Q>(job, kde, qbase, tid, M, q2);

    // ---- main loop over this split's tiles -----------------------------------------------------
    // Reference exponent window.  Pair terms are t = 2^(mref - acc) in fp32 (flush below 2^-126,
    // overflow above 2^127).  The log-sum is carried by terms within ~30 binades of the closest
    // point, so any mref in (acc_min - 96, acc_min + 120) is exact to fp32 rounding.  mref is set to
    // (smallest acc seen) - REF_BIAS: points up to ~210 binades closer than anything seen so far are
    // absorbed without action; beyond that the SUB-point block overflows and is re-run once.
    float mref[Q];          // integer-valued reference exponent per query
    double S[Q];            // fp64 running sum at reference mref:  sum_j w_j 2^(-acc_j) = S * 2^(-mref)
    float nm2[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) { mref[q] = 0.f; S[q] = 0.0; nm2[q] = 0.f; }

    const int lane = tid & 31;
    unsigned n_rerun = 0, n_blocks = 0, n_culled = 0, n_visited = 0;

    float bmin[Q];          // running estimate of the smallest weighted exponent seen (absolute)
    const float ref_bias = kde->ref_bias;
    home_reference<D, Q>(kde, ntiles, q2, bmin, mref, nm2);
    // D == 2: second-level (cell) far-field records on?  A property of the fitted KDE (its leaves must be narrow enough for
    // the cells to serve what the tile record cannot), or forced either way for A/B runs
    const bool use_cells = (D == 2) && kde->ff2c != nullptr && (job.ff2_cells < 0 ? kde->ff2_cells_on != 0 : job.ff2_cells != 0);

    // ---- which tiles does this CTA need at all?  A tile that every warp can already cull with its initial
    // references is never requested from L2 (references only move down, so this is conservative).  Two
    // levels: a super-box that every query of the warp culls dismisses its SG tiles at once.
    if (job.cull) {
        const float* __restrict__ sbx = kde->sboxes;
        const float* __restrict__ bx = kde->boxes;
        const int s_first = t0 / SG, s_last_ = (t1 + SG - 1) / SG;
#pragma unroll 1
        for (int sidx = s_first; sidx < s_last_; sidx++) {
            bool need_s = false;
            {
                float lo[D], hi[D];
#pragma unroll
                for (int k = 0; k < D; k++) { lo[k] = __ldg(sbx + (size_t)sidx * 2 * D + k); hi[k] = __ldg(sbx + (size_t)sidx * 2 * D + D + k); }
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    float lb = 0.f;
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        const float zq = q2[q][k];
                        const float gap = fmaxf(fmaxf(lo[k] - zq, zq - hi[k]), 0.f);
                        lb = fmaf(gap, gap, lb);
                    }
                    need_s |= !(lb - mref[q] > CULL_MARGIN);
                }
                if (!(lo[0] <= hi[0])) need_s = false;
            }
            if (!__any_sync(0xffffffffu, need_s)) continue;
            unsigned word = 0u;                                        // SG = 16 tiles -> half a bitmap word
            const int tb = sidx * SG;
            constexpr int kTestUnroll = SKB_TEST_UNROLL;
#pragma unroll kTestUnroll
            for (int b = 0; b < SG; b++) {
                const int t = tb + b;
                if (t < t0 || t >= t1) continue;
                float lo[D], hi[D];
#pragma unroll
                for (int k = 0; k < D; k++) { lo[k] = __ldg(bx + (size_t)t * 2 * D + k); hi[k] = __ldg(bx + (size_t)t * 2 * D + D + k); }
                bool need = false;
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    float lb = 0.f;
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        const float zq = q2[q][k];
                        const float gap = fmaxf(fmaxf(lo[k] - zq, zq - hi[k]), 0.f);
                        lb = fmaf(gap, gap, lb);
                    }
                    need |= !(lb - mref[q] > CULL_MARGIN);
                }
                if (!(lo[0] <= hi[0])) need = false;                  // tile without weight
                if (__any_sync(0xffffffffu, need)) word |= 1u << b;
            }
            if (lane == 0 && word) {
                const int rel = tb - t0;                               // bit position of the super-box's first tile
                if (rel >= 0 && (rel & (SG - 1)) == 0) atomicOr(&need_bits[rel >> 5], word << (rel & 31));
                else {                                                 // split boundary inside a super-box
                    for (int b = 0; b < SG; b++)
                        if ((word >> b) & 1u) { const int r2 = tb + b - t0; atomicOr(&need_bits[r2 >> 5], 1u << (r2 & 31)); }
                }
            }
        }
    } else {
        for (int w = tid; w < nwords; w += NT) {
            const int rem = ntl - w * 32;
            need_bits[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
        }
    }
    __syncthreads();

    // cursors over the set bits: cc = tile this warp consumes next, pc (thread 0) = tile to request next
    int cw = -1, pw = -1;
    unsigned cbits = 0u, pbits = 0u;
    auto next_tile = [&](int& w, unsigned& bits) -> int {
        while (bits == 0u) {
            if (++w >= nwords) return -1;
            bits = need_bits[w];
        }
        const int b = __ffs(bits) - 1;
        bits &= bits - 1u;
        return w * 32 + b;
    };
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            const int t = next_tile(pw, pbits);
            if (t < 0) break;
            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
            tma_bulk_g2s(stages + (size_t)s * STAGE_FLOATS, gtiles + (size_t)t * STAGE_FLOATS, STAGE_BYTES, &full_bar[s]);
        }
    }

    for (int it = 0;; it++) {
        const int trel = next_tile(cw, cbits);           // tile index within this split
        if (trel < 0) break;
        n_visited++;
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        // producer: once every warp has released the stage of the previous tile, refill it
        if (tid == 0 && it >= 1) {
            const int t = next_tile(pw, pbits);
            if (t >= 0) {
                const int sp = (it - 1) % NSTAGE;
                mbar_wait(&empty_bar[sp], ((it - 1) / NSTAGE) & 1);
                mbar_expect_tx(&full_bar[sp], STAGE_BYTES);
                tma_bulk_g2s(stages + (size_t)sp * STAGE_FLOATS, gtiles + (size_t)t * STAGE_FLOATS, STAGE_BYTES,
                             &full_bar[sp]);
            }
        }
        __syncwarp();
        mbar_wait(&full_bar[s], ph);
        const float* tile = stages + (size_t)s * STAGE_FLOATS;

        u64 sum2[Q];
        float amin[Q];
        // cull: if for every query of this warp the nearest point of the tile's box is more than 127
        // binades beyond the reference, every term 2^(m - acc) of the tile flushes to zero in fp32 — the
        // tile contributes exactly nothing and is skipped (warp uniform; other queries' boxes never change
        // a query's own arithmetic).
        bool own_cull[Q];                                  // this query's own box test (the tile is visited if ANY query needs it)
#pragma unroll
        for (int q = 0; q < Q; q++) own_cull[q] = false;
        if (job.cull) {
            const float* hdr = tile + (D + 1) * TN;
            bool skip = true;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                float lb = 0.f;
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const float zq = q2[q][k];
                    const float gap = fmaxf(fmaxf(hdr[k] - zq, zq - hdr[D + k]), 0.f);
                    lb = fmaf(gap, gap, lb);
                }
                own_cull[q] = (lb - mref[q] > CULL_MARGIN);
                skip &= own_cull[q];
            }
            if (__all_sync(0xffffffffu, skip)) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                n_culled++;
                continue;
            }
        }
        // ---- D == 2: far-field record of the tile (skb_fit.cu: skb_moment2_kernel; it rides in the tile's header row and
        // arrived with the same TMA copy).  About the box centre c (s = q - c, t = x - c) a pair term is
        //     2^-|s - t|^2 = 2^-|s|^2 * e^(2 ln2 s.t) * 2^-|t|^2,
        // so the tile contributes 2^-|s|^2 * sum_{i+j<=12} B_ij sx^i sy^j; the Taylor series of e^u is cut after u^12 only
        // where |u| <= 2 ln2 (|sx| rx + |sy| ry) <= 1.7: relative remainder < 5e-6 of the tile's own positive sum.  One exponential
        // and 104 FMAs instead of 128 exponentials (the pair loop sits on the MUFU.EX2 roofline in 2-D).  Which path a
        // (query, tile) takes depends on the query and the fitted KDE alone; a query whose own box test culls the tile
        // takes the value 0 that the pair loop would give it; a value at the overflow limit goes to the pair loop, which
        // re-centres.  When every query of the warp is served the pair loop is skipped.
        bool ffq[Q];
        float ffv[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) { ffq[q] = false; ffv[q] = 0.f; }
        if constexpr (D == 2 && Q == queries_per_thread(2)) {
            if (job.cull) {
                float tv[Q];
                bool tok[Q];
                ff2_eval<Q>(tile + (D + 1) * TN + FF2_OFF, q2, nm2, tv, tok);
                bool all = true;
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    ffq[q] = own_cull[q] || (tok[q] && (tv[q] < OVERFLOW_LIMIT));
                    ffv[q] = own_cull[q] ? 0.f : tv[q];
                    all &= ffq[q];
                }
                // second level: the tile's four 32-point cells (median splits of the leaf, skb_fit.cu) carry records of their
                // own with half the radius; a query the tile record cannot serve is served when ALL FOUR cell records can
                bool try_cells = false;
                if (!__all_sync(0xffffffffu, all) && use_cells) {
                    // the cell records are fetched when some query still unserved passes all four of its remainder bounds (a
                    // query's own path never depends on its companions: it is served by the cells iff ITS bounds hold)
                    const float4* __restrict__ hdrs = reinterpret_cast<const float4*>(kde->ff2c + (size_t)(t0 + trel) * (4 * FF2_CELL));
                    bool pre[Q];
#pragma unroll
                    for (int q = 0; q < Q; q++) pre[q] = !ffq[q];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const float4 cr = __ldg(hdrs + c * (FF2_CELL / 4));
#pragma unroll
                        for (int q = 0; q < Q; q++)
                            pre[q] &= (cr.z < 0.f) || (fabsf(q2[q][0] - cr.x) * cr.z + fabsf(q2[q][1] - cr.y) * cr.w <= FF2_SR_MAX);
                    }
                    bool want = false;
#pragma unroll
                    for (int q = 0; q < Q; q++) want |= pre[q];
                    try_cells = __any_sync(0xffffffffu, want);
                }
                if (try_cells) {
                    const float4* __restrict__ src = reinterpret_cast<const float4*>(kde->ff2c + (size_t)(t0 + trel) * (4 * FF2_CELL));
                    float4* dst = reinterpret_cast<float4*>(cellbuf);
#pragma unroll
                    for (int i = 0; i < 4 * FF2_CELL / 4 / 32; i++) dst[lane + 32 * i] = __ldg(src + lane + 32 * i);
                    __syncwarp();
                    float cv[Q];
                    bool cok[Q];
#pragma unroll
                    for (int q = 0; q < Q; q++) { cv[q] = 0.f; cok[q] = true; }
#pragma unroll 1
                    for (int c = 0; c < 4; c++) {
                        float v[Q];
                        bool ok[Q];
                        ff2_eval<Q>(cellbuf + c * FF2_CELL, q2, nm2, v, ok);
#pragma unroll
                        for (int q = 0; q < Q; q++) { cv[q] += v[q]; cok[q] = cok[q] && ok[q]; }
                    }
                    __syncwarp();
                    all = true;
#pragma unroll
                    for (int q = 0; q < Q; q++) {
                        if (!ffq[q] && cok[q] && (cv[q] < OVERFLOW_LIMIT)) { ffq[q] = true; ffv[q] = cv[q]; }
                        all &= ffq[q];
                    }
                }
                if (__all_sync(0xffffffffu, all)) {
#pragma unroll
                    for (int q = 0; q < Q; q++) {
                        S[q] += (double)ffv[q];
                        if (ffv[q] > 0.f)
                            bmin[q] = fminf(bmin[q], mref[q] - (float)((__float_as_int(ffv[q]) >> 23) - 128));
                        ffq[q] = false;                                               // consumed
                    }
                    n_blocks++;
                    goto tile_done;
                }
            }
        }
#pragma unroll 1
        for (int sub = 0; sub < TN; sub += SUB) {
            const float* blk = tile + sub;
            n_blocks++;
            for (;;) {
#pragma unroll
                for (int q = 0; q < Q; q++) sum2[q] = pk(0.f, 0.f);
                tile_pass<D, Q, false>(blk, SUB, q2, nm2, sum2, amin);
                bool bad = false;
                float ts[Q];
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    float lo, hi; upk(sum2[q], lo, hi);
                    ts[q] = ffq[q] ? 0.f : lo + hi;                  // a query served by the far-field record ignores the pair loop
                    bad |= !(ts[q] < OVERFLOW_LIMIT);
                }
                if (!__any_sync(0xffffffffu, bad)) {
#pragma unroll
                    for (int q = 0; q < Q; q++) {
                        S[q] += (double)ts[q];
                        // ts >= its largest term 2^(m - acc_min): the exponent of the block sum bounds the
                        // closest point of the block to within the block size
                        if (ts[q] > 0.f)
                            bmin[q] = fminf(bmin[q], mref[q] - (float)((__float_as_int(ts[q]) >> 23) - 127));
                    }
                    break;
                }
                // rare: a query met a point far closer than its window allows -> recentre, re-run block
                n_rerun++;
#pragma unroll
                for (int q = 0; q < Q; q++) amin[q] = INFINITY;
                tile_pass<D, Q, true>(blk, SUB, q2, nm2, sum2, amin);
                bool fixable = false;
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    if (!(ts[q] < OVERFLOW_LIMIT)) {
                        const float delta = floorf(amin[q]) - ref_bias;   // new reference relative to mref
                        if (delta < 0.f && delta > -1.0e30f) {
                            fixable = true;
                            bmin[q] = fminf(bmin[q], mref[q] + floorf(amin[q]));
                            mref[q] += delta;
                            nm2[q] = -mref[q];
                            double e = (double)delta;
                            if (e < -4000.0) e = -4000.0;
                            S[q] = ldexp(S[q], (int)e);
                        } else {
                            ts[q] = NAN;                             // NaN/inf input row: propagate
                        }
                    }
                }
                if (!__any_sync(0xffffffffu, fixable)) {
#pragma unroll
                    for (int q = 0; q < Q; q++) S[q] += (double)ts[q];
                    break;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; q++)
            if (ffq[q]) {
                S[q] += (double)ffv[q];
                if (ffv[q] > 0.f) bmin[q] = fminf(bmin[q], mref[q] - (float)((__float_as_int(ffv[q]) >> 23) - 128));
            }
    tile_done:
        // move the reference down when the sums have shown a closer point than it was built for (exact:
        // S is rescaled by a power of two); a tighter reference culls more of the tiles still to come
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const float want = floorf(bmin[q]) - ref_bias;
            if (mref[q] - want >= 8.f) {
                S[q] = ldexp(S[q], (int)fmaxf(want - mref[q], -4000.f));
                mref[q] = want;
                nm2[q] = -want;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    if (job.stats && lane == 0) {
        if (n_rerun) atomicAdd(&job.stats[0], (unsigned long long)n_rerun * Q * SUB);
        atomicAdd(&job.stats[1], (unsigned long long)n_blocks * Q * SUB);
        atomicAdd(&job.stats[2], (unsigned long long)(n_culled + (unsigned)ntl - n_visited));
        atomicAdd(&job.stats[3], (unsigned long long)ntl);      // tiles of the split (culled = box-culled at either level)
    }

    finish_queries<D, Q>(L, job, qbase, tid, M, mref, S, s_last);
}

template <int D, bool STATS>
__global__ void __launch_bounds__(32, D <= 8 ? SKB_PQ_MINB : (D <= 12 ? SKB_PQ_MINB_HI : SKB_PQ_MINB_TOP))
skb_score_pq_kernel(const __grid_constant__ ScoreLaunch L)
{
    constexpr int Q = 1;
    constexpr int TN = tile_points(D);
    constexpr int QS = pq_qstride(D);
    constexpr int STAGE_FLOATS = stage_floats(D);
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(TN == 64 && NT == 32, "PQ: one warp per CTA, two tile points per lane");

    const ScoreJob& job = L.jobs[blockIdx.y];
    const long long M = job.M;
    const long long qbase = (long long)blockIdx.x * 32;
    if (qbase >= M || (int)blockIdx.z >= job.nsplit) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* rows = reinterpret_cast<float*>(smem_raw);                 // 32 query rows: z[0..D), -m
    float* part = rows + 32 * QS;                                     // part[query][lane], rows of PQ_ROW words
    unsigned* lmask = reinterpret_cast<unsigned*>(part + 32 * PQ_ROW);   // work list: query mask ...
    unsigned char* ltile = reinterpret_cast<unsigned char*>(lmask + PQ_CAP);   // ... and tile offset in the chunk
    __shared__ int s_last[32 / FB];

    const int lane = threadIdx.x;
    const KdeDev* __restrict__ kde = job.kde;
    const int ntiles = kde->ntiles;
    const int nsplit = job.nsplit;
    const int t0 = (int)(((long long)ntiles * blockIdx.z) / nsplit);
    const int t1 = (int)(((long long)ntiles * (blockIdx.z + 1)) / nsplit);

    float q2[Q][D];
    load_queries<D, Q>(job, kde, qbase, lane, M, q2);
    float mref[Q], bmin[Q];
    double S[Q];
    float nm2[Q];
    mref[0] = 0.f; S[0] = 0.0; nm2[0] = 0.f;
    const float ref_bias = kde->ref_bias;
    const float PQ_LIMIT = exp2f((float)SKB_PQ_SLACK - ref_bias);     // a lane term this large re-centres
#if SKB_PQ_COOP_SCAN
    // exponent reference: the home super-box by branch and bound with one query per lane; its tiles are then scanned by the
    // WHOLE warp, one query at a time (lanes = tiles for the box tests, lanes = points for the scans: coalesced 256-byte rows
    // instead of 32 private tile walks), nearest box first, while a tile's box can still beat the best point found.  The
    // result is the exact minimum over the home super-box, as before.
    {
        int sbest[1];
        home_superbox<D, 1>(kde, q2, sbest);
        const bool has_w = kde->has_w != 0;
        float mine = INFINITY;
#pragma unroll 1
        for (int src = 0; src < 32; src++) {
            float zq[D];
#pragma unroll
            for (int k = 0; k < D; k++) zq[k] = __shfl_sync(FULL, q2[0][k], src);
            const int si = __shfl_sync(FULL, sbest[0], src);
            const int t = si * SG + (lane & 15);
            const float lbt = (lane < SG && t < ntiles) ? box_lb<D>(kde->boxes + (size_t)t * 2 * D, zq) : INFINITY;
            float best = INFINITY;
            unsigned todo = __ballot_sync(FULL, lbt < INFINITY);
            while (todo) {
                float v = ((todo >> lane) & 1u) ? lbt : INFINITY;
                int l = lane;
                warp_argmin(v, l);
                if (!(v < best)) break;
                todo &= ~(1u << l);
                best = fminf(best, coop_scan_tile<D>(kde->tiles + (size_t)(si * SG + (l & 15)) * STAGE_FLOATS, zq, has_w, lane));
            }
            if (lane == src) mine = best;
        }
        if (!(mine < 1.0e30f)) mine = ref_bias;                   // inf/NaN rows (or padding only): finite reference
        bmin[0] = mine;
        mref[0] = floorf(mine) - ref_bias;
        nm2[0] = -mref[0];
    }
#else
    home_reference<D, Q>(kde, ntiles, q2, bmin, mref, nm2);
#endif

#pragma unroll
    for (int k = 0; k < D; k++) rows[lane * QS + k] = q2[0][k];
    rows[lane * QS + D] = -mref[0];
    for (int l = 0; l < 32; l++) part[l * PQ_ROW + lane] = 0.f;
    __syncwarp();
    // the hot loops address the rows and the partials through 32-bit shared-window addresses (ld/st.shared): with generic
    // pointers ptxas re-derived the window base (S2R SR_CgaCtaId + 4 ALU ops) inside every loop to save a register
    const uint32_t rows_s = smem_u32(rows);
    const uint32_t mypart_s = smem_u32(part) + 4u * (uint32_t)lane;    // + query * PQ_ROW * 4

    const float* __restrict__ cbx = kde->cboxes;
    const float* __restrict__ sbx = kde->sboxes;
    const float* __restrict__ bx = kde->boxes;
    const float* __restrict__ gtiles = kde->tiles;
    const bool cull = job.cull != 0;
    unsigned n_rerun = 0, n_visited = 0;
    unsigned long long n_pairs = 0;
    int epoch_cur = -1;

    const int s_first = (t0 / SG / PQ_CH) * PQ_CH, s_end = (t1 + SG - 1) / SG;   // chunk aligned; tiles outside [t0, t1) are masked
#pragma unroll 1
    for (int s_base = s_first; s_base < s_end; s_base += PQ_CH) {
        // ---- (1) chunk-box test, then super-box tests, one query per lane: lane j keeps the ballot of
        //          super-box s_base + j ----
        unsigned smask = 0u;
        {
            float zq[D];
#pragma unroll
            for (int k = 0; k < D; k++) zq[k] = q2[0][k];
            if (cull) {
                float lo[D], hi[D];
                load_box<D>(cbx + (size_t)(s_base / PQ_CH) * 2 * D, lo, hi);
                float lb = 0.f;
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const float gap = fmaxf(fmaxf(lo[k] - zq[k], zq[k] - hi[k]), 0.f);
                    lb = fmaf(gap, gap, lb);
                }
                if (!__any_sync(FULL, !(lb - mref[0] > CULL_MARGIN))) continue;
            }
            const int nsb = s_end - s_base < PQ_CH ? s_end - s_base : PQ_CH;
#pragma unroll 2
            for (int j = 0; j < nsb; j++) {
                float lo[D], hi[D];
                load_box<D>(sbx + (size_t)(s_base + j) * 2 * D, lo, hi);
                float lb = 0.f;
                const bool valid = (lo[0] <= hi[0]);
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const float gap = fmaxf(fmaxf(lo[k] - zq[k], zq[k] - hi[k]), 0.f);
                    lb = fmaf(gap, gap, lb);
                }
                const bool need = valid && (!cull || !(lb - mref[0] > CULL_MARGIN));
                const unsigned bal = __ballot_sync(FULL, need);
                if (lane == j) smask = bal;
            }
        }
        // ---- (2) tile tests, one tile per lane (two super-boxes per pass), against the queries of the
        //          super-box's ballot only; compact the (tile, query mask) pairs into the work list ----
        int cnt = 0;
        unsigned act = __ballot_sync(FULL, smask != 0u);
        while (act) {
            const int ja = __ffs(act) - 1;
            act &= act - 1u;
            int jb = -1;
            if (act) { jb = __ffs(act) - 1; act &= act - 1u; }
            const int j = (lane < 16) ? ja : jb;
            unsigned qm = __shfl_sync(FULL, smask, j < 0 ? 0 : j);
            const int toff = (j < 0 ? 0 : j) * SG + (lane & 15);
            const int t = s_base * SG + toff;
            if (j < 0 || t < t0 || t >= t1) qm = 0u;
            float lo[D], hi[D];
            {
                const int tc = t < ntiles ? t : ntiles - 1;
                const float* b = bx + (size_t)tc * 2 * D;
                if (D % 4 == 0) {
#pragma unroll
                    for (int k = 0; k < D; k += 4) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(b + k));
                        const float4 c = __ldg(reinterpret_cast<const float4*>(b + D + k));
                        lo[k] = a.x; lo[k + 1] = a.y; lo[k + 2] = a.z; lo[k + 3] = a.w;
                        hi[k] = c.x; hi[k + 1] = c.y; hi[k + 2] = c.z; hi[k + 3] = c.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < D; k++) { lo[k] = __ldg(b + k); hi[k] = __ldg(b + D + k); }
                }
            }
            if (!(lo[0] <= hi[0])) qm = 0u;                               // tile without weight
            unsigned tmask = 0u;
            while (__any_sync(FULL, qm != 0u)) {
                const bool on = qm != 0u;
                int i = pop_top_bit(qm);                                  // an empty mask stays empty (bfind gives -1, bmsk clamps to 0)
                i = on ? i : 0;
                float r[QS];
                pq_load_row<D>(rows_s + (uint32_t)i * (QS * 4), r);
                float lb = 0.f;
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const float gap = fmaxf(fmaxf(lo[k] - r[k], r[k] - hi[k]), 0.f);
                    lb = fmaf(gap, gap, lb);
                }
                if (on && (!cull || !(lb + r[D] > CULL_MARGIN))) tmask |= 1u << i;
            }
            const unsigned nz = __ballot_sync(FULL, tmask != 0u);
            if (tmask) {
                const int pos = cnt + __popc(nz & ((1u << lane) - 1u));
                lmask[pos] = tmask;
                ltile[pos] = (unsigned char)toff;
            }
            cnt += __popc(nz);
        }
        __syncwarp();
        if (cnt == 0) continue;
        if constexpr (STATS) n_visited += (unsigned)cnt;

        // ---- (3) stream the work list: every lane reads its two points of a tile straight from L2 (one coalesced
        //          256-byte row per dimension, LDG.64) and keeps them in registers while the tile's query mask is walked:
        //          no shared-memory staging, no barriers (the TMA ring this replaces cost ~60 issue slots per tile) ----
        const int tchunk = s_base * SG;
#pragma unroll 1
        for (int e = 0; e < cnt; e++) {
            unsigned mask = lmask[e];
            const int tcur = tchunk + ltile[e];
            const int epoch = tcur >> PQ_EPOCH_SHIFT;
            if (epoch != epoch_cur) {
                if (epoch_cur >= 0) { __syncwarp(); S[0] += pq_fold_row(part, lane); __syncwarp(); }
                epoch_cur = epoch;
            }
            u64 p2[D];
            float w0, w1;
            {
                const float* __restrict__ tile = gtiles + (size_t)tcur * STAGE_FLOATS + 2 * lane;
#pragma unroll
                for (int k = 0; k < D; k++) p2[k] = __ldg(reinterpret_cast<const u64*>(tile + k * TN));
                upk(__ldg(reinterpret_cast<const u64*>(tile + D * TN)), w0, w1);
            }
            if constexpr (STATS) n_pairs += 2u * __popc(mask);

            // rare: a point >= SLACK binades closer than the reference was built for (or a NaN row) — fold the
            // row, move the query's reference to the tile's closest point, re-evaluate
            auto recentre = [&](int i, float a0, float a1) -> float {
                float am = fminf(a0 + fminf(-__log2f(w0), LW_CAP), a1 + fminf(-__log2f(w1), LW_CAP));   // weighted exponents
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) am = fminf(am, __shfl_xor_sync(FULL, am, off));
                const float delta = floorf(am) - ref_bias;
                if constexpr (STATS) n_rerun++;
                if (delta < 0.f && delta > -1.0e30f) {
                    __syncwarp();
                    if (lane == i) {
                        S[0] += pq_fold_row(part, i);
                        S[0] = ldexp(S[0], (int)fmaxf(delta, -4000.f));
                        mref[0] += delta;
                        nm2[0] = -mref[0];
                        sts32(rows_s + (uint32_t)(i * QS + D) * 4u, -mref[0]);
                    }
                    __syncwarp();
                    pq_acc<D>(rows_s + (uint32_t)i * (QS * 4), p2, a0, a1);
                    return __fmaf_rn(w1, ex2_neg(a1), __fmul_rn(w0, ex2_neg(a0)));
                }
                if (lane == i) S[0] = NAN;                                  // NaN / inf query row: propagate
                return 0.f;
            };

            // NQ queries of the mask per step: independent chains for the FMA / SFU latencies
            auto eval = [&](auto nqc) {
                constexpr int NQ = decltype(nqc)::value;
                int idx[NQ];
                float pa[NQ], a0[NQ], a1[NQ], t[NQ];
#pragma unroll
                for (int n = 0; n < NQ; n++) idx[n] = pop_top_bit(mask);
#pragma unroll
                for (int n = 0; n < NQ; n++) pa[n] = lds32(mypart_s + (uint32_t)idx[n] * (PQ_ROW * 4));   // issued early: off the critical path
#pragma unroll
                for (int n = 0; n < NQ; n++) pq_acc<D>(rows_s + (uint32_t)idx[n] * (QS * 4), p2, a0[n], a1[n]);
                bool bad = false;
#pragma unroll
                for (int n = 0; n < NQ; n++) {
                    t[n] = __fmaf_rn(w1, ex2_neg(a1[n]), __fmul_rn(w0, ex2_neg(a0[n])));
                    bad |= !(t[n] < PQ_LIMIT);
                }
                if (__any_sync(FULL, bad)) {
#pragma unroll
                    for (int n = 0; n < NQ; n++)
                        if (__any_sync(FULL, !(t[n] < PQ_LIMIT))) {
                            t[n] = recentre(idx[n], a0[n], a1[n]);
                            pa[n] = lds32(mypart_s + (uint32_t)idx[n] * (PQ_ROW * 4));
                        }
                }
#pragma unroll
                for (int n = 0; n < NQ; n++) sts32(mypart_s + (uint32_t)idx[n] * (PQ_ROW * 4), pa[n] + t[n]);
            };
            if constexpr (SKB_PQ_UNROLL >= 4 || D <= SKB_PQ_UNROLL4_MAX_D)
                while (__popc(mask) >= 4) eval(std::integral_constant<int, 4>{});
            while (mask & (mask - 1u)) eval(std::integral_constant<int, 2>{});       // at least two bits left
            if (mask) eval(std::integral_constant<int, 1>{});
        }
        __syncwarp();                                                      // the list is rebuilt next chunk
    }
    __syncwarp();
    S[0] += pq_fold_row(part, lane);

    if (STATS && job.stats && lane == 0) {
        if (n_rerun) atomicAdd(&job.stats[0], (unsigned long long)n_rerun * 2);
        atomicAdd(&job.stats[1], n_pairs);
        atomicAdd(&job.stats[2], (unsigned long long)((unsigned)(t1 - t0) - n_visited));
        atomicAdd(&job.stats[3], (unsigned long long)(t1 - t0));
    }
    finish_queries<D, Q>(L, job, qbase, lane, M, mref, S, s_last);
}

// ---------------------------------------------------------------------------------------------
// D == 1.  One coordinate means the tiles are SORTED intervals, and a 128-point tile is far narrower than the
// kernel: a tile's 128 terms collapse into one exponential times a degree-5 polynomial in s = q - centre
// (skb_fit.cu: skb_moment_kernel; cut when 2 ln2 |s| r <= 1/4, relative remainder < 6e-7 of the tile's own positive
// contribution).  A query needs 14-37 % of ALL points in 1-D, so the pair loop sat on the MUFU.EX2 roofline (95 ms for
// 1e6 x 2^20); per tile this is ~12 instructions instead of 128 exponentials.
//   lanes = queries (Morton order = sorted, so a warp's queries walk the same tiles and every table load is a broadcast);
//   reference exponent: the nearest point is in the tile that holds q or the one before it (binary search over the tile
//     intervals, then an exact weighted scan of the home tile and its two neighbours) — the true weighted minimum is
//     within 64 binades of it (LW_CAP), the fp32 window has 216: no term can overflow and there is no re-centring;
//   window: the same box test as everywhere (a tile whose interval is > 128 binades past the reference holds only terms
//     that flush), evaluated per query on a range found by two binary searches; tiles too wide for the series (sparse
//     tails) are evaluated point by point for exactly the queries that need it;
//   fp32 tile values are folded into the query's fp64 sum at fixed tile indices.
// Everything is a function of the query and the fitted KDE alone: results do not depend on the batch.
// ---------------------------------------------------------------------------------------------
constexpr int D1_NT = 128;
constexpr int D1_FOLD = 64;
constexpr float D1_SR_MAX = 0.25f / 1.3862943611198906f;      // |s| r <= 1/4 / (2 ln 2)

__global__ void __launch_bounds__(D1_NT)
skb_score_d1_kernel(const __grid_constant__ ScoreLaunch L)
{
    constexpr int TN = tile_points(1);
    constexpr int STAGE_FLOATS = stage_floats(1);
    constexpr unsigned FULL = 0xffffffffu;
    const ScoreJob& job = L.jobs[blockIdx.y];
    const long long M = job.M;
    const long long qbase = (long long)blockIdx.x * D1_NT;
    if (qbase >= M || (int)blockIdx.z >= job.nsplit) return;
    __shared__ int s_last[D1_NT / FB];

    const int tid = threadIdx.x;
    const KdeDev* __restrict__ kde = job.kde;
    const int ntiles = kde->ntiles;
    const int nsplit = job.nsplit;
    const int t0 = (int)(((long long)ntiles * blockIdx.z) / nsplit);
    const int t1 = (int)(((long long)ntiles * (blockIdx.z + 1)) / nsplit);
    const float* __restrict__ bx = kde->boxes;                      // [tile][lo, hi]
    const float* __restrict__ gtiles = kde->tiles;
    const float4* __restrict__ mom = reinterpret_cast<const float4*>(kde->mom);
    const bool has_w = kde->has_w != 0;
    const float ref_bias = kde->ref_bias;

    float q2[1][1];
    load_queries<1, 1>(job, kde, qbase, tid, M, q2);
    const float zq = q2[0][0];

    // ---- reference exponent from the nearest point: home tile = first interval that ends at or after q ----
    int th;
    {
        int lo = 0, hi = ntiles - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(bx + 2 * mid + 1) < zq) lo = mid + 1; else hi = mid;
        }
        th = lo;
    }
    float best = INFINITY;
    {
        const float zz[1] = {zq};
        const int ta = th > 0 ? th - 1 : 0, tb = th + 1 < ntiles ? th + 1 : ntiles - 1;
        for (int t = ta; t <= tb; t++) best = scan_tile_min<1>(gtiles + (size_t)t * STAGE_FLOATS, zz, has_w, best);
    }
    if (!(best < 1.0e30f)) best = ref_bias;                        // inf / NaN rows: finite reference
    const float m = floorf(best) - ref_bias;
    const float nm = -m;

    // ---- the tiles this query can see: intervals within sqrt(m + 128) of q (conservative range, exact test below) ----
    int ta = 1, tb = 0;                                            // empty
    if (zq == zq) {
        const float R = sqrtf(fmaxf(m + CULL_MARGIN, 0.f)) * 1.0001f + 1.0e-3f;
        const float qa = zq - R, qb = zq + R;
        int lo = 0, hi = ntiles;                                   // first tile with hi_t >= qa
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(bx + 2 * mid + 1) < qa) lo = mid + 1; else hi = mid;
        }
        ta = lo;
        lo = 0; hi = ntiles;                                       // first tile with lo_t > qb
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(bx + 2 * mid) <= qb) lo = mid + 1; else hi = mid;
        }
        tb = lo - 1;
        ta = ta < t0 ? t0 : ta;
        tb = tb > t1 - 1 ? t1 - 1 : tb;
    }
    const int wa = __reduce_min_sync(FULL, ta <= tb ? ta : 0x7fffffff);
    const int wb = __reduce_max_sync(FULL, ta <= tb ? tb : -1);

    double S = 0.0;
    float sum = 0.f;
#pragma unroll 1
    for (int t = wa; t <= wb; t++) {
        const float4 r0 = __ldg(mom + 3 * (size_t)t);              // lo, hi, centre, radius
        const float gap = fmaxf(fmaxf(r0.x - zq, zq - r0.y), 0.f);
        const bool in = (t >= ta) && (t <= tb) && (r0.x <= r0.y) && !(gap * gap - m > CULL_MARGIN);
        const float s = zq - r0.z;
        const bool wide = in && !(fabsf(s) * r0.w <= D1_SR_MAX);
        if (__any_sync(FULL, in && !wide)) {
            const float4 r1 = __ldg(mom + 3 * (size_t)t + 1);      // B0..B3
            const float4 r2 = __ldg(mom + 3 * (size_t)t + 2);      // B4, B5
            float p = fmaf(r2.y, s, r2.x);
            p = fmaf(p, s, r1.w);
            p = fmaf(p, s, r1.z);
            p = fmaf(p, s, r1.y);
            p = fmaf(p, s, r1.x);
            const float e = ex2_neg(fmaf(s, s, nm));
            if (in && !wide) sum = fmaf(e, p, sum);
        }
        if (__any_sync(FULL, wide)) {                               // sparse tails: this tile point by point
            const float* __restrict__ tile = gtiles + (size_t)t * STAGE_FLOATS;
            float sd = 0.f;
#pragma unroll 2
            for (int j = 0; j < TN; j += 4) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(tile + j));
                const float4 w = __ldg(reinterpret_cast<const float4*>(tile + TN + j));
                const float d0 = zq - x.x, d1 = zq - x.y, d2 = zq - x.z, d3 = zq - x.w;
                sd = fmaf(w.x, ex2_neg(fmaf(d0, d0, nm)), sd);
                sd = fmaf(w.y, ex2_neg(fmaf(d1, d1, nm)), sd);
                sd = fmaf(w.z, ex2_neg(fmaf(d2, d2, nm)), sd);
                sd = fmaf(w.w, ex2_neg(fmaf(d3, d3, nm)), sd);
            }
            if (wide) sum += sd;
        }
        if ((t & (D1_FOLD - 1)) == D1_FOLD - 1) { S += (double)sum; sum = 0.f; }
    }
    S += (double)sum;
    if (!(zq == zq)) S = NAN;

    const float mref[1] = {m};
    const double Sv[1] = {S};
    finish_queries<1, 1, D1_NT>(L, job, qbase, tid, M, mref, Sv, s_last);
}

template <int D, int Q>
static cudaError_t launch_score_dq(const ScoreLaunch& L, int njobs, long long max_rows, int max_split,
                                   int max_ntiles, cudaStream_t st)
{
    constexpr int QB = pq_mode(D, Q) ? 32 : NT * Q;
    dim3 grid((unsigned)((max_rows + QB - 1) / QB), (unsigned)njobs, (unsigned)max_split);
    if constexpr (D == 1 && Q == queries_per_thread(1)) {
        // the culled 1-D path: sorted intervals + far-field tile records (the all-pairs instantiation, Q = 8, keeps the pair loop)
        bool all_mom = true;
        for (int j = 0; j < njobs; j++) all_mom = all_mom && L.jobs[j].cull;
        if (all_mom) {
            dim3 g1((unsigned)((max_rows + D1_NT - 1) / D1_NT), (unsigned)njobs, (unsigned)max_split);
            skb_score_d1_kernel<<<g1, D1_NT, 0, st>>>(L);
            count_launch();
            return cudaGetLastError();
        }
    }
    if constexpr (pq_mode(D, Q)) {
        const size_t smem = pq_smem_bytes(D);
        bool stats = false;
        for (int j = 0; j < njobs; j++) stats |= (L.jobs[j].stats != nullptr);
        auto kern = stats ? skb_score_pq_kernel<D, true> : skb_score_pq_kernel<D, false>;   // counters are compiled out of the release kernel
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, 32, smem, st>>>(L);
    } else {
        constexpr int NSTAGE = nstage(D);
        const size_t smem = (size_t)NSTAGE * stage_floats(D) * sizeof(float) + 2 * NSTAGE * sizeof(uint64_t) +
                            (((size_t)max_ntiles + 31) / 32) * 4 + 16 + (D == 2 ? (size_t)4 * FF2_CELL * 4 : 0);
        cudaError_t e = cudaFuncSetAttribute(skb_score_kernel<D, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        skb_score_kernel<D, Q><<<grid, NT, smem, st>>>(L);
    }
    count_launch();
    return cudaGetLastError();
}

template <int D, int Q>
static int occupancy_dq()
{
    int n = 0;
    cudaError_t e;
    if constexpr (pq_mode(D, Q)) {
        const size_t smem = pq_smem_bytes(D);
        e = cudaFuncSetAttribute(skb_score_pq_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, skb_score_pq_kernel<D, false>, 32, smem);
    } else {
        constexpr int NSTAGE = nstage(D);
        const size_t smem = (size_t)NSTAGE * stage_floats(D) * sizeof(float) + 2 * NSTAGE * sizeof(uint64_t) + 1024;
        e = cudaFuncSetAttribute(skb_score_kernel<D, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, skb_score_kernel<D, Q>, NT, smem);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 2;
    }
    return n > 0 ? n : 1;
}

int score_ctas_per_sm(int D, bool dense_q)
{
    static int cache[2][SKB_MAXD + 1] = {{0}, {0}};
    if (D < 1 || D > SKB_MAXD) return 1;
    int& c = cache[dense_q ? 1 : 0][D];
    if (c == 0) {
        switch (D) {
#define SKB_CASE(d) case d: c = dense_q ? occupancy_dq<d, queries_per_thread_dense(d)>() \
                                        : occupancy_dq<d, queries_per_thread(d)>(); break;
            SKB_CASE(1) SKB_CASE(2) SKB_CASE(3) SKB_CASE(4) SKB_CASE(5) SKB_CASE(6) SKB_CASE(7) SKB_CASE(8)
            SKB_CASE(9) SKB_CASE(10) SKB_CASE(11) SKB_CASE(12) SKB_CASE(13) SKB_CASE(14) SKB_CASE(15) SKB_CASE(16)
#undef SKB_CASE
        }
    }
    return c;
}

static_assert(sizeof(ScoreLaunch) <= 4000, "ScoreLaunch must fit the classic 4 KB kernel-parameter space");

cudaError_t launch_score(int D, bool dense_q, const ScoreLaunch& L, int njobs, long long max_rows, int max_split,
                         int max_ntiles, cudaStream_t st)
{
    switch (D) {
#define SKB_CASE(d) case d: return dense_q ? launch_score_dq<d, queries_per_thread_dense(d)>(L, njobs, max_rows, max_split, max_ntiles, st) \
                                           : launch_score_dq<d, queries_per_thread(d)>(L, njobs, max_rows, max_split, max_ntiles, st);
        SKB_CASE(1) SKB_CASE(2) SKB_CASE(3) SKB_CASE(4) SKB_CASE(5) SKB_CASE(6) SKB_CASE(7) SKB_CASE(8)
        SKB_CASE(9) SKB_CASE(10) SKB_CASE(11) SKB_CASE(12) SKB_CASE(13) SKB_CASE(14) SKB_CASE(15) SKB_CASE(16)
#undef SKB_CASE
    }
    return cudaErrorInvalidValue;
}

}  // namespace skb

// The culled score path for sm_100a: plan -> bucket -> evaluate -> finalize.
//
// What the path needs.  A query's log-density is carried by the training tiles (kd leaves) whose bounding box lies within
// ~38 binades of its largest weighted term; in 8-D that is ~1.3 % of the leaves, and DIFFERENT leaves for different
// queries — 2^20 queries in 8-D are sparse: even 256 Morton neighbours span more than the kernel radius, so any grouping
// of queries (a warp, a CTA) needs a tile for only ~10 % of its members.  The all-pairs kernel (skb_score.cu: lanes =
// queries, tile words broadcast from shared memory) runs the pair arithmetic at 0.82 of the FP32-pipe roofline but evaluates
// a tile for every query of the warp; round 1's per-query kernel evaluated only needed (query, tile) pairs but with lanes =
// two training points (query rows re-read, shared-memory partial sums, per-tile ring traffic): 0.30 of the roofline; a
// CTA-level work-list kernel (profiles/r02d_*) filled its lanes only half and sat on barriers.  So the incidence is
// sorted GLOBALLY, by tile:
//
//   plan      one warp per 32 queries of the visit order: fused fp64 whitening, exponent reference (home super-box scan;
//             queries far from everything run an exact, warp-cooperative branch and bound), then per chunk of 256 tiles the
//             chunk-box / super-box tests with one query per lane and the tile tests with one TILE per lane -> entries
//             (tile, query warp, 32-bit mask of the warp's queries that need the tile), appended to a global list
//   bucket    counting sort of the entries by tile (per-CTA shared-memory histograms, one scan, scatter)
//   evaluate  a warp takes 64 consecutive sorted entries (almost always one tile), stages the tile in shared memory and
//             runs the all-pairs inner loop — 4 points per step, packed FADD2/FFMA2 + MUFU.EX2, tile words broadcast —
//             with TWO needing queries per lane; every lane is busy and nothing but the needed pairs is evaluated.  A
//             (query, tile) sum is added to the query's 128-bit FIXED-POINT accumulator with integer atomics: exact and
//             order independent, so the result is a function of the query and the fitted KDE alone (any batching /
//             sharding gives the same bits) although the evaluation order is not
//   finalize  accumulators -> fp64 (m, S) -> log p, ratio and accept test
//
// The exponent reference m is fixed per query by the plan step (an upper bound of the smallest weighted exponent, from
// the home super-box).  The accumulator spans 2^-128 .. 2^128 relative to 2^m, i.e. points up to ~190 binades closer than
// the reference are simply absorbed.  A (query, tile) sum beyond that (the home estimate missed the true nearest point
// by more than 190 binades: far-tail queries) flags the query with the tile's exact minimum (atomicMin, order
// independent); flagged queries are re-referenced to that minimum — it IS the global one, every unflagged tile being
// farther — and their entries are evaluated again: a superset of what the new reference needs.
#include "skb_score_common.cuh"
#include <cub/device/device_scan.cuh>

namespace skb {

constexpr float SUM_LIMIT = 1.6225928e32f;         // 2^107: a (query, tile) sum at or above this re-references the query
constexpr float NEWBEST_NONE = 3.0e38f;           // newbest[] holds 0x7f7f7f7f (3.39e38) until a query is flagged
constexpr int PLAN_WARPS = 4;                     // warps per CTA of the plan kernel (independent of each other)
constexpr int ENTRY_BLOCK = 256;                  // entries a plan warp reserves at a time (one global atomic per block)
#ifndef SKB_EVAL_EPL
#define SKB_EVAL_EPL 4
#endif
constexpr int EVAL_EPL = SKB_EVAL_EPL;            // sorted entries per lane of an evaluate work unit
constexpr int EVAL_EB = 32 * EVAL_EPL;            // ... = entries per work unit: ~7 pairs each, i.e. ~14 full passes
#ifndef SKB_EVAL_WARPS
#define SKB_EVAL_WARPS 8
#endif
#ifndef SKB_EVAL_MINB_LO
#define SKB_EVAL_MINB_LO 3      // resident CTAs per SM of the evaluate kernel for D <= 8 (<= 85 registers) ...
#endif
#ifndef SKB_EVAL_MINB_HI
#define SKB_EVAL_MINB_HI 2      // ... and above
#endif
__host__ __device__ constexpr int eval_minb(int D) { return D <= 8 ? SKB_EVAL_MINB_LO : SKB_EVAL_MINB_HI; }
#ifndef SKB_EVAL_STEP8
#define SKB_EVAL_STEP8 8        // > 0: 8 points per inner step (eight independent FMA chains per lane) up to this dimension
#endif
constexpr int EVAL_WARPS = SKB_EVAL_WARPS;
constexpr int NB_CTAS = 128;                      // CTAs of the bucket kernels (slices of the entry list)
constexpr int MAX_BUCKETS = 48 * 1024;            // tiles per bucket pass (shared-memory histogram)
constexpr int ACC_LIMBS = 4;                      // 256-bit fixed-point accumulator per query, LSB 2^-128 (relative to 2^m)

__host__ __device__ constexpr int sp_qstride(int D) { return (D + 4) & ~3; }                  // z[0..D), -m, padding to 16 B
__host__ __device__ constexpr int sp_tile_floats(int D) { return (D + 1) * tile_points(D); }  // coordinates + weights

// acc += s, exactly: s = mant 2^(e-150) lands at bit e - 22 of the accumulator (unit 2^-128); integer atomics with carry
// propagation are associative, so the sum does not depend on the order in which (query, tile) sums arrive.  Callers pass
// sums below SUM_LIMIT = 2^107 only, so the top limb cannot overflow.
__device__ __forceinline__ void fx_add(unsigned long long* __restrict__ acc, float s)
{
    if (!(s > 0.f)) return;
    const unsigned bits = __float_as_uint(s);
    const int e = (int)((bits >> 23) & 0xffu);
    unsigned long long mant = (unsigned long long)(bits & 0x7fffffu) | (e ? 0x800000ull : 0ull);
    int p = (e ? e : 1) - 22;
    if (p < 0) {                                   // below 2^-128: the low bits fall off (at least 38 bits under the reference term)
        mant = (-p >= 24) ? 0ull : (mant >> (-p));
        p = 0;
    }
    if (mant == 0ull) return;
    int limb = p >> 6;
    const int sh = p & 63;
    const unsigned long long lo = mant << sh;
    unsigned long long hi = sh > 40 ? (mant >> (64 - sh)) : 0ull;
    const unsigned long long old = atomicAdd(acc + limb, lo);
    if (old + lo < old) hi += 1ull;
    for (limb++; hi != 0ull && limb < ACC_LIMBS; limb++) {
        const unsigned long long o2 = atomicAdd(acc + limb, hi);
        hi = (o2 + hi < o2) ? 1ull : 0ull;
    }
}

// position of the n-th (0-based) set bit of w; n < popc(w)
__device__ __forceinline__ int nth_set_bit(unsigned w, unsigned n)
{
    int pos = 0;
    unsigned c = __popc(w & 0xffffu);
    if (n >= c) { n -= c; pos = 16; w >>= 16; }
    c = __popc(w & 0xffu);
    if (n >= c) { n -= c; pos += 8; w >>= 8; }
    c = __popc(w & 0xfu);
    if (n >= c) { n -= c; pos += 4; w >>= 4; }
    c = __popc(w & 0x3u);
    if (n >= c) { n -= c; pos += 2; w >>= 2; }
    if (n >= (w & 1u)) pos += 1;
    return pos;
}

// Exponent reference of ONE query (coordinates broadcast to the warp), warp cooperative with lanes = boxes / points:
//   (1) the super-box nearest to the query by box distance (exact branch and bound over chunk boxes -> super-boxes): "home";
//   (2) its tiles are scanned in order of box distance for the smallest weighted exponent |c dz|^2 - log2 w, as long as a
//       tile's box can still beat the best point found.
// A real point's exponent is an upper bound of the true minimum, which is all the window needs for accuracy; a tight
// one (the nearest neighbour usually lives in the home super-box) is what lets the box tests cull.  Should a point far
// below it turn up later, the evaluate step detects the overflow and the query is re-referenced exactly (see there).
template <int D>
__device__ __forceinline__ float coop_home_reference(const KdeDev* __restrict__ kde, int ntiles, const float (&zq)[D], int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    const bool has_w = kde->has_w != 0;
    const int nchunk = kde->nchunk, nsuper = kde->nsuper;
    // nearest chunk box
    float cb = INFINITY;
    int ci = 0;
#pragma unroll 1
    for (int c0 = 0; c0 < nchunk; c0 += 32) {
        const int c = c0 + lane;
        const float lb = c < nchunk ? box_lb<D>(kde->cboxes + (size_t)c * 2 * D, zq) : INFINITY;
        if (lb < cb) { cb = lb; ci = c; }
    }
    warp_argmin(cb, ci);
    // nearest super-box: the nearest chunk's first, then every chunk whose box is closer than the best super-box so far
    float sb = INFINITY;
    int si = ci * CG;
    auto scan_chunk = [&](int c) {
        const int s = c * CG + (lane & 15);
        float lb = (lane < CG && s < nsuper) ? box_lb<D>(kde->sboxes + (size_t)s * 2 * D, zq) : INFINITY;
        int idx = s;
        warp_argmin(lb, idx);
        if (lb < sb) { sb = lb; si = idx; }
    };
    scan_chunk(ci);
#pragma unroll 1
    for (int c0 = 0; c0 < nchunk; c0 += 32) {
        const int c = c0 + lane;
        const float lb = (c < nchunk && c != ci) ? box_lb<D>(kde->cboxes + (size_t)c * 2 * D, zq) : INFINITY;
        unsigned m = __ballot_sync(FULL, lb < sb);
        while (m) {
            const int cc = c0 + __ffs(m) - 1;
            m &= m - 1u;
            scan_chunk(cc);
        }
    }
    // tiles of the home super-box, nearest box first
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
        best = fminf(best, coop_scan_tile<D>(kde->tiles + (size_t)(si * SG + (l & 15)) * stage_floats(D), zq, has_w, lane));
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(32 * PLAN_WARPS)
skb_plan_kernel(const __grid_constant__ ScoreLaunch L, const SparseLaunch* __restrict__ SLp)
{
    constexpr int QS = sp_qstride(D);
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ __align__(16) float rows_s[PLAN_WARPS][32 * QS];

    const ScoreJob& job = L.jobs[blockIdx.y];
    const SparseLaunch& SL = *SLp;
    const SparseJob& sj = SL.sj[job.fin_slot];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long M = job.M;
    const long long vend = SL.vend < M ? SL.vend : M;
    const long long wbase = SL.vbeg + ((long long)blockIdx.x * PLAN_WARPS + wid) * 32;
    if (wbase >= vend) return;                              // warps are independent: no CTA barrier below
    float* rows = rows_s[wid];

    const KdeDev* __restrict__ kde = job.kde;
    const int ntiles = kde->ntiles;
    const float ref_bias = kde->ref_bias;

    // ---- this lane's query: whitening, exponent reference ----
    bool alive = wbase + lane < vend;
    float q2[1][D];
    load_queries<D, 1>(job, kde, wbase, lane, M, q2);
    bool isnan_row = false;
#pragma unroll
    for (int k = 0; k < D; k++) isnan_row |= (q2[0][k] != q2[0][k]);
    float mref = 0.f;
    {
        // one query at a time, the whole warp on it (lanes = boxes, then lanes = points)
        unsigned todo = __ballot_sync(FULL, alive && !isnan_row);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            float zq[D];
#pragma unroll
            for (int k = 0; k < D; k++) zq[k] = __shfl_sync(FULL, q2[0][k], src);
            float best = coop_home_reference<D>(kde, ntiles, zq, lane);
            if (!(best < 1.0e30f)) best = ref_bias;                  // inf rows (or padding only): finite reference
            if (lane == src) mref = floorf(best) - ref_bias;
        }
    }
    // rows past the end of the table and NaN rows need nothing (a NaN row's -m is NaN: finalize propagates it)
    const float negm = !alive ? INFINITY : (isnan_row ? NAN : -mref);
    alive = alive && !isnan_row;
#pragma unroll
    for (int k = 0; k < D; k++) rows[lane * QS + k] = q2[0][k];
    rows[lane * QS + D] = negm;
#pragma unroll
    for (int k = D + 1; k < QS; k++) rows[lane * QS + k] = 0.f;
    __syncwarp();
    if (wbase + lane < vend) {
        float4* dst = reinterpret_cast<float4*>(sj.qz + (size_t)(wbase + lane) * QS);
        const float4* src = reinterpret_cast<const float4*>(rows + lane * QS);
#pragma unroll
        for (int k = 0; k < QS / 4; k++) dst[k] = src[k];
    }

    // ---- entries: (tile, this warp, mask of its queries that need the tile) ----
    const unsigned my_warp = (unsigned)(wbase >> 5);
    const unsigned bucket0 = (unsigned)sj.tile_base;
    unsigned ebase = 0xffffffffu, eused = ENTRY_BLOCK;      // no block reserved yet
    unsigned n_entries = 0;
    auto emit = [&](bool has, unsigned tile, unsigned mask) {
        const unsigned nz = __ballot_sync(FULL, has);
        const unsigned c = __popc(nz);
        if (c == 0) return;
        if (eused + c > (unsigned)ENTRY_BLOCK) {
            // pad the rest of the current block with empty entries, reserve the next one
            if (ebase != 0xffffffffu)
                for (unsigned e = eused + lane; e < (unsigned)ENTRY_BLOCK; e += 32) {
                    SL.ent_tile[ebase + e] = bucket0; SL.ent_warp[ebase + e] = 0u; SL.ent_mask[ebase + e] = 0u;
                }
            unsigned nb = 0;
            if (lane == 0) nb = (unsigned)atomicAdd(SL.counter, (unsigned long long)ENTRY_BLOCK);
            nb = __shfl_sync(FULL, nb, 0);
            if ((unsigned long long)nb + ENTRY_BLOCK > (unsigned long long)SL.cap) {     // cannot happen with the worst-case capacity
                if (lane == 0) atomicExch(SL.overflow, 1u);
                ebase = 0xffffffffu; eused = ENTRY_BLOCK;
                return;
            }
            ebase = nb; eused = 0;
        }
        if (has) {
            const unsigned pos = ebase + eused + __popc(nz & ((1u << lane) - 1u));
            SL.ent_tile[pos] = bucket0 + tile; SL.ent_warp[pos] = my_warp; SL.ent_mask[pos] = mask;
        }
        eused += c;
        n_entries += c;
    };

    const float* __restrict__ cbx = kde->cboxes;
    const float* __restrict__ sbx = kde->sboxes;
    const float* __restrict__ bx = kde->boxes;
    const int nchunk = kde->nchunk, nsuper = kde->nsuper;
    // {lo, hi} of box b as packed pairs, or "empty" (no query can need it)
    auto load_box = [&](const float* __restrict__ b, bool in_range, u64 (&box2)[D]) -> bool {
        float lo[D], hi[D];
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
#pragma unroll
        for (int k = 0; k < D; k++) box2[k] = pk(lo[k], hi[k]);
        return in_range && (lo[0] <= hi[0]);                              // (inf, -inf): a box without weight
    };
    // this lane's box against the queries of qm (walked highest bit first; the warp iterates until every lane is done)
    auto test_box = [&](const u64 (&box2)[D], unsigned qm) -> unsigned {
        unsigned out = 0u;
        while (__any_sync(FULL, qm != 0u)) {
            const bool on = qm != 0u;
            const int i = on ? 31 - __clz(qm) : 0;
            qm &= ~(1u << i);                                             // 0 stays 0
            const float* qrow = rows + i * QS;
            float r[QS];
#pragma unroll
            for (int k = 0; k < QS; k += 4) {
                const float4 v = *reinterpret_cast<const float4*>(qrow + k);
                r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
            }
            float lb = 0.f;
#pragma unroll
            for (int k = 0; k < D; k++) {
                // (a packed {z - lo, z - hi} costs more than it saves here: the result pair has to be moved around)
                float blo, bhi;
                upk(box2[k], blo, bhi);
                const float gap = fmaxf(fmaxf(blo - r[k], r[k] - bhi), 0.f);
                lb = fmaf(gap, gap, lb);
            }
            if (on && !(lb + r[D] > CULL_MARGIN)) out |= 1u << i;
        }
        return out;
    };
#pragma unroll 1
    for (int c = 0; c < nchunk; c += 2) {
        // chunk boxes, one query per lane, two chunks per round
        const bool needA = alive && !(box_lb<D>(cbx + (size_t)c * 2 * D, q2[0]) + negm > CULL_MARGIN);
        const bool needB = alive && c + 1 < nchunk && !(box_lb<D>(cbx + (size_t)(c + 1) * 2 * D, q2[0]) + negm > CULL_MARGIN);
        const unsigned cmA = __ballot_sync(FULL, needA), cmB = __ballot_sync(FULL, needB);
        if ((cmA | cmB) == 0u) continue;
        // super-box tests, one SUPER-BOX per lane (lanes 0-15: chunk c, lanes 16-31: chunk c + 1) against the queries that
        // need its chunk: lane j ends up with the mask of queries that need super-box j
        unsigned smask;
        {
            const int sidx = (c + (lane >> 4)) * CG + (lane & 15);
            u64 box2[D];
            const bool ok = load_box(sbx + (size_t)(sidx < nsuper ? sidx : nsuper - 1) * 2 * D, sidx < nsuper, box2);
            smask = test_box(box2, ok ? (lane < 16 ? cmA : cmB) : 0u);
        }
        // tile tests, one TILE per lane (two super-boxes per pass), against the queries of its super-box's mask
        unsigned act = __ballot_sync(FULL, smask != 0u);
        while (act) {
            const int ja = __ffs(act) - 1;
            act &= act - 1u;
            int jb = -1;
            if (act) { jb = __ffs(act) - 1; act &= act - 1u; }
            const int j = (lane < 16) ? ja : jb;
            const int jj = j < 0 ? 0 : j;
            const unsigned qm = __shfl_sync(FULL, smask, jj);
            const int t = ((c + (jj >> 4)) * CG + (jj & 15)) * SG + (lane & 15);
            u64 box2[D];
            const bool ok = load_box(bx + (size_t)(t < ntiles ? t : ntiles - 1) * 2 * D, j >= 0 && t < ntiles, box2);
            const unsigned tmask = test_box(box2, ok ? qm : 0u);
            emit(tmask != 0u, (unsigned)t, tmask);
        }
    }
    // pad the tail of the last block
    if (ebase != 0xffffffffu)
        for (unsigned e = eused + lane; e < (unsigned)ENTRY_BLOCK; e += 32) {
            SL.ent_tile[ebase + e] = bucket0; SL.ent_warp[ebase + e] = 0u; SL.ent_mask[ebase + e] = 0u;
        }
    if (job.stats && lane == 0) {
        atomicAdd(&job.stats[2], (unsigned long long)((unsigned)ntiles - n_entries));   // tiles no query of the warp needs
        atomicAdd(&job.stats[3], (unsigned long long)ntiles);
    }
}

// ---------------------------------------------------------------------------------------------
// bucket: counting sort of the entries by tile, buckets [b_lo, b_hi) per pass
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void entry_slice(const SparseLaunch& SL, unsigned& a, unsigned& b)
{
    unsigned long long n = *SL.counter;
    if (n > SL.cap) n = SL.cap;
    const unsigned per = (unsigned)((n + NB_CTAS - 1) / NB_CTAS);
    a = blockIdx.x * per;
    b = a + per;
    if (a > n) a = (unsigned)n;
    if (b > n) b = (unsigned)n;
}

__global__ void __launch_bounds__(1024) skb_hist_kernel(const SparseLaunch* __restrict__ SLp, unsigned b_lo, unsigned b_hi, int redo)
{
    extern __shared__ unsigned sh[];
    const SparseLaunch& SL = *SLp;
    if (redo && *SL.nflag == 0u) return;
    const unsigned nb = b_hi - b_lo;
    for (unsigned i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    unsigned a, b;
    entry_slice(SL, a, b);
    for (unsigned i = a + threadIdx.x; i < b; i += blockDim.x) {
        const unsigned t = SL.ent_tile[i];
        if (t >= b_lo && t < b_hi && SL.ent_mask[i] != 0u) atomicAdd(&sh[t - b_lo], 1u);
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < nb; i += blockDim.x) SL.hist[(size_t)i * NB_CTAS + blockIdx.x] = sh[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) SL.hist[(size_t)nb * NB_CTAS] = 0u;      // the scan leaves the total here
}

__global__ void __launch_bounds__(1024) skb_scatter_kernel(const SparseLaunch* __restrict__ SLp, unsigned b_lo, unsigned b_hi, int redo)
{
    extern __shared__ unsigned sh[];
    const SparseLaunch& SL = *SLp;
    if (redo && *SL.nflag == 0u) return;
    const unsigned nb = b_hi - b_lo;
    for (unsigned i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = SL.hist[(size_t)i * NB_CTAS + blockIdx.x];
    __syncthreads();
    unsigned a, b;
    entry_slice(SL, a, b);
    for (unsigned i = a + threadIdx.x; i < b; i += blockDim.x) {
        const unsigned t = SL.ent_tile[i];
        const unsigned m = SL.ent_mask[i];
        if (t >= b_lo && t < b_hi && m != 0u) {
            const unsigned pos = atomicAdd(&sh[t - b_lo], 1u);
            SL.srt_tile[pos] = t; SL.srt_warp[pos] = SL.ent_warp[i]; SL.srt_mask[pos] = m;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// evaluate
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(32 * EVAL_WARPS, eval_minb(D))
skb_eval_kernel(const SparseLaunch* __restrict__ SLp, unsigned nbuckets_pass, int redo)
{
    constexpr int TN = tile_points(D);
    constexpr int QS = sp_qstride(D);
    constexpr int TILE_FLOATS = sp_tile_floats(D);
    constexpr int STAGE_FLOATS = stage_floats(D);
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* slot = reinterpret_cast<float*>(smem_raw) + (size_t)wid * TILE_FLOATS;                 // this warp's tile
    unsigned* tab = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(smem_raw) + (size_t)EVAL_WARPS * TILE_FLOATS) + wid * 4 * EVAL_EB;
    unsigned* s_scan = tab;                      // inclusive pair counts
    unsigned* s_tile = tab + EVAL_EB;
    unsigned* s_warp = tab + 2 * EVAL_EB;
    unsigned* s_mask = tab + 3 * EVAL_EB;

    const SparseLaunch& SL = *SLp;
    if (redo && *SL.nflag == 0u) return;                                  // nothing was flagged: the usual case
    const unsigned n = SL.hist[(size_t)nbuckets_pass * NB_CTAS];          // sorted entries of this pass (total of the scan)
    const unsigned nblk = (n + EVAL_EB - 1) / EVAL_EB;
    const unsigned gw = blockIdx.x * EVAL_WARPS + wid, nw = gridDim.x * EVAL_WARPS;
    unsigned long long n_pairs = 0;

#pragma unroll 1
    for (unsigned blk = gw; blk < nblk; blk += nw) {
        __syncwarp();
        // ---- the block's entries (EVAL_EPL per lane) and the inclusive scan of their pair counts ----
        unsigned pc[EVAL_EPL];
#pragma unroll
        for (int h = 0; h < EVAL_EPL; h++) {
            const unsigned e = blk * EVAL_EB + h * 32 + lane;
            const bool in = e < n;
            unsigned m = in ? SL.srt_mask[e] : 0u;
            const unsigned tl = in ? SL.srt_tile[e] : 0xffffffffu;
            const unsigned wp = in ? SL.srt_warp[e] : 0u;
            if (redo && m) {                                              // second round: only the flagged queries of each entry
                int jsel = 0;
#pragma unroll 1
                for (int j = 1; j < SL.njobs; j++)
                    if (SL.sj[SL.jobs_of_group[j]].tile_base <= (int)tl) jsel = j;
                m &= SL.sj[SL.jobs_of_group[jsel]].wflag[wp];
            }
            s_tile[h * 32 + lane] = tl;
            s_warp[h * 32 + lane] = wp;
            s_mask[h * 32 + lane] = m;
            pc[h] = __popc(m);
        }
        unsigned carry = 0u;
#pragma unroll
        for (int h = 0; h < EVAL_EPL; h++) {
            unsigned incl = pc[h];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            incl += carry;
            s_scan[h * 32 + lane] = incl;
            carry = __shfl_sync(FULL, incl, 31);
        }
        const unsigned total = carry;
        __syncwarp();

        // ---- tile by tile (sorted: the entries of a tile are contiguous, and so are their ranks) ----
        unsigned r0 = 0u;
        int e_first = 0;                              // first entry of the current tile
#pragma unroll 1
        while (r0 < total) {
            // skip entries that end at or before r0 (none in practice: padding entries never reach the sorted list)
            while (e_first < EVAL_EB && s_scan[e_first] <= r0) e_first++;
            const unsigned cur = s_tile[e_first];
            // last entry of this tile: entries are sorted, so compare each lane's entries with `cur`
            int e_last = e_first;
#pragma unroll
            for (int h = 0; h < EVAL_EPL; h++) {
                const unsigned mh = __ballot_sync(FULL, s_tile[h * 32 + lane] == cur);
                if (mh) e_last = h * 32 + (31 - __clz(mh));
            }
            const unsigned r_end = s_scan[e_last];

            // which job, which tile
            int jsel = 0;
#pragma unroll 1
            for (int j = 1; j < SL.njobs; j++)
                if (SL.sj[SL.jobs_of_group[j]].tile_base <= (int)cur) jsel = j;
            const SparseJob& sj = SL.sj[SL.jobs_of_group[jsel]];
            const float* __restrict__ gt = sj.tiles + (size_t)(cur - (unsigned)sj.tile_base) * STAGE_FLOATS;
            // stage the tile (coordinates + weights) in this warp's slot
            __syncwarp();
#pragma unroll
            for (int i = lane; i < TILE_FLOATS / 4; i += 32)
                reinterpret_cast<float4*>(slot)[i] = __ldg(reinterpret_cast<const float4*>(gt) + i);
            __syncwarp();

#pragma unroll 1
            for (unsigned r = r0; r < r_end; r += 64u) {
                // this lane's two queries: ranks ra, ra + 1 among the tile's needers (the second may fall off the end)
                const unsigned ra = r + 2u * (unsigned)lane;
                if (ra < r_end) {
                    unsigned qv[2];
                    const bool two = ra + 1u < r_end;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const unsigned rk = (h == 1 && !two) ? ra : ra + (unsigned)h;
                        // entry e in [e_first, e_last] with scan[e - 1] <= rk < scan[e]
                        int lo = e_first, hi = e_last;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (s_scan[mid] > rk) hi = mid; else lo = mid + 1;
                        }
                        const unsigned before = lo ? s_scan[lo - 1] : 0u;
                        qv[h] = s_warp[lo] * 32u + (unsigned)nth_set_bit(s_mask[lo], rk - before);
                    }
                    bool want[2];
                    want[0] = true;
                    want[1] = two;
                    if (redo) {                                           // second round: flagged queries only
                        want[0] = sj.newbest[qv[0]] < NEWBEST_NONE;
                        want[1] = two && sj.newbest[qv[1]] < NEWBEST_NONE;
                        if (!want[0] && !want[1]) continue;
                    }
                    float qq[2][D], nm[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const float4* qrow = reinterpret_cast<const float4*>(sj.qz + (size_t)qv[h] * QS);
                        float rr[QS];
#pragma unroll
                        for (int k = 0; k < QS / 4; k++) {
                            const float4 v = qrow[k];                       // (not __ldg: the fix-up kernel rewrites -m between rounds)
                            rr[4 * k] = v.x; rr[4 * k + 1] = v.y; rr[4 * k + 2] = v.z; rr[4 * k + 3] = v.w;
                        }
#pragma unroll
                        for (int k = 0; k < D; k++) qq[h][k] = rr[k];
                        nm[h] = rr[D];
                    }
                    u64 sum2[2];
                    float unused[2];
                    sum2[0] = pk(0.f, 0.f);
                    sum2[1] = pk(0.f, 0.f);
#if SKB_EVAL_STEP8
                    if constexpr (D <= SKB_EVAL_STEP8) {
#pragma unroll 1
                        for (int pt = 0; pt < TN; pt += 8) step_compute8<D, 2, false>(slot, pt, qq, nm, sum2, unused);
                    } else
#endif
                    {
#pragma unroll 1
                        for (int pt = 0; pt < TN; pt += 4) {
                            StepOperands<D> A;
                            A.template load<true>(slot, pt);
                            step_compute<D, 2, false, false>(A, qq, nm, sum2, unused);
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (!want[h]) continue;
                        float sl, sh_;
                        upk(sum2[h], sl, sh_);
                        const float sm = sl + sh_;
                        if (sm < SUM_LIMIT) {
                            fx_add(sj.acc + (size_t)qv[h] * ACC_LIMBS, sm);
                        } else if (!redo) {
                            // a point far closer than the reference allows (or inf / NaN arithmetic): the tile's exact minimum
                            // of the weighted exponent goes to newbest[] (absolute, >= 0: ordered like its bit pattern)
                            float amin = INFINITY;
#pragma unroll 1
                            for (int pt = 0; pt < TN; pt++) {
                                float a = fminf(-__log2f(slot[D * TN + pt]), LW_CAP);
#pragma unroll
                                for (int k = 0; k < D; k++) {
                                    const float dz = qq[h][k] - slot[k * TN + pt];
                                    a = fmaf(dz, dz, a);
                                }
                                amin = fminf(amin, a);
                            }
                            if (amin < NEWBEST_NONE) {
                                atomicMin(reinterpret_cast<int*>(sj.newbest + qv[h]), __float_as_int(fmaxf(amin, 0.f)));
                                atomicOr(sj.wflag + (qv[h] >> 5), 1u << (qv[h] & 31u));
                                atomicAdd(SL.nflag, 1u);
                            } else {
                                atomicExch(SL.overflow, 1u);               // NaN arithmetic on a finite row: fail loudly
                            }
                        } else {
                            atomicExch(SL.overflow, 1u);                   // cannot happen after re-referencing
                        }
                    }
                    n_pairs += (unsigned)TN * ((want[0] ? 1u : 0u) + (want[1] ? 1u : 0u));
                }
            }
            r0 = r_end;
            e_first = e_last + 1;
        }
    }
    if (SL.stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_pairs += __shfl_xor_sync(FULL, n_pairs, o);
        if (lane == 0 && n_pairs) atomicAdd(&SL.stats[1], n_pairs / 32ull);      // counters are "per lane at warp level": x 32 = pairs
    }
}

// ---------------------------------------------------------------------------------------------
// fix-up between the two evaluate rounds: flagged queries get the exact reference and a clean accumulator
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) skb_refix_kernel(const SparseLaunch* __restrict__ SLp)
{
    const SparseLaunch& SL = *SLp;
    if (*SL.nflag == 0u) return;
    const SparseJob& sj = SL.sj[SL.jobs_of_group[blockIdx.y]];
    const long long vend = SL.vend < sj.M ? SL.vend : sj.M;
    const long long v = SL.vbeg + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= vend) return;
    const float nb = sj.newbest[v];
    if (!(nb < NEWBEST_NONE)) return;
    sj.qz[(size_t)v * sj.qs + sj.D] = -(floorf(nb) - sj.ref_bias);
#pragma unroll
    for (int k = 0; k < ACC_LIMBS; k++) sj.acc[(size_t)v * ACC_LIMBS + k] = 0ull;
}

// ---------------------------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
skb_sparse_fin_kernel(const __grid_constant__ ScoreLaunch L, const SparseLaunch* __restrict__ SLp, int first, int count, long long M)
{
    const SparseLaunch& SL = *SLp;
    const int f0 = first + (count == 1 ? (int)blockIdx.y : 0);            // ungrouped launches finalize one job per blockIdx.y
    const long long Mj = count == 1 ? SL.sj[f0].M : M;
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= Mj) return;
    const bool overflow = *SL.overflow != 0u;
    if (count == 1 && SL.sj[f0].acc == nullptr) return;                    // that job finalized itself (round-1 kernel, ungrouped)
    finalize_query_with(L.fin, f0, count, v, [&](int f, long long vv, double& mstar, double& S) {
        const SparseJob& sj = SL.sj[f];
        if (sj.acc == nullptr) {                                            // group member scored by a round-1 kernel
            merge_partials(L.fin.fj[f].partial, L.fin.fj[f].nsplit, sj.M, vv, mstar, S);
            return;
        }
        const float negm = sj.qz[(size_t)vv * sj.qs + sj.D];
        const unsigned long long* a = sj.acc + (size_t)vv * ACC_LIMBS;
        // limbs are worth 2^-128, 2^-64, 1, 2^64: summed top down in fp64
        S = __dadd_rn(__dadd_rn(__dmul_rn((double)a[3], 18446744073709551616.0), (double)a[2]),
                      __dadd_rn(__dmul_rn((double)a[1], 5.421010862427522e-20), __dmul_rn((double)a[0], 2.938735877055719e-39)));
        mstar = -(double)negm;
        if (overflow || negm != negm) { S = NAN; mstar = 0.0; }
    });
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int D>
static cudaError_t launch_group_d(const ScoreLaunch& L, int nl, const SparseLaunch& SLh, const SparseLaunch* SLd,
                                  long long max_rows, int nbuckets, void* scan_tmp, size_t scan_tmp_bytes, int sms, cudaStream_t st)
{
    cudaError_t e;
    const long long nwarps = (max_rows + 31) / 32;
    dim3 pg((unsigned)((nwarps + PLAN_WARPS - 1) / PLAN_WARPS), (unsigned)nl, 1);
    skb_plan_kernel<D><<<pg, 32 * PLAN_WARPS, 0, st>>>(L, SLd);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const size_t eval_smem = (size_t)EVAL_WARPS * sp_tile_floats(D) * 4 + (size_t)EVAL_WARPS * 4 * EVAL_EB * 4;
    if ((e = cudaFuncSetAttribute(skb_eval_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eval_smem)) != cudaSuccess) return e;
    const int npass = (nbuckets + MAX_BUCKETS - 1) / MAX_BUCKETS;
    // round 0 evaluates everything; round 1 re-evaluates the queries flagged in round 0 (every kernel of round 1 returns at
    // once when nothing was flagged; with a single bucket pass the sorted list of round 0 is still in place)
    for (int redo = 0; redo < 2; redo++) {
        if (redo) {
            dim3 fg((unsigned)((max_rows + 255) / 256), (unsigned)nl, 1);
            skb_refix_kernel<<<fg, 256, 0, st>>>(SLd);
            count_launch();
        }
        for (int b_lo = 0; b_lo < nbuckets; b_lo += MAX_BUCKETS) {
            const int b_hi = b_lo + MAX_BUCKETS < nbuckets ? b_lo + MAX_BUCKETS : nbuckets;
            const int nb = b_hi - b_lo;
            const size_t hsmem = (size_t)nb * 4;
            if (!redo || npass > 1) {
                skb_hist_kernel<<<NB_CTAS, 1024, hsmem, st>>>(SLd, (unsigned)b_lo, (unsigned)b_hi, redo);
                // (the scan of a skipped histogram is harmless: round 1's evaluate returns before reading it)
                size_t tb = scan_tmp_bytes;
                if ((e = cub::DeviceScan::ExclusiveSum(scan_tmp, tb, SLh.hist, SLh.hist, nb * NB_CTAS + 1, st)) != cudaSuccess) return e;
                skb_scatter_kernel<<<NB_CTAS, 1024, hsmem, st>>>(SLd, (unsigned)b_lo, (unsigned)b_hi, redo);
                count_launch(4);
            }
            skb_eval_kernel<D><<<eval_minb(D) * sms, 32 * EVAL_WARPS, eval_smem, st>>>(SLd, (unsigned)nb, redo);
            count_launch();
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

size_t sparse_scan_temp_bytes()
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (unsigned*)nullptr, (unsigned*)nullptr, MAX_BUCKETS * NB_CTAS + 1);
    return bytes;
}

cudaError_t sparse_prepare()
{
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(skb_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_BUCKETS * 4);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(skb_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_BUCKETS * 4);
    if (e != cudaSuccess) return e;
    done = true;
    return cudaSuccess;
}

cudaError_t launch_sparse_group(int D, const ScoreLaunch& L, int nl, const SparseLaunch& SLh, const SparseLaunch* SLd,
                                long long max_rows, int nbuckets, void* scan_tmp, size_t scan_tmp_bytes, int sms, cudaStream_t st)
{
    switch (D) {
#define SKB_CASE(d) case d: return launch_group_d<d>(L, nl, SLh, SLd, max_rows, nbuckets, scan_tmp, scan_tmp_bytes, sms, st);
        SKB_CASE(1) SKB_CASE(2) SKB_CASE(3) SKB_CASE(4) SKB_CASE(5) SKB_CASE(6) SKB_CASE(7) SKB_CASE(8)
        SKB_CASE(9) SKB_CASE(10) SKB_CASE(11) SKB_CASE(12) SKB_CASE(13) SKB_CASE(14) SKB_CASE(15) SKB_CASE(16)
#undef SKB_CASE
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_sparse_finalize(const ScoreLaunch& L, const SparseLaunch* SLd, int first, int count, int njobs_y,
                                   long long max_rows, cudaStream_t st)
{
    dim3 grid((unsigned)((max_rows + 255) / 256), (unsigned)njobs_y, 1);
    skb_sparse_fin_kernel<<<grid, 256, 0, st>>>(L, SLd, first, count, max_rows);
    count_launch();
    return cudaGetLastError();
}

}  // namespace skb

// The culled score kernel for sm_100a: "work item per lane" (WIL).
//
// What the path needs.  A query's log-density is carried by the training tiles (kd leaves) whose bounding box lies within
// ~38 binades of its nearest weighted term; in 8-D that is ~1.3 % of the leaves, and DIFFERENT leaves for different
// queries.  The all-pairs kernel (skb_score.cu, lanes = queries, tile words broadcast from shared memory) runs the pair
// arithmetic at 0.82 of the FP32-pipe roofline but evaluates a tile for all 32 queries of a warp as soon as one needs it;
// round 1's per-query kernel evaluated only the needed (query, tile) pairs but with lanes = 2 training points, paying ~40
// issue slots per 64 pairs (query row reloads, shared-memory partial sums, per-tile ring traffic): 0.30 of the roofline.
//
// This kernel keeps the all-pairs inner loop and feeds it exactly the needed work:
//   * a CTA owns 256 consecutive queries of the visit order (Morton neighbours) and walks the training side in chunks of
//     256 tiles (one chunk box = 16 super-boxes = one kd subtree);
//   * phase A (per warp, as before): chunk-box / super-box tests with one query per lane, then tile tests with one TILE per
//     lane against the queries of the super-box's ballot -> mask[warp][tile] = which of the warp's queries need the tile;
//   * phase B (thread t <-> tile t): per-tile counts over the CTA's 8 warps, one block scan -> a compact list of needed
//     tiles with item / result offsets.  A WORK ITEM is (tile, two of the queries that need it);
//   * phase C: the chunk's item stream is cut into windows of <= 256 items (one per thread) spanning <= S tiles; the window's
//     tiles arrive in shared memory by 1-D TMA bulk copies (double buffered, next window in flight), every thread decodes
//     its item (tile slot, two query rows), runs the 4-points-per-step packed FADD2/FFMA2/MUFU.EX2 loop over the tile for
//     its two queries with the tile words broadcast from shared memory, and writes two fp32 tile sums;
//   * the owner thread of each query folds the tile sums into its fp64 state in TILE ORDER.
// A query's result is  sum over its needed tiles, in tile order, of fp32 tile sums computed in a fixed point order — a
// function of the query and the fitted KDE alone, so any batching / sharding gives the same bits.  The exponent reference
// is fixed per query before the first tile (home super-box scan; queries far from everything run an exact branch-and-bound
// so that no term can leave the fp32 exponent range), hence no re-centring inside the loop.
#include "skb_score_common.cuh"

namespace skb {

constexpr int WIL_NW = 8;                         // warps per CTA
constexpr int WIL_P = 32 * WIL_NW;                // queries per CTA = tiles per chunk = threads
static_assert(WIL_P == CG * SG, "one thread per tile of a chunk in phase B");
constexpr float FAR_LIMIT = 150.f;                // references farther than this from every home point are refined exactly

__host__ __device__ constexpr int wil_slots(int D) { return D <= 8 ? 16 : 8; }                 // tiles per window
__host__ __device__ constexpr int wil_qstride(int D) { return (D + 4) & ~3; }                   // z[0..D), -m, padding to 16 B
__host__ __device__ constexpr int wil_tile_floats(int D) { return (D + 1) * tile_points(D); }   // coordinates + weights (no header row)
__host__ __device__ constexpr int wil_slot_floats(int D) { return wil_tile_floats(D) + 4; }     // 16 B skew: slots start in different banks
__host__ __device__ constexpr size_t wil_smem_bytes(int D)
{
    return (size_t)2 * wil_slots(D) * wil_slot_floats(D) * 4      // tile slots, double buffered
           + (size_t)WIL_P * wil_qstride(D) * 4                  // query rows
           + (size_t)WIL_NW * 256 * 4                            // masks[warp][tile]
           + (size_t)256 * 8                                     // cpre[tile][warp] (bytes)
           + (size_t)260 * 4                                     // cro
           + (size_t)260 * 2                                     // cio
           + (size_t)256                                         // ct
           + (size_t)2 * 2 * WIL_P * 4                           // results, double buffered
           + (size_t)WIL_NW * 8 + 2 * 8 + 64;                    // scan totals, mbarriers, slack
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

// Exact branch and bound for a query whose home estimate `best` is far from everything: scan every tile whose box comes
// closer than best - slack.  Afterwards every point has |c dz|^2 >= best - slack, so no term can exceed 2^(slack - bias).
template <int D>
__device__ __forceinline__ float refine_far_reference(const KdeDev* __restrict__ kde, int ntiles, const float (&zq)[D],
                                                      float best, float slack)
{
    const bool has_w = kde->has_w != 0;
    const int nchunk = kde->nchunk, nsuper = kde->nsuper;
#pragma unroll 1
    for (int c = 0; c < nchunk; c++) {
        if (!(box_lb<D>(kde->cboxes + (size_t)c * 2 * D, zq) < best - slack)) continue;
        const int send = (c + 1) * CG < nsuper ? (c + 1) * CG : nsuper;
#pragma unroll 1
        for (int s = c * CG; s < send; s++) {
            if (!(box_lb<D>(kde->sboxes + (size_t)s * 2 * D, zq) < best - slack)) continue;
            const int tend = (s + 1) * SG < ntiles ? (s + 1) * SG : ntiles;
#pragma unroll 1
            for (int t = s * SG; t < tend; t++) {
                if (!(box_lb<D>(kde->boxes + (size_t)t * 2 * D, zq) < best - slack)) continue;
                best = scan_tile_min<D>(kde->tiles + (size_t)t * stage_floats(D), zq, has_w, best);
            }
        }
    }
    return best;
}

template <int D>
__global__ void __launch_bounds__(WIL_P, 2)
skb_score_wil_kernel(const __grid_constant__ ScoreLaunch L)
{
    constexpr int TN = tile_points(D);
    constexpr int QS = wil_qstride(D);
    constexpr int S = wil_slots(D);
    constexpr int TILE_FLOATS = wil_tile_floats(D);
    constexpr int SLOT_FLOATS = wil_slot_floats(D);
    constexpr int STAGE_FLOATS = stage_floats(D);
    constexpr uint32_t TILE_BYTES = TILE_FLOATS * sizeof(float);
    constexpr unsigned FULL = 0xffffffffu;

    const ScoreJob& job = L.jobs[blockIdx.y];
    const long long M = job.M;
    const long long qbase = (long long)blockIdx.x * WIL_P;
    if (qbase >= M || (int)blockIdx.z >= job.nsplit) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* slots = reinterpret_cast<float*>(smem_raw);                               // [2][S][SLOT_FLOATS]
    float* rows = slots + 2 * S * SLOT_FLOATS;                                       // [P][QS]
    unsigned* masks = reinterpret_cast<unsigned*>(rows + WIL_P * QS);                // [NW][256]
    unsigned long long* cpre64 = reinterpret_cast<unsigned long long*>(masks + WIL_NW * 256);   // [256] 8 prefix bytes per tile
    unsigned* cro = reinterpret_cast<unsigned*>(cpre64 + 256);                       // [257] result offsets of the needed tiles
    unsigned short* cio = reinterpret_cast<unsigned short*>(cro + 260);              // [257] item offsets
    unsigned char* ct = reinterpret_cast<unsigned char*>(cio + 260);                 // [256] needed tiles (offset in the chunk)
    float* results = reinterpret_cast<float*>(ct + 256);                             // [2][2P]
    unsigned long long* wtot = reinterpret_cast<unsigned long long*>(results + 2 * 2 * WIL_P);   // [NW]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(wtot + WIL_NW);                 // [2]
    __shared__ int s_last[WIL_P / FB];
    const unsigned char* cpre = reinterpret_cast<const unsigned char*>(cpre64);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const KdeDev* __restrict__ kde = job.kde;
    const int ntiles = kde->ntiles;
    const int nsplit = job.nsplit;
    const int t0 = (int)(((long long)ntiles * blockIdx.z) / nsplit);
    const int t1 = (int)(((long long)ntiles * (blockIdx.z + 1)) / nsplit);

    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_fence_init();
    }

    // ---- P0: this thread's query: whitening, exponent reference ----
    const bool alive = qbase + tid < M;
    float q2[1][D];
    load_queries<D, 1>(job, kde, qbase, tid, M, q2);
    float mref[1], bmin[1], nm2[1];
    double Sacc[1];
    Sacc[0] = 0.0;
    home_reference<D, 1>(kde, ntiles, q2, bmin, mref, nm2, alive);
    if (alive && bmin[0] > FAR_LIMIT) {
        const float best = refine_far_reference<D>(kde, ntiles, q2[0], bmin[0], FAR_LIMIT);
        if (best < bmin[0]) { bmin[0] = best; mref[0] = floorf(best) - kde->ref_bias; }
    }
    const float negm = alive ? -mref[0] : INFINITY;          // rows past the end of the table need nothing
#pragma unroll
    for (int k = 0; k < D; k++) rows[tid * QS + k] = q2[0][k];
    rows[tid * QS + D] = negm;
    __syncthreads();

    const float* __restrict__ cbx = kde->cboxes;
    const float* __restrict__ sbx = kde->sboxes;
    const float* __restrict__ bx = kde->boxes;
    const float* __restrict__ gtiles = kde->tiles;
    const bool cull = job.cull != 0;
    unsigned wcount = 0;                                     // windows so far: ring position, uniform over the CTA
    unsigned long long n_pairs = 0;
    unsigned n_needed = 0;

    const int s_first = (t0 / SG / CG) * CG, s_end = (t1 + SG - 1) / SG;     // chunk aligned; tiles outside [t0, t1) are masked
#pragma unroll 1
    for (int s_base = s_first; s_base < s_end; s_base += CG) {
        // ---- A: which (query, tile) pairs of this chunk are needed ----
        bool need_c = alive;
        if (cull) need_c = alive && !(box_lb<D>(cbx + (size_t)(s_base / CG) * 2 * D, q2[0]) + negm > CULL_MARGIN);
        if (!__syncthreads_or(need_c ? 1 : 0)) continue;
        {
            uint4* mrow = reinterpret_cast<uint4*>(masks + wid * 256);
            mrow[lane * 2] = make_uint4(0u, 0u, 0u, 0u);
            mrow[lane * 2 + 1] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
        if (__any_sync(FULL, need_c)) {
            // super-box tests, one query per lane: lane j keeps the ballot of super-box s_base + j
            unsigned smask = 0u;
            const int nsb = s_end - s_base < CG ? s_end - s_base : CG;
#pragma unroll 2
            for (int j = 0; j < nsb; j++) {
                const float* b = sbx + (size_t)(s_base + j) * 2 * D;
                float lb = 0.f;
                bool valid = true;
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const float lo = __ldg(b + k), hi = __ldg(b + D + k);
                    if (k == 0) valid = (lo <= hi);
                    const float gap = fmaxf(fmaxf(lo - q2[0][k], q2[0][k] - hi), 0.f);
                    lb = fmaf(gap, gap, lb);
                }
                const bool need = alive && valid && (!cull || !(lb + negm > CULL_MARGIN));
                const unsigned bal = __ballot_sync(FULL, need);
                if (lane == j) smask = bal;
            }
            // tile tests, one tile per lane (two super-boxes per pass), against the queries of the super-box's ballot only
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
                    const int i = on ? 31 - __clz(qm) : 0;
                    qm &= ~(1u << i);                                         // 0 stays 0
                    const float* qrow = rows + (wid * 32 + i) * QS;
                    float lb = 0.f;
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        const float zq = qrow[k];
                        const float gap = fmaxf(fmaxf(lo[k] - zq, zq - hi[k]), 0.f);
                        lb = fmaf(gap, gap, lb);
                    }
                    if (on && (!cull || !(lb + qrow[D] > CULL_MARGIN))) tmask |= 1u << i;
                }
                if (j >= 0) masks[wid * 256 + toff] = tmask;
            }
        }
        __syncthreads();

        // ---- B: thread tid <-> tile tid of the chunk: counts over the warps, block scan, compact list ----
        unsigned long long total;
        {
            unsigned n = 0;
            unsigned long long pre = 0ull;
#pragma unroll
            for (int w = 0; w < WIL_NW; w++) {
                pre |= (unsigned long long)n << (8 * w);                      // exclusive prefix, <= 224
                n += __popc(masks[w * 256 + tid]);
            }
            cpre64[tid] = pre;
            const unsigned long long key = n ? ((1ull << 40) | ((unsigned long long)((n + 1) >> 1) << 20) | n) : 0ull;
            unsigned long long incl = key;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) wtot[wid] = incl;
            __syncthreads();
            unsigned long long base = 0ull;
            total = 0ull;
#pragma unroll
            for (int w = 0; w < WIL_NW; w++) {
                const unsigned long long v = wtot[w];
                if (w < wid) base += v;
                total += v;
            }
            const unsigned long long excl = base + incl - key;
            if (n) {
                const int j = (int)(excl >> 40);
                ct[j] = (unsigned char)tid;
                cio[j] = (unsigned short)((excl >> 20) & 0xfffffu);
                cro[j] = (unsigned)(excl & 0xfffffu);
            }
            if (tid == 0) {
                const int T = (int)(total >> 40);
                cio[T] = (unsigned short)((total >> 20) & 0xfffffu);
                cro[T] = (unsigned)(total & 0xfffffu);
            }
        }
        __syncthreads();
        const int T = (int)(total >> 40);
        if (T == 0) continue;
        n_needed += (unsigned)T;

        // ---- C: windows of <= P items spanning <= S tiles ----
        const unsigned itot = cio[T];
        const int chunk_t0 = s_base * SG;
        auto make_window = [&](unsigned i0, int ja, unsigned& i1, int& jb) {
            const int jlim = ja + S < T ? ja + S : T;
            const unsigned cap = cio[jlim];
            i1 = i0 + WIL_P < cap ? i0 + WIL_P : cap;
            jb = ja;
            while (jb + 1 < jlim && cio[jb + 1] < i1) jb++;
        };
        auto issue = [&](int buf, int ja, int jb) {                            // thread 0 only
            mbar_expect_tx(&full_bar[buf], (uint32_t)(jb - ja + 1) * TILE_BYTES);
            for (int j = ja; j <= jb; j++)
                tma_bulk_g2s(slots + (size_t)(buf * S + (j - ja)) * SLOT_FLOATS,
                             gtiles + (size_t)(chunk_t0 + ct[j]) * STAGE_FLOATS, TILE_BYTES, &full_bar[buf]);
        };
        unsigned i0 = 0, i1;
        int ja = 0, jb;
        make_window(i0, ja, i1, jb);
        if (tid == 0) issue((int)(wcount & 1u), ja, jb);
#pragma unroll 1
        for (;;) {
            const int buf = (int)(wcount & 1u);
            const unsigned par = (wcount >> 1) & 1u;
            // the next window goes in flight before this one is consumed
            const unsigned ni0 = i1;
            const bool has_next = ni0 < itot;
            const int nja = (ni0 == cio[jb + 1]) ? jb + 1 : jb;
            unsigned ni1 = 0;
            int njb = 0;
            if (has_next) {
                make_window(ni0, nja, ni1, njb);
                if (tid == 0) issue(buf ^ 1, nja, njb);
            }
            mbar_wait(&full_bar[buf], par);

            const unsigned rbase = cro[ja] + 2u * (i0 - cio[ja]);              // first result of the window
            float* res = results + buf * 2 * WIL_P;
            const unsigned i = i0 + (unsigned)tid;
            if (i < i1) {
                // decode: tile j with cio[j] <= i < cio[j + 1], then the two queries of ranks 2p, 2p + 1 among its needers
                int lo = ja, hi = jb;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (cio[mid] <= i) lo = mid; else hi = mid - 1;
                }
                const int j = lo;
                const int t = ct[j];
                const unsigned p = i - cio[j];
                const unsigned nt = cro[j + 1] - cro[j];
                const uint2 cp = *reinterpret_cast<const uint2*>(cpre + t * 8);
                int qi[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    unsigned r = 2u * p + (unsigned)e;
                    if (r >= nt) r = nt - 1u;                                  // odd count: the last item repeats its query
                    const unsigned rr = r * 0x01010101u;
                    const int w = (__popc(__vcmpgeu4(rr, cp.x)) + __popc(__vcmpgeu4(rr, cp.y))) / 8 - 1;
                    const unsigned base = ((w < 4 ? cp.x >> (8 * w) : cp.y >> (8 * (w - 4))) & 0xffu);
                    qi[e] = w * 32 + nth_set_bit(masks[w * 256 + t], r - base);
                }
                float qq[2][D], nm[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const float* qrow = rows + qi[e] * QS;
                    float r[QS];
#pragma unroll
                    for (int k = 0; k < QS; k += 4) {
                        const float4 v = *reinterpret_cast<const float4*>(qrow + k);
                        r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
                    }
#pragma unroll
                    for (int k = 0; k < D; k++) qq[e][k] = r[k];
                    nm[e] = r[D];
                }
                const float* tile = slots + (size_t)(buf * S + (j - ja)) * SLOT_FLOATS;
                u64 sum2[2];
                float unused[2];
                sum2[0] = pk(0.f, 0.f);
                sum2[1] = pk(0.f, 0.f);
#pragma unroll 1
                for (int pt = 0; pt < TN; pt += 4) {
                    StepOperands<D> A;
                    A.template load<true>(tile, pt);
                    step_compute<D, 2, false, false>(A, qq, nm, sum2, unused);
                }
                float s0l, s0h, s1l, s1h;
                upk(sum2[0], s0l, s0h);
                upk(sum2[1], s1l, s1h);
                const unsigned ridx = cro[j] + 2u * p - rbase;
                res[ridx] = s0l + s0h;
                if (2u * p + 1u < nt) res[ridx + 1] = s1l + s1h;
                n_pairs += (unsigned)TN * (2u * p + 1u < nt ? 2u : 1u);
            }
            __syncthreads();
            // fold: the owner of each query adds its tile sums of this window in tile order
#pragma unroll 1
            for (int j = ja; j <= jb; j++) {
                const int t = ct[j];
                const unsigned word = masks[wid * 256 + t];
                if ((word >> lane) & 1u) {
                    const unsigned rank = (unsigned)cpre[t * 8 + wid] + __popc(word & ((1u << lane) - 1u));
                    const unsigned c0 = cio[j], c1 = cio[j + 1];
                    const unsigned plo = i0 > c0 ? i0 - c0 : 0u;
                    const unsigned phi = (i1 < c1 ? i1 : c1) - c0;
                    if (rank >= 2u * plo && rank < 2u * phi) Sacc[0] += (double)res[cro[j] + rank - rbase];
                }
            }
            wcount++;
            if (!has_next) break;
            i0 = ni0; ja = nja; i1 = ni1; jb = njb;
        }
    }

    if (job.stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_pairs += __shfl_xor_sync(FULL, n_pairs, o);
        if (lane == 0) atomicAdd(&job.stats[1], n_pairs / 32ull);           // counters are "per lane at warp level": x 32 = pairs
        if (tid == 0) {
            atomicAdd(&job.stats[2], (unsigned long long)((unsigned)(t1 - t0) - n_needed));
            atomicAdd(&job.stats[3], (unsigned long long)(t1 - t0));
        }
    }
    finish_queries<D, 1, WIL_P>(L, job, qbase, tid, M, mref, Sacc, s_last);
}

template <int D>
static cudaError_t launch_wil_d(const ScoreLaunch& L, int njobs, long long max_rows, int max_split, cudaStream_t st)
{
    const size_t smem = wil_smem_bytes(D);
    cudaError_t e = cudaFuncSetAttribute(skb_score_wil_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((max_rows + WIL_P - 1) / WIL_P), (unsigned)njobs, (unsigned)max_split);
    skb_score_wil_kernel<D><<<grid, WIL_P, smem, st>>>(L);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_score_wil(int D, const ScoreLaunch& L, int njobs, long long max_rows, int max_split, cudaStream_t st)
{
    switch (D) {
#define SKB_CASE(d) case d: return launch_wil_d<d>(L, njobs, max_rows, max_split, st);
        SKB_CASE(1) SKB_CASE(2) SKB_CASE(3) SKB_CASE(4) SKB_CASE(5) SKB_CASE(6) SKB_CASE(7) SKB_CASE(8)
        SKB_CASE(9) SKB_CASE(10) SKB_CASE(11) SKB_CASE(12) SKB_CASE(13) SKB_CASE(14) SKB_CASE(15) SKB_CASE(16)
#undef SKB_CASE
    }
    return cudaErrorInvalidValue;
}

}  // namespace skb

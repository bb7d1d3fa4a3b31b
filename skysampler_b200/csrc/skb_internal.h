// Internal declarations shared by the sm_100a kernels and the C-ABI glue (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/skb.h"

namespace skb {

typedef unsigned long long u64;

// ---- training-tile geometry -------------------------------------------------------------------
// The fitted KDE lives in HBM as tiles of TN = tile_points(D) points in SoA order: tile t is one contiguous block
//   [D+2][TN] fp32 = D rows of prescaled whitened coordinates + 1 row of normalised weights + 1 header
//   row whose first 2D words are the tile's bounding box (lo[D], hi[D]),
// so that ONE 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) moves a whole tile into shared memory
// and an LDS.128 of row k yields two packed {x_j, x_j+1} operands for FADD2/FFMA2.
// Tiles are the leaves of a kd-tree over the training points (built on the device at fit time), so a
// tile is spatially compact and its box test can prove that every pair term flushes to zero.
#ifndef SKB_TN_LO
#define SKB_TN_LO 128
#endif
#ifndef SKB_TN_MID
#define SKB_TN_MID 64
#endif
#ifndef SKB_TN_HI
#define SKB_TN_HI 64
#endif
// training points per tile (= kd leaf size).  Smaller leaves have tighter boxes and cull more, which pays
// in higher dimensions (measured d=8 +6 %, d=16 +14 % for 64 vs 128); below d=5 128 is as good or better.
#ifndef SKB_TN_34
#define SKB_TN_34 128
#endif
__host__ __device__ constexpr int tile_points(int D) { return D <= 2 ? SKB_TN_LO : (D <= 4 ? SKB_TN_34 : (D <= 8 ? SKB_TN_MID : SKB_TN_HI)); }
#ifndef SKB_NSTAGE
#define SKB_NSTAGE 3
#endif
#ifndef SKB_NSTAGE_HI
#define SKB_NSTAGE_HI 2
#endif
// TMA ring depth (shared memory per CTA = depth x tile bytes) of the lanes = queries kernel.
__host__ __device__ constexpr int nstage(int D) { return D <= 4 ? SKB_NSTAGE : SKB_NSTAGE_HI; }
#ifndef SKB_NT
#define SKB_NT 32
#endif
constexpr int NT = SKB_NT;           // threads per CTA of the score kernel
constexpr int FB = 32;               // queries per finalize block (arrival-counter granularity)
#ifndef SKB_SG
#define SKB_SG 16
#endif
constexpr int SG = SKB_SG;               // tiles per super-box (consecutive kd leaves = one kd subtree)
#ifndef SKB_CG
#define SKB_CG 16
#endif
constexpr int CG = SKB_CG;               // super-boxes per chunk box (third level of the box hierarchy)
constexpr float PAD_COORD = 1.0e18f; // coordinate of padding / zero-weight points: exp2(-1e36) == 0

#ifndef SKB_Q_LE2
#define SKB_Q_LE2 2
#endif
#ifndef SKB_Q_LE4
#define SKB_Q_LE4 2
#endif
#ifndef SKB_Q_LE8
#define SKB_Q_LE8 1
#endif
#ifndef SKB_Q_LE16
#define SKB_Q_LE16 1
#endif
// resident CTAs per SM the register allocator is asked to allow (measured on B200, tools/perf_scan.py:
// d<=3 and d>=9 like 4 CTAs/SM with 128 registers, d=4 likes 2 CTAs with 232, d=5..8 like 3 with 168)
#ifndef SKB_MB_HI
#define SKB_MB_HI 4
#endif
__host__ __device__ constexpr int min_blocks_per_sm(int D) { return D <= 8 ? 4 : SKB_MB_HI; }   // x (128 / NT) CTAs
__host__ __device__ constexpr int queries_per_thread(int D)
{
    return D <= 2 ? SKB_Q_LE2 : (D <= 4 ? SKB_Q_LE4 : (D <= 8 ? SKB_Q_LE8 : SKB_Q_LE16));
}
// Second instantiation, tuned for the all-pairs evaluation (culling off, SKB_NO_CULL=1): more queries per
// thread = more independent FMA chains and fewer LDS per pair.  A query's arithmetic does not depend on Q
// (each query is a private register chain), so both variants give bit-identical results.
#ifndef SKB_DENSE_QMUL
#define SKB_DENSE_QMUL 1
#endif
__host__ __device__ constexpr int queries_per_thread_dense(int D) { return (D <= 4 ? 8 : (D <= 8 ? 4 : 2)) * SKB_DENSE_QMUL; }
__host__ __device__ constexpr int stage_floats(int D) { return (D + 2) * tile_points(D); }   // D coords, weights, header (box)

// D == 2 far-field record in a tile's header row: floats [FF2_OFF, FF2_OFF + 4) = centre x, y, half widths rx, ry (rx < 0:
// no record), then the (FF2_P + 1)(FF2_P + 2) / 2 coefficients B_ij, i + j <= FF2_P, row i = B_i0 .. B_i,P-i
constexpr int FF2_P = 12;
constexpr int FF2_OFF = 8;
constexpr int FF2_NCOEF = (FF2_P + 1) * (FF2_P + 2) / 2;
constexpr int FF2_CELL = 96;                                  // floats per cell record in KdeDev::ff2c: cx, cy, rx, ry, B[91], pad
constexpr float FF2_CELLS_MAX_R = 2.0f;                        // measured: never slower from 5e4 rows up (r ~ 1.1); beyond this even a cell is wider than the kernel
constexpr float FF2_SR_MAX = 1.7f / 1.3862943611198906f;      // z = 2 ln2 (|sx| rx + |sy| ry) <= 1.7: z^13 / 13! e^2z < 5e-6
static_assert(4 + FF2_NCOEF <= FF2_CELL && (4 * FF2_CELL) % 128 == 0, "cell record layout");
static_assert(FF2_OFF + 4 + FF2_NCOEF <= SKB_TN_LO, "the far-field record must fit the header row of a 2-D tile");

// Device-resident description of one fitted KDE (one per handle, uploaded at skb_create).
struct KdeDev {
    const float* tiles;      // [ntiles][D+2][tile_points(D)]
    const double* zt64;      // [N][D] whitened fp64 rows (draw centres)
    const double* cumsum_w;  // [N] sequential fp64 cumsum (NULL for unweighted)
    int ff2_cells_on;        // D == 2: 1 when the mean leaf radius (rx + ry, prescaled units) is at most FF2_CELLS_MAX_R
    const float* ff2c;       // D == 2 only: [ntiles][4][FF2_CELL] far-field records of a tile's four 32-point cells
    const float* mom;        // D == 1 only: [ntiles][12] far-field record of a tile: lo, hi, centre, radius, B0..B5, 0, 0
    const int* cs_index;     // [cs_buckets + 1] cs_index[b] = searchsorted(cumsum_w, b / cs_buckets * cumsum_w[N-1]): brackets the draw's search
    int cs_buckets;
    const float* boxes;      // [ntiles][2][D] lo / hi of the packed fp32 coordinates
    const float* sboxes;     // [nsuper][2][D] union boxes of SG consecutive tiles
    const float* cboxes;     // [nchunk][2][D] union boxes of CG consecutive super-boxes
    long long N;
    int ntiles;
    int nsuper;
    int nchunk;
    int D;
    int has_w;
    int has_winv;
    double c;                // prescale sqrt(log2(e)/2)/h : exp(-d^2/2h^2) == exp2(-|c dz|^2)
    double h;
    double sum_w;
    double log_norm;         // log_knorm - log(sum_w) + log(w_max)   (natural log)
    float ref_bias;          // the exponent reference sits this many binades below the closest weighted term (90 up to 4e6 rows)
    float glo[SKB_MAXD], ghi[SKB_MAXD];   // bounding box of all packed points (Morton quantisation of queries)
    double mu[SKB_MAXD];
    double Wc[SKB_MAXD * SKB_MAXD];    // forward map prescaled by c:  c*z_k = sum_i (x_i-mu_i) Wc[i*D+k]
    double Winv[SKB_MAXD * SKB_MAXD];  // inverse map: x_i = sum_k z_k Winv[k*D+i] + mu_i
};

// One (KDE, query table) job of a score launch; blockIdx.y indexes these.
struct ScoreJob {
    const KdeDev* kde;
    const void* q;
    long long M;
    long long ldq;
    int q_f32;
    int q_is_whitened;
    int cols_identity;       // cols == 0..D-1 and rows are dense & 16B aligned -> vector loads
    int nsplit;              // training-set splits (blockIdx.z)
    int fin_first, fin_count;  // finalize group = jobs [fin_first, fin_first+fin_count) of FinTable
    int* counters;           // [ceil(M/FB)] arrival counters of the group (NULL -> finalize from registers)
    int expected;            // arrivals per finalize block
    int fin_slot;            // this job's index in the FinTable
    int cull;                // 1: skip tiles whose box proves every term flushes to zero
    const int* perm;         // visit order -> row of the query table (Morton order), or NULL
    unsigned long long* stats;   // 4 diagnostic counters, or NULL (production: nothing is counted)
    int ff2_cells;           // D == 2 cell records: -1 per the fitted KDE, 0 / 1 forced off / on (skb_set_option("ff2_cells"))
    signed char cols[SKB_MAXD];
};

struct FinJob {
    double2* partial;        // [nsplit][M]  (m, S):  sum_j w_j exp2(-acc_j) = S * exp2(-m)
    double* out_logp;        // [M] natural log density, or NULL
    double* out_m;           // partial-mode outputs (natural-log m, s) or NULL
    double* out_s;
    const int* perm;         // visit order -> row (outputs and u are indexed by row), or NULL
    double log_norm;
    double log_wmax;         // partial-mode outputs carry the true weights: m += log(w_max)
    double sign_logjac;      // sign * log|jac|
    int nsplit;
    int sign;
};

struct FinTable {
    FinJob fj[SKB_MAX_JOBS];
    // ratio / accept part (group mode)
    const double* u;         // [M] uniforms or NULL
    double log_m;
    double band;
    double* out_log_ratio;
    unsigned char* accept;
    unsigned char* in_band;
    int ratio;               // 1 -> form the ratio over the group
};

struct ScoreLaunch {
    ScoreJob jobs[SKB_MAX_JOBS];
    FinTable fin;
};

// ---- the culled score path (skb_sparse.cu): plan -> bucket -> evaluate -> finalize ----
struct SparseJob {
    const float* tiles;          // the KDE's tile array
    float* qz;                   // [M][qs] prescaled whitened query rows: z[0..D), -m, padding
    unsigned long long* acc;     // [M][4] 256-bit fixed-point sums (limbs low to high), LSB 2^-128 relative to 2^m
    float* newbest;              // [M] 0x7f7f7f7f, or the exact smallest weighted exponent of a query whose reference overflowed
    unsigned* wflag;             // [ceil(M / 32)] the same flags, one bit per query of a query warp
    long long M;
    float ref_bias;
    int tile_base;               // first bucket of this job within its dimension group
    int qs;
    int D;
};

struct SparseLaunch {
    SparseJob sj[SKB_MAX_JOBS];          // indexed by the job's slot in the finalize table
    int jobs_of_group[SKB_MAX_JOBS];     // slots of this dimension group's jobs, by increasing tile_base
    int njobs;                           // jobs in this dimension group
    unsigned* ent_tile;                  // [cap] entries as emitted: bucket (tile_base + tile), query warp, mask of its queries
    unsigned* ent_warp;
    unsigned* ent_mask;
    unsigned* srt_tile;                  // [cap] the same, sorted by bucket
    unsigned* srt_warp;
    unsigned* srt_mask;
    unsigned long long* counter;         // entries reserved so far
    unsigned* overflow;                  // set if the entry list ever overflowed (results are then NaN)
    unsigned* nflag;                     // queries flagged for re-referencing in this sub-chunk
    unsigned* hist;                      // [buckets of a pass][NB_CTAS] (+1): histogram, scanned in place
    unsigned long long* stats;           // diagnostic counters or NULL
    unsigned long long cap;
    long long vbeg, vend;                // visit positions planned by this launch
};

// kernels / launchers (defined in the .cu files)
cudaError_t launch_fit(const void* train, int f32, long long N, int D, long long ld, int is_whitened,
                       const KdeDev* kde_dev, double* zt64, cudaStream_t st);
cudaError_t launch_pack(const double* zt64, const int* perm, int D, double c, const double* w, double inv_wmax,
                        float* tiles, int ntiles, float* boxes, float* sboxes, float* cboxes, cudaStream_t st);
size_t sort_temp_bytes(long long M);
cudaError_t launch_query_order(const KdeDev* kde, const void* q, int q_f32, long long M, long long ldq,
                               int q_is_whitened, const signed char* cols_dev, unsigned long long* keys_in,
                               unsigned long long* keys_out, int* idx_in, int* idx_out, void* temp, size_t temp_bytes,
                               cudaStream_t st);
cudaError_t launch_kd_order(const double* zt64, const double* w, long long N, long long nv, int D, double c,
                            int* perm, KdeDev* kde_dev, cudaStream_t st);
cudaError_t launch_features(const skb_feature_spec* specs, int n_feat, long long n_rows, double* out, unsigned char* keep,
                            cudaStream_t st);
size_t sparse_scan_temp_bytes();
cudaError_t sparse_prepare();
cudaError_t launch_sparse_group(int D, const ScoreLaunch& L, int nl, const SparseLaunch& SLh, const SparseLaunch* SLd,
                                long long max_rows, int nbuckets, void* scan_tmp, size_t scan_tmp_bytes, int sms, cudaStream_t st);
cudaError_t launch_sparse_finalize(const ScoreLaunch& L, const SparseLaunch* SLd, int first, int count, int njobs_y,
                                   long long max_rows, cudaStream_t st);
cudaError_t launch_score(int D, bool dense_q, const ScoreLaunch& L, int njobs, long long max_rows, int max_split,
                         int max_ntiles, cudaStream_t st);
int score_ctas_per_sm(int D, bool dense_q);
cudaError_t launch_draw(const KdeDev* kde, int D, long long M, const double* u, const double* z,
                        uint64_t seed, uint64_t offset, const long long* rows,
                        int rcol, double rmin, double rmax, int max_attempts,
                        long long* out_idx, double* out_x, long long ld_out, double* out_z,
                        unsigned char* out_flag, unsigned long long* n_flagged, cudaStream_t st);
cudaError_t launch_moments2(float* tiles, const float* boxes, int ntiles, float* cells, KdeDev* kde_dev, cudaStream_t st);
cudaError_t launch_moments(const float* tiles, const float* boxes, int ntiles, float* mom, cudaStream_t st);
cudaError_t launch_cs_index(const double* cumsum, long long N, int* index, int buckets, cudaStream_t st);
cudaError_t launch_compact(const unsigned char* flags, long long M, const void* rows, int elem_bytes,
                           long long ld, void* out, long long* out_index,
                           unsigned long long* block_counts, unsigned long long* total,
                           cudaStream_t st);
long long compact_nblocks(long long M);
long long compact_ws_words(long long M);
cudaError_t measure_pipes(double ms_each, double* fma_lane_per_s, double* ex2_per_s);

void count_launch(int n = 1);

}  // namespace skb

/*
 * skb.h — C ABI of the B200 (sm_100a) Gaussian-KDE score / draw / accept backend.
 *
 * This is the drop-in boundary for ONE path of vargatn/skysampler: everything that
 * `skysampler/emulator.py`'s KDEContainer delegates to scikit-learn's KernelDensity
 * (emulator.py:351-385), plus the accept test of read_concentric (emulator.py:737-746).
 * The reference is pure Python; the binding a maintainer adds is a ctypes stub
 * (see INTEGRATION.md, and skysampler_b200/_cabi.py for the one this repo ships).
 *
 * Conventions
 *   - every entry point returns 0 on success or a negative SKB_E* code; the message is in
 *     skb_last_error() (thread-local).  The Python shim re-raises ValueError / RuntimeError
 *     where sklearn would (dimension mismatch: _binary_tree.pxi.tp:1474-1476; negative
 *     weights: _kde.py:236-238).
 *   - data pointers may be HOST or DEVICE pointers (detected with cudaPointerGetAttributes);
 *     host buffers are staged by the library, device buffers are used in place (zero copy).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls
 *     with only device buffers are asynchronous and stream-ordered; calls touching host
 *     buffers return after the results are in the host buffer.
 *   - dtype codes: SKB_F64 / SKB_F32 for row-major [rows][ld] matrices.
 *   - no torch types, no C++ types; plain pointers and sizes only.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     SKB_ECUDA.
 *   - every (query, training) pair is accounted for exactly as the all-pairs sum in fp32 terms /
 *     fp64 log-sum-exp would: tiles of training points whose bounding box proves that all their terms
 *     flush to zero are skipped (kd-ordered tiles, Morton-ordered queries).  SKB_NO_CULL=1 in the
 *     environment evaluates every pair (diagnostics; it runs a different kernel, so its results agree
 *     with the default path to fp32 summation-order rounding — a few 1e-5 in log p — not in bits).  A
 *     query's result never depends on the other rows of the call: any batching / sharding of the query
 *     table is bit-identical.
 *   - handles own no scratch: per-call buffers come from the CUDA stream-ordered memory pool on the call's
 *     stream, so calls on different streams may share a handle.
 */
#ifndef SKB_H
#define SKB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKB_MAXD 16          /* feature dimensions 1..16 (reference uses 1..8) */
#define SKB_MAX_GROUP 8      /* KDEs combined in one ratio/accept test */
#define SKB_MAX_JOBS 16      /* (KDE, query-table) jobs batched in one launch */

#define SKB_F64 0
#define SKB_F32 1

#define SKB_OK        0
#define SKB_EINVAL   -1      /* bad argument (-> ValueError) */
#define SKB_ECUDA    -2      /* CUDA runtime error / no device (-> RuntimeError) */
#define SKB_ENOMEM   -3
#define SKB_ESTATE   -4      /* handle not usable for this call (e.g. draw without inverse map) */

typedef struct skb_kde* skb_handle;

/* Library / device info. */
const char* skb_last_error(void);
int  skb_version(void);
int  skb_device_count(void);                       /* 0 when no CUDA device is usable */

/*
 * Fit = upload.  Replaces KernelDensity(...).fit(_data, sample_weight=weights)
 * (emulator.py:351-356 -> sklearn _kde.py:198-250; no tree is built: every pair is evaluated).
 *
 *   train        [N][ld_train] rows, dtype train_dtype.  If train_is_whitened != 0 these are the
 *                whitened rows (`KDEContainer._data`); otherwise raw rows that the library whitens
 *                with (mu, W) in fp64.
 *   w            N sample weights (>= 0) or NULL for an unweighted KDE.
 *   cumsum_w     N sequential fp64 partial sums of w as numpy.cumsum gives them (sklearn
 *                _kde.py:345), or NULL -> the library forms the same sequential sum on the host.
 *                Only used by skb_draw.
 *   h            bandwidth (> 0).
 *   mu, W        forward whitening  z = (x - mu) . W,  W is [d][d] row-major; this is
 *                pca_transform (emulator.py:322-327) with mu = mean1 + pca.mean_ and
 *                W = components_.T . diag(1/std2).  NULL,NULL -> identity (queries already whitened).
 *   Winv         inverse map  x = z . Winv + mu  (pca_inverse_transform, emulator.py:329-335,
 *                Winv = diag(std2) . components_).  NULL -> draws are returned in whitened space.
 *   device       CUDA device ordinal.
 */
int skb_create(const void* train, int train_dtype, int64_t N, int d, int64_t ld_train,
               int train_is_whitened,
               const double* w, const double* cumsum_w, double h,
               const double* mu, const double* W, const double* Winv,
               int device, skb_handle* out);

int skb_destroy(skb_handle h);                      /* drop_kde (emulator.py:387-389) */

int64_t skb_n_train(skb_handle h);
int     skb_ndim(skb_handle h);
double  skb_sum_weight(skb_handle h);
int     skb_device(skb_handle h);
double  skb_fit_ms(skb_handle h);                   /* device time of the fit inside skb_create (whitening, kd ordering, packing) */

/*
 * score_samples (emulator.py:380-385 -> sklearn _kde.py:252-288 -> kernel_density
 * _binary_tree.pxi.tp:1393-1535): out_logp[i] = log p(q_i), natural log, fp64.
 *
 *   q               [M][ldq] rows.  Column cols[k] of a row feeds feature k (cols == NULL ->
 *                   0..d-1): the reference selects column subsets of one proposal table per KDE
 *                   (emulator.py:671,679,687).
 *   q_is_whitened   rows are already in the whitened space (skip (mu, W)).
 */
int skb_score(skb_handle h, const void* q, int q_dtype, int64_t M, int64_t ldq,
              const int32_t* cols, int q_is_whitened,
              double* out_logp, void* stream);

/*
 * Training-sharded mode: per-query partial log-sum-exp state over THIS handle's rows only,
 *     sum_j w_j exp(-|z - z_j|^2 / 2h^2) = out_s[i] * exp(out_m[i])        (natural log units)
 * Partials from shards combine as  m = max m_g,  s = sum_g s_g exp(m_g - m);  the normalisation
 * (log_knorm - log sum_w over ALL shards) is added by the caller (skb_log_knorm helps).
 */
int skb_score_partial(skb_handle h, const void* q, int q_dtype, int64_t M, int64_t ldq,
                      const int32_t* cols, int q_is_whitened,
                      double* out_m, double* out_s, void* stream);

double skb_log_knorm(int d, double h);             /* -d/2 log(2 pi) - d log h (_binary_tree.pxi.tp:448-478) */

/*
 * Batched scoring: n_jobs independent (KDE, query table) jobs in ONE launch per feature
 * dimension — the concentric-shell sampler's "10 shell KDEs as a batch"
 * (bin/epsilon_concentric_sampler_v05.py:127-181 loops them serially).
 */
typedef struct skb_job {
    skb_handle     kde;
    const void*    q;
    int32_t        q_dtype;
    int32_t        q_is_whitened;
    int64_t        M;
    int64_t        ldq;
    const int32_t* cols;           /* HOST pointer, d entries, or NULL */
    double*        out_logp;       /* [M] */
} skb_job;

int skb_score_batch(const skb_job* jobs, int n_jobs, void* stream);

/*
 * Fused ratio + accept (emulator.py:737-746): all KDEs of a group score the SAME proposal
 * table in one pass and the last-arriving thread block of each query block forms
 *     log_ratio = sum_k sign[k] * (log p_k + log_abs_jac[k]) - log_m
 *     accept[i] = log(u[i]) < log_ratio          (u < p_ref / (m * prod p_proposal))
 *     in_band[i] = |log_ratio - log(u[i])| <= band     (knife-edge flag for parity checks)
 *   The test runs in LOG space; the reference compares in linear space (exp of each log-density), where a log-density
 *   below -745 underflows to 0 and turns the ratio into 0/0 = NaN (reject) or x/0 = inf (accept).  Rows that deep in the
 *   tails can therefore decide differently; they are not flagged by in_band (out_logp shows them).
 *   sign[k] = +1 for reference-density terms (wcr), -1 for proposal terms (dc, wr).
 *   cols    [n_kde][SKB_MAXD] HOST int32 column indices into the proposal table (-1 padded), or NULL.
 *   u       [M] uniforms (fp64) or NULL -> no accept test, only log p's.
 *   out_logp  [n_kde][M] fp64 or NULL.   accept / in_band  [M] uint8 or NULL.
 *   out_log_ratio [M] fp64 or NULL.
 */
int skb_ratio_accept(const skb_handle* kdes, int n_kde, const int32_t* sign,
                     const double* log_abs_jac, double log_m,
                     const void* q, int q_dtype, int64_t M, int64_t ldq,
                     const int32_t* cols,
                     const double* u, double band,
                     double* out_logp, double* out_log_ratio,
                     uint8_t* accept, uint8_t* in_band, void* stream);

/*
 * KDE draw: KernelDensity.sample (sklearn _kde.py:312-349) + pca_inverse_transform
 * (emulator.py:329-335), optionally with the radial window of random_draw (emulator.py:362-375).
 *
 *   u, z       host- or device-supplied streams (parity mode): u[M] uniforms in [0,1),
 *              z[M][d] standard normals, consumed exactly like sklearn does
 *              (i = searchsorted(cumsum_w, u*sum_w) or int64(u*N); x = data[i] + h*z).
 *              Both NULL -> Philox4x32-10 counter RNG keyed by (seed, draw index + offset, attempt).
 *   rows       NULL, or M int64 destination row indices (draw i overwrites row rows[i] of the
 *              outputs) — the redraw step `_res[inds, :] = vals` (emulator.py:373).
 *   rcol/rmin/rmax   radial window test on RAW column rcol; rcol < 0 disables it.  In Philox mode
 *              out-of-window draws are redrawn inside the kernel (attempt counter) until they pass
 *              (at most max_attempts; rows still failing are flagged).  In stream mode the kernel
 *              draws once and only flags.
 *   out_idx    [*] int64 selected centre index (nullable).
 *   out_x      [*][ld_out] fp64 raw-space samples (whitened-space if the handle has no Winv).
 *   out_z      [*][d] fp64 whitened-space samples (nullable) — the `_res` array of random_draw.
 *   out_flag   [*] uint8 1 = outside the window after this call (nullable).
 *   n_flagged  HOST int64: number of flagged rows (nullable; forces a sync when given).
 */
int skb_draw(skb_handle h, int64_t M, const double* u, const double* z,
             uint64_t seed, uint64_t offset,
             const int64_t* rows,
             int rcol, double rmin, double rmax, int max_attempts,
             int64_t* out_idx, double* out_x, int64_t ld_out, double* out_z,
             uint8_t* out_flag, int64_t* n_flagged, void* stream);

/*
 * Order-preserving stream compaction of accepted rows (`samples[inds]`, emulator.py:746-747).
 *   flags [M] uint8, rows [M][ld] (dtype), out [count][ld]; out_index [count] int64 source row
 *   ids (nullable); n_out receives the count: a HOST int64 (forces a sync) or a DEVICE int64 (the call stays asynchronous).
 */
int skb_compact(const uint8_t* flags, int64_t M, const void* rows, int dtype, int64_t ld,
                void* out, int64_t* out_index, int64_t* n_out, void* stream);

/*
 * Feature construction in front of the KDE fit (BaseContainer.construct_features, emulator.py:74-132): every feature is
 * a catalogue column, a literal, or a binary expression of two of them; rows are kept when every feature lies inside its
 * OPEN interval (lo, hi); features flagged log10 are replaced by their decimal logarithm.  One kernel, one thread per
 * catalogue row; operands are DEVICE columns (element stride in elements, so a component of a vector column is just an
 * offset pointer with the vector length as stride).  Arithmetic follows numpy: float32 when every array operand is
 * float32 (literals are weak scalars), else float64; the stored feature is float64.
 *   op: 0 = a, 1 = a - b, 2 = a + b, 3 = a * b, 4 = a / b, 5 = sqrt(a^2 + b^2)
 *   out  [n_rows][n_feat] fp64 (all rows; rows with keep == 0 are to be dropped, e.g. with skb_compact), keep [n_rows].
 */
typedef struct skb_feature_spec {
    const void* a;            /* device column or NULL (-> a_const) */
    const void* b;            /* device column or NULL (-> b_const; ignored for op 0) */
    int64_t     a_stride, b_stride;
    double      a_const, b_const;
    double      lo, hi;       /* open interval; -inf / +inf disable a side */
    int32_t     a_dtype, b_dtype;   /* SKB_F64 / SKB_F32 */
    int32_t     op;
    int32_t     log10;
} skb_feature_spec;

int skb_features(const skb_feature_spec* specs /* HOST */, int n_feat, int64_t n_rows,
                 double* out, uint8_t* keep, void* stream);

/*
 * FP32-FMA / MUFU.EX2 pipe peaks measured live on the device (roofline denominators, H8 of
 * SURVEY.md): lane-FMAs per second through FFMA2, EX2 per second, over ~ms_each ms each.
 */
int skb_measure_pipes(int device, double ms_each, double* fma_lane_per_s, double* ex2_per_s);

/*
 * Process-wide diagnostic switches (the release kernels carry no counters and read no environment per call):
 *   "cull"          1 (default) skip training tiles whose box proves every term flushes to zero; 0 = evaluate
 *                   every pair with the all-pairs kernel (roofline measurements; agrees to fp32 rounding).
 *                   The initial value is 0 when SKB_NO_CULL=1 is in the environment at load time.
 *   "force_nsplit"  > 0: split the training tiles over that many CTAs per query block (experiments).
 *   "stats"         1: launches made while it is set count evaluated pairs / culled tiles (skb_debug_counters).
 *   "kernel"        0 (default): route by dimension (d = 1 sorted intervals + far-field records, d = 2 lanes = queries with
 *                   far-field records, d = 3..5 sorted-incidence pipeline, d >= 6 per-query kernel); 1 / 2 force the
 *                   single-launch kernels / the pipeline (A/B runs).
 *   "ff2_cells"     -1 (default): the 2-D second-level far-field records are used per the fitted KDE; 0 / 1 force them off / on.
 *   "host_stage_rows" > 0: rows per staged chunk of a host-buffer call (default 0: one chunk up to 2^21 rows — measured
 *                   faster than overlapping two half uploads — and chunks of 2^19 beyond).
 */
int skb_set_option(const char* name, int value);
int skb_get_option(const char* name);               /* -1 for an unknown name */

/*
 * Diagnostics of the score kernel since the last reset (synchronises the device); all zero unless
 * skb_set_option("stats", 1) was in effect for the launches of interest:
 *   out4[0] = (query, point) pairs per lane re-run after an exponent-window overflow (warp level)
 *   out4[1] = (query, point) pairs per lane evaluated (warp level): x 32 lanes = pairs evaluated
 *   out4[2] = tiles skipped by the box tests (warp level);  out4[3] = tiles in range (warp level).
 */
int skb_debug_counters(int device, uint64_t* out4, int reset);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t skb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SKB_H */

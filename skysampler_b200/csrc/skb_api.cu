// C-ABI glue of libskb (include/skb.h): handle lifetime, host<->device staging, launch planning.
// No compute happens here and there is no CPU fallback: every path ends in a kernel launch or an
// error code.
#include "skb_internal.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <vector>

namespace skb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define SKB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            cudaGetLastError();                                                                 \
            return fail(_e == cudaErrorMemoryAllocation ? SKB_ENOMEM : SKB_ECUDA, "%s: %s", #expr, \
                        cudaGetErrorString(_e));                                                \
        }                                                                                       \
    } while (0)

// true when p can be dereferenced by a kernel on the current device
static bool is_device_ptr(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Per-call device scratch from the CUDA stream-ordered memory pool: allocated and freed on the call's
// stream, so a handle owns no scratch, concurrent calls on different streams never share a buffer, and the
// pool (release threshold raised at first use) hands the same blocks back call after call.
struct Arena {
    cudaStream_t st;
    void* blocks[128];
    int n = 0;
    explicit Arena(cudaStream_t s) : st(s) {}
    cudaError_t get(void** out, size_t bytes)
    {
        *out = nullptr;
        if (n >= 128) return cudaErrorMemoryAllocation;
        cudaError_t e = cudaMallocAsync(out, bytes ? bytes : 16, st);
        if (e == cudaSuccess) blocks[n++] = *out;
        return e;
    }
    ~Arena() { for (int i = 0; i < n; i++) cudaFreeAsync(blocks[i], st); }
};

static void raise_pool_threshold(int device)
{
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;               // keep freed blocks cached instead of returning them to the OS
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    done[device] = true;
}

// Process-wide diagnostic switches (skb_set_option).  Defaults come from the environment ONCE, at load.
static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
static std::atomic<int> g_opt_cull{env_int("SKB_NO_CULL", 0) ? 0 : 1};
static std::atomic<int> g_opt_nsplit{env_int("SKB_FORCE_NSPLIT", 0)};
static std::atomic<int> g_opt_stats{0};
static std::atomic<int> g_opt_ff2_cells{-1};             // diagnostic: D == 2 cell records -1 auto / 0 off / 1 on
static std::atomic<int> g_opt_kernel{env_int("SKB_KERNEL", 0)};     // 0: route by dimension; 1: single-launch kernels; 2: sorted-incidence pipeline (A/B runs)
static unsigned long long* g_stats_dev[64] = {nullptr};

}  // namespace skb

using namespace skb;

struct skb_kde {
    int device = 0;
    int sms = 148;
    KdeDev hk;                 // host copy of the device descriptor
    KdeDev* dk = nullptr;      // device descriptor
    void* slab = nullptr;      // the one device allocation everything below lives in
    float* tiles = nullptr;
    float* boxes = nullptr;
    float* sboxes = nullptr;
    double* zt64 = nullptr;
    double* cumsum = nullptr;
    double log_wmax = 0.0;
    float fit_ms = 0.f;        // device time of the fit (whitening, kd ordering, packing)
};

namespace skb {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- launch planning --------------------------------------------------------------------------

// Training-set splits of a launch (blockIdx.z).  Never a function of the number of query rows: the
// arithmetic a query sees must be the same whatever batch, chunk or GPU it lands in, so that results
// are bit-identical under any query sharding.
static bool cull_enabled() { return g_opt_cull.load(std::memory_order_relaxed) != 0; }

static int plan_nsplit(const skb_kde* h, long long /*M*/, int /*njobs*/)
{
    const int v = g_opt_nsplit.load(std::memory_order_relaxed);      // skb_set_option("force_nsplit"): experiments only
    if (v >= 1) return v > h->hk.ntiles ? h->hk.ntiles : v;
    // Every CTA walks the box table of ALL tiles once to find its queries' exponent references; splitting
    // the tiles over CTAs would repeat that walk per split, and with culling a CTA only visits a fraction
    // of the tiles anyway.  One split also keeps a query's arithmetic a function of the fitted KDE alone.
    return 1;
}

struct GroupSpec {
    int ratio = 0;
    const int32_t* sign = nullptr;
    const double* log_abs_jac = nullptr;
    double log_m = 0.0;
    const double* u = nullptr;          // device
    double band = 0.0;
    double* out_log_ratio = nullptr;    // device
    unsigned char* accept = nullptr;    // device
    unsigned char* in_band = nullptr;   // device
};

struct DevJob {
    skb_kde* h;
    const void* q;        // device
    int q_f32;
    int q_is_whitened;
    long long M, ldq;
    const int32_t* cols;  // host
    double* out_logp;     // device or NULL
    double* out_m;        // device or NULL
    double* out_s;
};

// All pointers are device pointers here.  grouped: every job scores the same M rows and the
// finalize step sees all of them; otherwise each job finalizes alone.  Scratch lives in `ar` (freed on
// the stream when the caller's Arena goes out of scope, i.e. after everything enqueued here).
static int run_jobs_device(const DevJob* jobs, int n, bool grouped, const GroupSpec& g, Arena& ar, cudaStream_t st)
{
    if (n <= 0) return SKB_OK;
    if (n > SKB_MAX_JOBS) return fail(SKB_EINVAL, "too many jobs in one launch (%d > %d)", n, SKB_MAX_JOBS);
    // ---- validate every job BEFORE anything is enqueued: a bad column index must come back as SKB_EINVAL,
    //      not as an out-of-bounds read of the ordering kernel ----
    for (int j = 0; j < n; j++) {
        const DevJob& dj = jobs[j];
        const int D = dj.h->hk.D;
        if (!dj.cols && dj.ldq < D)
            return fail(SKB_EINVAL, "query has %lld features, the KDE was fitted with %d", dj.ldq, D);
        for (int k = 0; k < D; k++) {
            const int c = dj.cols ? dj.cols[k] : k;
            if (c < 0 || c >= dj.ldq || c > 127)
                return fail(SKB_EINVAL, "query column index %d out of range for row stride %lld", c, dj.ldq);
        }
        if (dj.M > 0x7fffffffLL) return fail(SKB_EINVAL, "too many query rows for one launch");
    }
    ScoreLaunch L;
    memset(&L, 0, sizeof(L));
    int nsplit[SKB_MAX_JOBS];
    int expected_group = 0;
    for (int j = 0; j < n; j++) {
        const DevJob& dj = jobs[j];
        int same = 0;
        for (int k = 0; k < n; k++) same += (jobs[k].h->hk.D == dj.h->hk.D);
        nsplit[j] = plan_nsplit(dj.h, dj.M, same);
        expected_group += nsplit[j];
    }
    const bool cull = cull_enabled();
    // which culled kernel scores a job (a function of its dimension only, so that batching never changes the arithmetic):
    // the sorted-incidence pipeline (plan -> bucket -> evaluate -> finalize) where it is the fastest (measured on B200,
    // profiles/r04f_scan_routes.txt: d = 3..5), the lanes = queries kernel for d <= 2 (dense incidence, already at the SFU
    // roofline of its pairs) and the per-query kernel for d >= 6 (since its tiles come straight from L2 it is ahead of the
    // pipeline at d = 6 and 7 too: 21.1 vs 21.6 ms, 22.0 vs 22.6 ms).  skb_set_option("kernel", 1 / 2) forces either family.
    const int kopt = g_opt_kernel.load(std::memory_order_relaxed);
    bool use_sparse[SKB_MAX_JOBS];
    bool sparse = false;
    for (int j = 0; j < n; j++) {
        const int D = jobs[j].h->hk.D;
        // (small training sets — up to two chunks of tiles — stay on the single-launch round-1 kernels: the pipeline's
        //  fixed cost of a dozen launches would dominate; like D, the tile count is a property of the fitted KDE)
        use_sparse[j] = cull && (kopt == 2 || (kopt == 0 && D >= 3 && D <= 5 && jobs[j].h->hk.ntiles > 2 * CG * SG));
        sparse |= use_sparse[j];
    }
    const bool fin_by_kernel = sparse && grouped;       // a group with a sparse member is finalized by the finalize kernel
    // finalize table
    for (int j = 0; j < n; j++) {
        const DevJob& dj = jobs[j];
        FinJob& fj = L.fin.fj[j];
        const bool need_partial = !use_sparse[j] && (grouped || nsplit[j] > 1);
        if (need_partial) {
            void* p = nullptr;
            SKB_CUDA(ar.get(&p, (size_t)nsplit[j] * (size_t)dj.M * sizeof(double2)));
            fj.partial = reinterpret_cast<double2*>(p);
        }
        fj.out_logp = dj.out_logp;
        fj.out_m = dj.out_m;
        fj.out_s = dj.out_s;
        fj.log_norm = dj.h->hk.log_norm;
        fj.log_wmax = dj.h->log_wmax;
        fj.nsplit = nsplit[j];
        fj.sign = (grouped && g.sign) ? g.sign[j] : 0;
        fj.sign_logjac = (grouped && g.sign && g.log_abs_jac) ? g.sign[j] * g.log_abs_jac[j] : 0.0;
    }
    L.fin.ratio = grouped ? g.ratio : 0;
    L.fin.u = g.u;
    L.fin.log_m = g.log_m;
    L.fin.band = g.band;
    L.fin.out_log_ratio = g.out_log_ratio;
    L.fin.accept = g.accept;
    L.fin.in_band = g.in_band;

    // ---- visit order of the query rows: Morton order in the prescaled space of the widest KDE ----
    const int* perm_of[SKB_MAX_JOBS];
    for (int j = 0; j < n; j++) perm_of[j] = nullptr;
    if (cull) {
        int lead = 0;
        for (int j = 1; j < n; j++) if (jobs[j].h->hk.D > jobs[lead].h->hk.D) lead = j;
        for (int j = 0; j < n; j++) {
            if (grouped && j != lead) continue;
            const DevJob& dj = jobs[j];
            if (dj.M < 1024) continue;
            const size_t M_ = (size_t)dj.M;
            const size_t tb = sort_temp_bytes(dj.M);
            const size_t need = M_ * (8 + 8 + 4 + 4) + 64 + tb + 256;
            void* blk = nullptr;
            SKB_CUDA(ar.get(&blk, need));
            char* base = reinterpret_cast<char*>(blk);
            unsigned long long* keys_in = reinterpret_cast<unsigned long long*>(base);
            unsigned long long* keys_out = keys_in + M_;
            int* idx_in = reinterpret_cast<int*>(keys_out + M_);
            int* idx_out = idx_in + M_;
            signed char* cols_dev = reinterpret_cast<signed char*>(idx_out + M_);
            void* temp = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(cols_dev) + 64 + 255) & ~(uintptr_t)255);
            signed char ch[SKB_MAXD];
            for (int k = 0; k < SKB_MAXD; k++) ch[k] = (signed char)(k < dj.h->hk.D ? (dj.cols ? dj.cols[k] : k) : 0);
            SKB_CUDA(cudaMemcpyAsync(cols_dev, ch, SKB_MAXD, cudaMemcpyHostToDevice, st));
            SKB_CUDA(launch_query_order(dj.h->dk, dj.q, dj.q_f32, dj.M, dj.ldq, dj.q_is_whitened, cols_dev, keys_in,
                                        keys_out, idx_in, idx_out, temp, tb, st));
            perm_of[j] = idx_out;
        }
        if (grouped) for (int j = 0; j < n; j++) perm_of[j] = perm_of[lead];
    }
    for (int j = 0; j < n; j++) L.fin.fj[j].perm = perm_of[j];

    // arrival counters: zeroed on the call's own stream, every call (the kernels' self-reset is not relied on)
    auto counters_for = [&](long long M, int** out) -> cudaError_t {
        const size_t nfb = (size_t)((M + FB - 1) / FB);
        void* p = nullptr;
        cudaError_t e = ar.get(&p, nfb * sizeof(int));
        if (e != cudaSuccess) return e;
        *out = reinterpret_cast<int*>(p);
        return cudaMemsetAsync(p, 0, nfb * sizeof(int), st);
    };
    int* group_counters = nullptr;
    if (grouped && !(sparse && [&] { for (int j = 0; j < n; j++) if (!use_sparse[j]) return false; return true; }()))
        SKB_CUDA(counters_for(jobs[0].M, &group_counters));
    unsigned long long* stats = nullptr;
    if (g_opt_stats.load(std::memory_order_relaxed)) {
        const int dev = jobs[0].h->device;
        if (dev >= 0 && dev < 64) {
            if (!g_stats_dev[dev]) {
                SKB_CUDA(cudaMalloc(reinterpret_cast<void**>(&g_stats_dev[dev]), 4 * sizeof(unsigned long long)));
                SKB_CUDA(cudaMemset(g_stats_dev[dev], 0, 4 * sizeof(unsigned long long)));
            }
            stats = g_stats_dev[dev];
        }
    }
    // ---- culled path: per-job query rows + accumulators, one entry list per dimension group ----
    SparseLaunch SLh;
    SparseLaunch* SLd_all = nullptr;              // device copies, one per (dimension group, sub-chunk)
    int n_sl = 0, sl_cap = 0;
    if (sparse) {
        SKB_CUDA(sparse_prepare());
        memset(&SLh, 0, sizeof(SLh));
        for (int j = 0; j < n; j++) {
            const DevJob& dj = jobs[j];
            const int D = dj.h->hk.D;
            const int qs = (D + 4) & ~3;
            SLh.sj[j].M = dj.M;
            SLh.sj[j].D = D;
            if (!use_sparse[j]) continue;                // scored by a round-1 kernel: (m, S) partials instead of accumulators
            void* p = nullptr;
            SKB_CUDA(ar.get(&p, (size_t)dj.M * qs * sizeof(float)));
            SLh.sj[j].qz = reinterpret_cast<float*>(p);
            SKB_CUDA(ar.get(&p, (size_t)dj.M * 32));
            SKB_CUDA(cudaMemsetAsync(p, 0, (size_t)dj.M * 32, st));
            SLh.sj[j].acc = reinterpret_cast<unsigned long long*>(p);
            SKB_CUDA(ar.get(&p, (size_t)dj.M * 4));
            SKB_CUDA(cudaMemsetAsync(p, 0x7f, (size_t)dj.M * 4, st));
            SLh.sj[j].newbest = reinterpret_cast<float*>(p);
            SKB_CUDA(ar.get(&p, (size_t)((dj.M + 31) / 32) * 4));
            SKB_CUDA(cudaMemsetAsync(p, 0, (size_t)((dj.M + 31) / 32) * 4, st));
            SLh.sj[j].wflag = reinterpret_cast<unsigned*>(p);
            SLh.sj[j].ref_bias = dj.h->hk.ref_bias;
            SLh.sj[j].tiles = dj.h->tiles;
            SLh.sj[j].M = dj.M;
            SLh.sj[j].qs = qs;
            SLh.sj[j].D = D;
        }
        void* p = nullptr;
        SKB_CUDA(ar.get(&p, 64));
        SKB_CUDA(cudaMemsetAsync(p, 0, 64, st));
        SLh.counter = reinterpret_cast<unsigned long long*>(p);
        SLh.overflow = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(p) + 16);
        SLh.nflag = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(p) + 8);
        SLh.stats = stats;
        // worst-case number of device descriptors: one per sub-chunk per dimension group (bounded below)
        sl_cap = 0;
        long long maxM = 0;
        for (int j = 0; j < n; j++) if (jobs[j].M > maxM) maxM = jobs[j].M;
        sl_cap = (int)(n * ((maxM + 32 * 128 - 1) / (32 * 128)) + n + 1);
        if (sl_cap > 4096) sl_cap = 4096;
        SKB_CUDA(ar.get(&p, (size_t)sl_cap * sizeof(SparseLaunch)));
        SLd_all = reinterpret_cast<SparseLaunch*>(p);
    }
    // one launch per distinct feature dimension
    bool done[SKB_MAX_JOBS] = {false};
    for (int j0 = 0; j0 < n; j0++) {
        if (done[j0]) continue;
        const int D = jobs[j0].h->hk.D;
        int nl = 0;
        long long max_rows = 0;
        int max_split = 1, max_ntiles = 0;
        for (int j = j0; j < n; j++) {
            if (done[j] || jobs[j].h->hk.D != D) continue;
            done[j] = true;
            const DevJob& dj = jobs[j];
            ScoreJob& sj = L.jobs[nl++];
            memset(&sj, 0, sizeof(sj));
            sj.kde = dj.h->dk;
            sj.q = dj.q;
            sj.M = dj.M;
            sj.ldq = dj.ldq;
            sj.q_f32 = dj.q_f32;
            sj.q_is_whitened = dj.q_is_whitened;
            sj.nsplit = nsplit[j];
            sj.fin_slot = j;
            sj.cull = cull ? 1 : 0;
            sj.ff2_cells = g_opt_ff2_cells.load(std::memory_order_relaxed);
            sj.perm = perm_of[j];
            sj.stats = stats;
            bool ident = true;
            for (int k = 0; k < D; k++) {
                const int c = dj.cols ? dj.cols[k] : k;
                sj.cols[k] = (signed char)c;
                ident &= (c == k);
            }
            const size_t esz = dj.q_f32 ? 4 : 8;
            sj.cols_identity = ident && ((reinterpret_cast<uintptr_t>(dj.q) & 15) == 0) && ((dj.ldq * esz) % 16 == 0);
            if (grouped) {
                sj.fin_first = 0; sj.fin_count = n;
                sj.counters = group_counters;
                sj.expected = fin_by_kernel ? 0x7fffffff : expected_group;    // nobody "arrives last": the finalize kernel does it
            } else {
                sj.fin_first = j; sj.fin_count = 1;
                if (nsplit[j] > 1) {
                    SKB_CUDA(counters_for(dj.M, &sj.counters));
                    sj.expected = nsplit[j];
                }
            }
            if (dj.M > max_rows) max_rows = dj.M;
            if (nsplit[j] > max_split) max_split = nsplit[j];
            if (dj.h->hk.ntiles > max_ntiles) max_ntiles = dj.h->hk.ntiles;
        }
        if (max_rows <= 0) continue;
        if (!use_sparse[j0]) {
            SKB_CUDA(launch_score(D, !cull, L, nl, max_rows, max_split, max_ntiles, st));
            continue;
        }
        // ---- this dimension group's entry list: worst-case-safe capacity, query rows planned in sub-chunks ----
        SLh.njobs = nl;
        int nbuckets = 0;
        double per_warp = 0.0;                       // entries one query warp can reserve over all jobs of the group
        for (int i = 0; i < nl; i++) {
            const int slot = L.jobs[i].fin_slot;
            SLh.jobs_of_group[i] = slot;
            SLh.sj[slot].tile_base = nbuckets;
            nbuckets += jobs[slot].h->hk.ntiles;
            per_warp += 1.14 * jobs[slot].h->hk.ntiles + 2.0 * 256;
        }
        const double kCapEntries = 134217728.0;      // 2^27 entries (3.2 GB of scratch with the sorted copy) per sub-chunk
        long long sub_warps = (long long)(kCapEntries / per_warp);
        if (sub_warps < 128) sub_warps = 128;
        const long long all_warps = (max_rows + 31) / 32;
        if (sub_warps > all_warps) sub_warps = all_warps;
        const unsigned long long cap = (unsigned long long)((double)(sub_warps + 1) * per_warp) + 1024;
        if (cap > 0xfff00000ull) return fail(SKB_ENOMEM, "training set too large for one entry list (%d tiles): shard it over GPUs", nbuckets);
        {
            void* p = nullptr;
            SKB_CUDA(ar.get(&p, (size_t)cap * 4 * 6));
            unsigned* base = reinterpret_cast<unsigned*>(p);
            SLh.ent_tile = base; SLh.ent_warp = base + cap; SLh.ent_mask = base + 2 * cap;
            SLh.srt_tile = base + 3 * cap; SLh.srt_warp = base + 4 * cap; SLh.srt_mask = base + 5 * cap;
            SLh.cap = cap;
            const int nbp = nbuckets < 49152 ? nbuckets : 49152;
            SKB_CUDA(ar.get(&p, ((size_t)nbp * 128 + 1) * 4));
            SLh.hist = reinterpret_cast<unsigned*>(p);
        }
        const size_t scan_bytes = sparse_scan_temp_bytes();
        void* scan_tmp = nullptr;
        SKB_CUDA(ar.get(&scan_tmp, scan_bytes));
        for (long long w0 = 0; w0 < all_warps; w0 += sub_warps) {
            if (n_sl >= sl_cap) return fail(SKB_ENOMEM, "too many planning sub-chunks");
            SLh.vbeg = w0 * 32;
            SLh.vend = (w0 + sub_warps) * 32 < max_rows ? (w0 + sub_warps) * 32 : max_rows;
            SparseLaunch* SLd = SLd_all + n_sl++;
            SKB_CUDA(cudaMemcpyAsync(SLd, &SLh, sizeof(SparseLaunch), cudaMemcpyHostToDevice, st));   // pageable source: staged before return
            SKB_CUDA(cudaMemsetAsync(SLh.counter, 0, 16, st));       // entry counter + flag counter
            SKB_CUDA(launch_sparse_group(D, L, nl, SLh, SLd, SLh.vend - SLh.vbeg, nbuckets, scan_tmp, scan_bytes,
                                         jobs[j0].h->sms, st));
        }
    }
    if (sparse) {
        if (n_sl == 0) return SKB_OK;
        long long maxM = 0;
        for (int j = 0; j < n; j++) if (jobs[j].M > maxM) maxM = jobs[j].M;
        // the last descriptor carries every job's rows / accumulators: finalize all jobs from it
        if (grouped) SKB_CUDA(launch_sparse_finalize(L, SLd_all + (n_sl - 1), 0, n, 1, jobs[0].M, st));
        else SKB_CUDA(launch_sparse_finalize(L, SLd_all + (n_sl - 1), 0, 1, n, maxM, st));
    }
    return SKB_OK;
}

// rows per chunk of a device-resident table (bounds the per-call scratch) and of a host table (copy / compute overlap)
static const long long STAGE_ROWS = 1LL << 22;
static const long long HOST_STAGE_ROWS = 1LL << 19;
static std::atomic<long long> g_opt_host_rows{0};       // diagnostic: rows per staged chunk of host-buffer calls (0: the policy below)
// Host-buffer calls: up to 2^21 rows go as ONE chunk (one ordering pass, one launch set), longer tables in chunks of 2^19 with
// the upload of chunk i+1 overlapping the kernels of chunk i.
static long long host_stage_step(long long M)
{
    const long long g = g_opt_host_rows.load(std::memory_order_relaxed);
    if (g > 0) return M > g ? g : M;
    return M > (1 << 21) ? HOST_STAGE_ROWS : M;
}

// Copy / compute overlap for calls whose buffers live in host memory: uploads and downloads run on a private
// copy stream, ordered against the caller's stream with events (two buffers in flight).
struct HostStager {
    cudaStream_t main = nullptr, copy = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, kd[2] = {nullptr, nullptr}, down[2] = {nullptr, nullptr}, start = nullptr;
    bool down_pending[2] = {false, false};
    cudaError_t init(cudaStream_t st)
    {
        main = st;
        cudaError_t e = cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
        for (int i = 0; i < 2; i++) {
            if ((e = cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&kd[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&down[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        if ((e = cudaEventCreateWithFlags(&start, cudaEventDisableTiming)) != cudaSuccess) return e;
        // the copy stream must not run ahead of work already queued on the caller's stream (buffer allocation)
        if ((e = cudaEventRecord(start, main)) != cudaSuccess) return e;
        return cudaStreamWaitEvent(copy, start, 0);
    }
    cudaError_t upload(void* dst, const void* src, size_t bytes, int b)
    {
        cudaError_t e;
        if (kd[b] && (e = cudaStreamWaitEvent(copy, kd[b], 0)) != cudaSuccess) return e;   // buffer b's last kernels are done
        if ((e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, copy)) != cudaSuccess) return e;
        return cudaEventRecord(up[b], copy);
    }
    cudaError_t mark_upload(int b) { return cudaEventRecord(up[b], copy); }
    cudaError_t wait_upload(int b) { return cudaStreamWaitEvent(main, up[b], 0); }
    cudaError_t wait_download(int b) { return down_pending[b] ? cudaStreamWaitEvent(main, down[b], 0) : cudaSuccess; }
    cudaError_t kernels_done(int b) { return cudaEventRecord(kd[b], main); }
    cudaError_t download(void* dst, const void* src, size_t bytes, int b)
    {
        cudaError_t e = cudaStreamWaitEvent(copy, kd[b], 0);                                // buffer b's kernels are done
        if (e != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, copy)) != cudaSuccess) return e;
        down_pending[b] = true;
        return cudaEventRecord(down[b], copy);
    }
    cudaError_t finish()
    {
        cudaError_t e = cudaStreamSynchronize(copy);
        if (e != cudaSuccess) return e;
        // scratch of the call is freed on `main`: order it after the copies too, and leave nothing in flight
        return cudaStreamSynchronize(main);
    }
    ~HostStager()
    {
        if (copy) { cudaStreamSynchronize(copy); cudaStreamDestroy(copy); }
        for (int i = 0; i < 2; i++) {
            if (up[i]) cudaEventDestroy(up[i]);
            if (kd[i]) cudaEventDestroy(kd[i]);
            if (down[i]) cudaEventDestroy(down[i]);
        }
        if (start) cudaEventDestroy(start);
    }
};

}  // namespace skb

// =================================================================================================
extern "C" {

const char* skb_last_error(void) { return g_err; }
int skb_version(void) { return 100; }
int64_t skb_launch_count(void) { return g_launches.load(); }

int skb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

double skb_log_knorm(int d, double h) { return -0.5 * d * std::log(2.0 * M_PI) - d * std::log(h); }

int skb_create(const void* train, int train_dtype, int64_t N, int d, int64_t ld_train, int train_is_whitened,
               const double* w, const double* cumsum_w, double h, const double* mu, const double* W,
               const double* Winv, int device, skb_handle* out)
{
    if (!out) return fail(SKB_EINVAL, "out handle is NULL");
    *out = nullptr;
    if (!train || N <= 0) return fail(SKB_EINVAL, "training set is empty");
    if (d < 1 || d > SKB_MAXD) return fail(SKB_EINVAL, "feature dimension %d outside 1..%d", d, SKB_MAXD);
    if (ld_train < d) return fail(SKB_EINVAL, "ld_train %lld < d %d", (long long)ld_train, d);
    if (!(h > 0.0) || !std::isfinite(h)) return fail(SKB_EINVAL, "bandwidth must be positive and finite");
    if (train_dtype != SKB_F64 && train_dtype != SKB_F32) return fail(SKB_EINVAL, "bad train dtype %d", train_dtype);
    if ((mu == nullptr) != (W == nullptr)) return fail(SKB_EINVAL, "mu and W must be given together");
    if (!train_is_whitened && !W) return fail(SKB_EINVAL, "raw training rows need the whitening map (mu, W)");
    if (N >= (1LL << 31)) return fail(SKB_EINVAL, "training set too large (row ids are 32-bit): shard it over GPUs");
    int ndev = skb_device_count();
    if (ndev <= 0) return fail(SKB_ECUDA, "no CUDA device available (this backend has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(SKB_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(SKB_ECUDA, "cudaSetDevice(%d) failed", device);

    skb_kde* k = new (std::nothrow) skb_kde();
    if (!k) return fail(SKB_ENOMEM, "out of host memory");
    k->device = device;
    cudaDeviceGetAttribute(&k->sms, cudaDevAttrMultiProcessorCount, device);
    KdeDev& hk = k->hk;
    memset(&hk, 0, sizeof(hk));
    hk.N = N;
    hk.D = d;
    const int TN = tile_points(d);
    hk.ntiles = (int)((N + TN - 1) / TN);
    hk.h = h;
    hk.c = std::sqrt(0.5 * 1.4426950408889634074) / h;   // sqrt(log2(e)/2)/h
    hk.has_w = (w != nullptr);
    hk.has_winv = (Winv != nullptr);
    for (int i = 0; i < d; i++) hk.mu[i] = mu ? mu[i] : 0.0;
    for (int i = 0; i < d; i++)
        for (int kk = 0; kk < d; kk++) {
            hk.Wc[i * d + kk] = (W ? W[i * d + kk] : (i == kk ? 1.0 : 0.0)) * hk.c;
            hk.Winv[i * d + kk] = Winv ? Winv[i * d + kk] : (i == kk ? 1.0 : 0.0);
        }

    int rc = SKB_OK;
    raise_pool_threshold(device);
    cudaStream_t st = nullptr;              // the fit runs on a private stream and is synchronised before returning
    std::vector<double> wh, cs;
    void* d_train_tmp = nullptr;
    double* d_w = nullptr;
    int* d_perm = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    auto cleanup = [&](int code) {
        if (st) {
            // fit-time temporaries come from the stream-ordered pool (cudaFree would synchronise the device three times)
            if (d_train_tmp) cudaFreeAsync(d_train_tmp, st);
            if (d_w) cudaFreeAsync(d_w, st);
            if (d_perm) cudaFreeAsync(d_perm, st);
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (code != SKB_OK) { skb_destroy(k); }
        return code;
    };
#define SKB_CUDA_C(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            cudaGetLastError();                                                                \
            return cleanup(fail(_e == cudaErrorMemoryAllocation ? SKB_ENOMEM : SKB_ECUDA, "%s: %s", #expr, \
                                cudaGetErrorString(_e)));                                      \
        }                                                                                      \
    } while (0)

    // Every upload below goes through THIS stream: a synchronous cudaMemcpy from pageable memory may return before the DMA
    // has landed, which only the legacy stream (not a non-blocking one) is ordered against.
    SKB_CUDA_C(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    SKB_CUDA_C(cudaEventCreateWithFlags(&ev0, cudaEventDisableTiming));
    SKB_CUDA_C(cudaEventRecord(ev0, 0));                     // device inputs produced on the legacy default stream (torch's
    SKB_CUDA_C(cudaStreamWaitEvent(st, ev0, 0));             // default) are complete before the fit reads them
    SKB_CUDA_C(cudaEventDestroy(ev0));
    ev0 = nullptr;
    // ---- weights: validate, sum, max, sequential cumsum (host, fp64: numpy's own rounding) ----
    double wmax = 1.0;
    hk.sum_w = (double)N;
    long long nv = N;                       // rows with positive weight
    if (w) {
        wh.resize((size_t)N);
        if (is_device_ptr(w)) SKB_CUDA_C(cudaMemcpy(wh.data(), w, (size_t)N * 8, cudaMemcpyDeviceToHost));
        else memcpy(wh.data(), w, (size_t)N * 8);
        wmax = 0.0;
        nv = 0;
        for (int64_t i = 0; i < N; i++) {
            const double v = wh[(size_t)i];
            if (!(v >= 0.0) || !std::isfinite(v))
                return cleanup(fail(SKB_EINVAL, "Negative values in data passed to `sample_weight` (or non-finite) at row %lld",
                                    (long long)i));
            if (v > wmax) wmax = v;
            nv += (v > 0.0);
        }
        if (!(wmax > 0.0)) return cleanup(fail(SKB_EINVAL, "all sample weights are zero"));
        {
            long double acc = 0.0L;          // differences to numpy's pairwise sum are O(1e-16) relative and only enter through log(sum_w)
            for (int64_t i = 0; i < N; i++) acc += wh[(size_t)i];
            hk.sum_w = (double)acc;
        }
        cs.resize((size_t)N);
        if (cumsum_w) {
            if (is_device_ptr(cumsum_w)) SKB_CUDA_C(cudaMemcpy(cs.data(), cumsum_w, (size_t)N * 8, cudaMemcpyDeviceToHost));
            else memcpy(cs.data(), cumsum_w, (size_t)N * 8);
        } else {
            double a = 0.0;                    // numpy.cumsum: strictly sequential fp64 adds
            for (int64_t i = 0; i < N; i++) { a += wh[(size_t)i]; cs[(size_t)i] = a; }
        }
    }
    hk.log_norm = skb_log_knorm(d, h) - std::log(hk.sum_w) + std::log(wmax);
    k->log_wmax = std::log(wmax);
    hk.ntiles = (int)((nv + TN - 1) / TN);
    hk.nsuper = (hk.ntiles + SG - 1) / SG;
    hk.nchunk = (hk.nsuper + CG - 1) / CG;
    // reference bias: the exponent window keeps terms down to 2^-(126 - bias) of the reference term, so the
    // worst-case truncation is N 2^-(126 - bias) relative; 90 up to 4e6 rows, one binade less per doubling beyond
    {
        int extra = 0;
        while (extra < 24 && ((long long)(1 << 22) << extra) < nv) extra++;
        hk.ref_bias = 90.f - (float)extra;
    }

    // ---- device buffers: one slab per handle (one allocation, one free) ----
    const size_t tile_bytes = (size_t)hk.ntiles * stage_floats(d) * sizeof(float);
    {
        size_t off = 0;
        auto carve = [&off](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        const size_t o_zt = carve((size_t)N * d * sizeof(double));
        const size_t o_cs = carve(w ? (size_t)N * sizeof(double) : 0);
        hk.cs_buckets = w ? (int)(N < (1ll << 24) ? N : (1ll << 24)) : 0;
        const size_t o_ci = carve(w ? ((size_t)hk.cs_buckets + 1) * sizeof(int) : 0);
        const size_t o_tiles = carve(tile_bytes);
        const size_t o_boxes = carve((size_t)hk.ntiles * 2 * d * sizeof(float));
        const size_t o_sb = carve((size_t)(hk.nsuper + hk.nchunk) * 2 * d * sizeof(float));   // super-boxes, then chunk boxes
        const size_t o_mom = carve(d == 1 ? (size_t)hk.ntiles * 12 * sizeof(float) : 0);
        const size_t o_ff2c = carve(d == 2 ? (size_t)hk.ntiles * 4 * FF2_CELL * sizeof(float) : 0);
        const size_t o_dk = carve(sizeof(KdeDev));
        SKB_CUDA_C(cudaMalloc(&k->slab, off));
        char* base = reinterpret_cast<char*>(k->slab);
        k->zt64 = reinterpret_cast<double*>(base + o_zt);
        k->cumsum = w ? reinterpret_cast<double*>(base + o_cs) : nullptr;
        hk.cs_index = w ? reinterpret_cast<int*>(base + o_ci) : nullptr;
        k->tiles = reinterpret_cast<float*>(base + o_tiles);
        k->boxes = reinterpret_cast<float*>(base + o_boxes);
        k->sboxes = reinterpret_cast<float*>(base + o_sb);
        k->dk = reinterpret_cast<KdeDev*>(base + o_dk);
        hk.mom = d == 1 ? reinterpret_cast<float*>(base + o_mom) : nullptr;
        hk.ff2c = d == 2 ? reinterpret_cast<float*>(base + o_ff2c) : nullptr;
    }
    if (w) {
        SKB_CUDA_C(cudaMemcpyAsync(k->cumsum, cs.data(), (size_t)N * 8, cudaMemcpyHostToDevice, st));
        SKB_CUDA_C(launch_cs_index(k->cumsum, N, const_cast<int*>(hk.cs_index), hk.cs_buckets, st));
        SKB_CUDA_C(cudaMallocAsync(reinterpret_cast<void**>(&d_w), (size_t)N * 8, st));
        SKB_CUDA_C(cudaMemcpyAsync(d_w, wh.data(), (size_t)N * 8, cudaMemcpyHostToDevice, st));
    }
    float* cboxes = k->sboxes + (size_t)hk.nsuper * 2 * d;
    SKB_CUDA_C(cudaMallocAsync(reinterpret_cast<void**>(&d_perm), (size_t)hk.ntiles * TN * sizeof(int), st));
    hk.zt64 = k->zt64;
    hk.cumsum_w = k->cumsum;
    hk.tiles = k->tiles;
    hk.boxes = k->boxes;
    hk.sboxes = k->sboxes;
    hk.cboxes = cboxes;
    SKB_CUDA_C(cudaMemcpyAsync(k->dk, &hk, sizeof(KdeDev), cudaMemcpyHostToDevice, st));

    const void* d_train = train;
    if (!is_device_ptr(train)) {
        const size_t esz = train_dtype == SKB_F32 ? 4 : 8;
        const size_t bytes = ((size_t)(N - 1) * ld_train + d) * esz;
        SKB_CUDA_C(cudaMallocAsync(&d_train_tmp, bytes, st));
        SKB_CUDA_C(cudaMemcpyAsync(d_train_tmp, train, bytes, cudaMemcpyHostToDevice, st));
        d_train = d_train_tmp;
    }
    // ---- the fit proper, all on the device: whitened fp64 rows (original order: the draw centres) -> kd order of
    //      the weighted rows (every tile of TN points is one leaf) -> packed tiles + the three box levels ----
    SKB_CUDA_C(cudaEventCreate(&ev0));
    SKB_CUDA_C(cudaEventCreate(&ev1));
    SKB_CUDA_C(cudaEventRecord(ev0, st));
    SKB_CUDA_C(launch_fit(d_train, train_dtype == SKB_F32, N, d, ld_train, train_is_whitened, k->dk, k->zt64, st));
    SKB_CUDA_C(launch_kd_order(k->zt64, d_w, N, nv, d, hk.c, d_perm, k->dk, st));     // also fills dk->glo / ghi
    SKB_CUDA_C(launch_pack(k->zt64, d_perm, d, hk.c, d_w, 1.0 / wmax, k->tiles, hk.ntiles, k->boxes, k->sboxes, cboxes, st));
    SKB_CUDA_C(launch_moments(k->tiles, k->boxes, hk.ntiles, const_cast<float*>(hk.mom), st));
    if (d == 2) SKB_CUDA_C(launch_moments2(k->tiles, k->boxes, hk.ntiles, const_cast<float*>(hk.ff2c), k->dk, st));
    SKB_CUDA_C(cudaEventRecord(ev1, st));
    SKB_CUDA_C(cudaMemcpyAsync(hk.glo, reinterpret_cast<char*>(k->dk) + offsetof(KdeDev, glo), sizeof(hk.glo) + sizeof(hk.ghi),
                               cudaMemcpyDeviceToHost, st));
    SKB_CUDA_C(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&k->fit_ms, ev0, ev1);
#undef SKB_CUDA_C
    rc = cleanup(SKB_OK);
    *out = k;
    return rc;
}

int skb_destroy(skb_handle h)
{
    if (!h) return SKB_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    if (h->slab) cudaFree(h->slab);
    cudaGetLastError();
    delete h;
    return SKB_OK;
}

int64_t skb_n_train(skb_handle h) { return h ? h->hk.N : 0; }
int skb_ndim(skb_handle h) { return h ? h->hk.D : 0; }
double skb_sum_weight(skb_handle h) { return h ? h->hk.sum_w : 0.0; }
int skb_device(skb_handle h) { return h ? h->device : -1; }
double skb_fit_ms(skb_handle h) { return h ? (double)h->fit_ms : 0.0; }

// ---- scoring ------------------------------------------------------------------------------------
static int check_query_args(skb_handle h, const void* q, int q_dtype, int64_t M, int64_t ldq)
{
    if (!h) return fail(SKB_EINVAL, "NULL handle");
    if (M < 0) return fail(SKB_EINVAL, "negative row count");
    if (M > 0 && !q) return fail(SKB_EINVAL, "NULL query pointer");
    if (q_dtype != SKB_F64 && q_dtype != SKB_F32) return fail(SKB_EINVAL, "bad query dtype %d", q_dtype);
    if (ldq < 1) return fail(SKB_EINVAL, "query row stride must be >= 1");
    return SKB_OK;
}

static int score_one(skb_handle h, const void* q, int q_dtype, int64_t M, int64_t ldq, const int32_t* cols,
                     int q_is_whitened, double* out_logp, double* out_m, double* out_s, void* stream)
{
    int rc = check_query_args(h, q, q_dtype, M, ldq);
    if (rc) return rc;
    if (M == 0) return SKB_OK;
    if (!cols && ldq < h->hk.D)
        return fail(SKB_EINVAL, "query has %lld features, the KDE was fitted with %d", (long long)ldq, h->hk.D);
    DeviceGuard guard(h->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool q_dev = is_device_ptr(q);
    double* outs[3] = {out_logp, out_m, out_s};
    bool any_host_out = false;
    for (double* o : outs) if (o && !is_device_ptr(o)) any_host_out = true;
    const GroupSpec g;
    const size_t esz = q_dtype == SKB_F32 ? 4 : 8;
    if (q_dev && !any_host_out) {
        // device resident: still walk the table in STAGE_ROWS chunks so the per-call scratch stays bounded
        for (int64_t a = 0; a < M; a += STAGE_ROWS) {
            const int64_t m = (M - a < STAGE_ROWS) ? (M - a) : STAGE_ROWS;
            Arena ar(st);
            DevJob dj{h, reinterpret_cast<const char*>(q) + (size_t)a * ldq * esz, q_dtype == SKB_F32, q_is_whitened,
                      m, ldq, cols, out_logp ? out_logp + a : nullptr, out_m ? out_m + a : nullptr,
                      out_s ? out_s + a : nullptr};
            rc = run_jobs_device(&dj, 1, false, g, ar, st);
            if (rc) return rc;
        }
        return SKB_OK;
    }
    // staged path: chunks of rows through device scratch, double buffered — the upload of chunk i+1 (copy stream)
    // overlaps the kernels of chunk i, and the download of chunk i is issued after the kernels of chunk i+1 are
    // queued, so that a blocking (pageable) download never leaves the GPU without work
    const int64_t step = host_stage_step(M);
    const int64_t nchunks = (M + step - 1) / step;
    void* qbuf[2] = {nullptr, nullptr};
    double* obuf[2] = {nullptr, nullptr};
    Arena keep(st);
    for (int b = 0; b < (nchunks > 1 ? 2 : 1); b++) {
        if (!q_dev) SKB_CUDA(keep.get(&qbuf[b], (size_t)step * ldq * esz));
        void* p = nullptr;
        SKB_CUDA(keep.get(&p, (size_t)step * 8 * 3));
        obuf[b] = reinterpret_cast<double*>(p);
    }
    HostStager hs;                                          // after `keep`: destroyed (and its stream drained) first
    SKB_CUDA(hs.init(st));
    auto upload = [&](int64_t c) -> cudaError_t {
        if (q_dev) return cudaSuccess;
        const int64_t a = c * step, m = (M - a < step) ? (M - a) : step;
        const char* qsrc = reinterpret_cast<const char*>(q) + (size_t)a * ldq * esz;
        return hs.upload(qbuf[c & 1], qsrc, (size_t)m * ldq * esz, (int)(c & 1));
    };
    auto download = [&](int64_t c) -> cudaError_t {
        const int64_t a = c * step, m = (M - a < step) ? (M - a) : step;
        const int b = (int)(c & 1);
        for (int t = 0; t < 3; t++)
            if (outs[t] && !is_device_ptr(outs[t])) {
                cudaError_t e = hs.download(outs[t] + a, obuf[b] + (size_t)t * m, (size_t)m * 8, b);
                if (e != cudaSuccess) return e;
            }
        return cudaSuccess;
    };
    SKB_CUDA(upload(0));
    for (int64_t c = 0; c < nchunks; c++) {
        const int64_t a = c * step, m = (M - a < step) ? (M - a) : step;
        const int b = (int)(c & 1);
        if (c + 1 < nchunks) SKB_CUDA(upload(c + 1));
        if (!q_dev) SKB_CUDA(hs.wait_upload(b));
        const void* qd = q_dev ? reinterpret_cast<const char*>(q) + (size_t)a * ldq * esz : qbuf[b];
        double* dst[3];
        for (int t = 0; t < 3; t++) {
            if (!outs[t]) dst[t] = nullptr;
            else if (is_device_ptr(outs[t])) dst[t] = outs[t] + a;
            else dst[t] = obuf[b] + (size_t)t * m;
        }
        SKB_CUDA(hs.wait_download(b));                     // the previous user of obuf[b] has been copied out
        {
            Arena ar(st);
            DevJob dj{h, qd, q_dtype == SKB_F32, q_is_whitened, m, ldq, cols, dst[0], dst[1], dst[2]};
            rc = run_jobs_device(&dj, 1, false, g, ar, st);
            if (rc) return rc;
        }
        SKB_CUDA(hs.kernels_done(b));
        if (c >= 1) SKB_CUDA(download(c - 1));
    }
    SKB_CUDA(download(nchunks - 1));
    SKB_CUDA(hs.finish());
    return SKB_OK;
}

int skb_score(skb_handle h, const void* q, int q_dtype, int64_t M, int64_t ldq, const int32_t* cols,
              int q_is_whitened, double* out_logp, void* stream)
{
    if (M > 0 && !out_logp) return fail(SKB_EINVAL, "NULL output pointer");
    return score_one(h, q, q_dtype, M, ldq, cols, q_is_whitened, out_logp, nullptr, nullptr, stream);
}

int skb_score_partial(skb_handle h, const void* q, int q_dtype, int64_t M, int64_t ldq, const int32_t* cols,
                      int q_is_whitened, double* out_m, double* out_s, void* stream)
{
    if (M > 0 && (!out_m || !out_s)) return fail(SKB_EINVAL, "NULL output pointer");
    return score_one(h, q, q_dtype, M, ldq, cols, q_is_whitened, nullptr, out_m, out_s, stream);
}

int skb_score_batch(const skb_job* jobs, int n_jobs, void* stream)
{
    if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail(SKB_EINVAL, "bad job list");
    if (n_jobs == 0) return SKB_OK;
    if (n_jobs > SKB_MAX_JOBS) return fail(SKB_EINVAL, "at most %d jobs per batch", SKB_MAX_JOBS);
    bool all_dev = true;
    for (int j = 0; j < n_jobs; j++) {
        int rc = check_query_args(jobs[j].kde, jobs[j].q, jobs[j].q_dtype, jobs[j].M, jobs[j].ldq);
        if (rc) return rc;
        if (jobs[j].M > 0 && !jobs[j].out_logp) return fail(SKB_EINVAL, "job %d: NULL output pointer", j);
        if (jobs[j].kde->device != jobs[0].kde->device) return fail(SKB_EINVAL, "jobs of one batch must share a device");
        if (jobs[j].M > 0 && (!is_device_ptr(jobs[j].q) || !is_device_ptr(jobs[j].out_logp))) all_dev = false;
        if (jobs[j].M > STAGE_ROWS) all_dev = false;        // oversized tables take the chunked single-job path
    }
    if (!all_dev) {
        // host buffers: serial staged calls (still correct, just not one launch)
        for (int j = 0; j < n_jobs; j++) {
            int rc = skb_score(jobs[j].kde, jobs[j].q, jobs[j].q_dtype, jobs[j].M, jobs[j].ldq, jobs[j].cols,
                               jobs[j].q_is_whitened, jobs[j].out_logp, stream);
            if (rc) return rc;
        }
        return SKB_OK;
    }
    DeviceGuard guard(jobs[0].kde->device);
    DevJob dj[SKB_MAX_JOBS];
    int n = 0;
    for (int j = 0; j < n_jobs; j++) {
        if (jobs[j].M == 0) continue;
        dj[n++] = DevJob{jobs[j].kde, jobs[j].q, jobs[j].q_dtype == SKB_F32, jobs[j].q_is_whitened, jobs[j].M,
                         jobs[j].ldq, jobs[j].cols, jobs[j].out_logp, nullptr, nullptr};
    }
    const GroupSpec g;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Arena ar(st);
    return run_jobs_device(dj, n, false, g, ar, st);
}

int skb_ratio_accept(const skb_handle* kdes, int n_kde, const int32_t* sign, const double* log_abs_jac,
                     double log_m, const void* q, int q_dtype, int64_t M, int64_t ldq, const int32_t* cols,
                     const double* u, double band, double* out_logp, double* out_log_ratio, uint8_t* accept,
                     uint8_t* in_band, void* stream)
{
    if (!kdes || n_kde < 1 || n_kde > SKB_MAX_GROUP) return fail(SKB_EINVAL, "group size must be 1..%d", SKB_MAX_GROUP);
    if (!sign || !log_abs_jac) return fail(SKB_EINVAL, "sign / log_abs_jac are required");
    for (int k = 0; k < n_kde; k++) {
        int rc = check_query_args(kdes[k], q, q_dtype, M, ldq);
        if (rc) return rc;
        if (kdes[k]->device != kdes[0]->device) return fail(SKB_EINVAL, "KDEs of one group must share a device");
        if (sign[k] != 1 && sign[k] != -1) return fail(SKB_EINVAL, "sign[%d] must be +1 or -1", k);
    }
    if (M == 0) return SKB_OK;
    skb_kde* h0 = kdes[0];
    DeviceGuard guard(h0->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t esz = q_dtype == SKB_F32 ? 4 : 8;
    const bool q_dev = is_device_ptr(q);
    const bool u_dev = !u || is_device_ptr(u);
    const bool lp_dev = !out_logp || is_device_ptr(out_logp);
    const bool lr_dev = !out_log_ratio || is_device_ptr(out_log_ratio);
    const bool ac_dev = !accept || is_device_ptr(accept);
    const bool ib_dev = !in_band || is_device_ptr(in_band);
    const bool all_dev = q_dev && u_dev && lp_dev && lr_dev && ac_dev && ib_dev;
    auto group_of = [&](GroupSpec& g) {
        g.ratio = 1; g.sign = sign; g.log_abs_jac = log_abs_jac; g.log_m = log_m; g.band = band;
    };
    if (all_dev) {
        for (int64_t a = 0; a < M; a += STAGE_ROWS) {
            const int64_t m = (M - a < STAGE_ROWS) ? (M - a) : STAGE_ROWS;
            GroupSpec g;
            group_of(g);
            g.u = u ? u + a : nullptr;
            g.out_log_ratio = out_log_ratio ? out_log_ratio + a : nullptr;
            g.accept = accept ? accept + a : nullptr;
            g.in_band = in_band ? in_band + a : nullptr;
            DevJob dj[SKB_MAX_GROUP];
            for (int k = 0; k < n_kde; k++)
                dj[k] = DevJob{kdes[k], reinterpret_cast<const char*>(q) + (size_t)a * ldq * esz, q_dtype == SKB_F32, 0, m, ldq,
                               cols ? cols + (size_t)k * SKB_MAXD : nullptr,
                               out_logp ? out_logp + (size_t)k * M + a : nullptr, nullptr, nullptr};
            Arena ar(st);
            int rc = run_jobs_device(dj, n_kde, true, g, ar, st);
            if (rc) return rc;
        }
        return SKB_OK;
    }
    // host buffers: chunks through device scratch with the copies on a side stream (see score_one)
    const int64_t step = host_stage_step(M);
    const int64_t nchunks = (M + step - 1) / step;
    const int nbuf = nchunks > 1 ? 2 : 1;
    void* qbuf[2] = {nullptr, nullptr};
    double* aux[2] = {nullptr, nullptr};      // [u | log_ratio | logp x n_kde] doubles
    unsigned char* fl[2] = {nullptr, nullptr};   // [accept | in_band]
    Arena keep(st);
    for (int b = 0; b < nbuf; b++) {
        void* p = nullptr;
        if (!q_dev) SKB_CUDA(keep.get(&qbuf[b], (size_t)step * ldq * esz));
        SKB_CUDA(keep.get(&p, (size_t)step * (2 + (size_t)n_kde) * 8));
        aux[b] = reinterpret_cast<double*>(p);
        SKB_CUDA(keep.get(&p, (size_t)step * 2));
        fl[b] = reinterpret_cast<unsigned char*>(p);
    }
    HostStager hs;
    SKB_CUDA(hs.init(st));
    auto upload = [&](int64_t c) -> cudaError_t {
        const int64_t a = c * step, m = (M - a < step) ? (M - a) : step;
        const int b = (int)(c & 1);
        cudaError_t e = cudaSuccess;
        if (!q_dev) e = hs.upload(qbuf[b], reinterpret_cast<const char*>(q) + (size_t)a * ldq * esz, (size_t)m * ldq * esz, b);
        if (e == cudaSuccess && u && !u_dev) e = hs.upload(aux[b], u + a, (size_t)m * 8, b);
        if (e == cudaSuccess) e = hs.mark_upload(b);
        return e;
    };
    auto download = [&](int64_t c) -> cudaError_t {
        const int64_t a = c * step, m = (M - a < step) ? (M - a) : step;
        const int b = (int)(c & 1);
        cudaError_t e = cudaSuccess;
        if (out_log_ratio && !lr_dev) e = hs.download(out_log_ratio + a, aux[b] + m, (size_t)m * 8, b);
        if (out_logp && !lp_dev)
            for (int k = 0; k < n_kde && e == cudaSuccess; k++)
                e = hs.download(out_logp + (size_t)k * M + a, aux[b] + (size_t)(2 + k) * m, (size_t)m * 8, b);
        if (e == cudaSuccess && accept && !ac_dev) e = hs.download(accept + a, fl[b], (size_t)m, b);
        if (e == cudaSuccess && in_band && !ib_dev) e = hs.download(in_band + a, fl[b] + m, (size_t)m, b);
        return e;
    };
    SKB_CUDA(upload(0));
    for (int64_t c = 0; c < nchunks; c++) {
        const int64_t a = c * step, m = (M - a < step) ? (M - a) : step;
        const int b = (int)(c & 1);
        if (c + 1 < nchunks) SKB_CUDA(upload(c + 1));
        SKB_CUDA(hs.wait_upload(b));
        SKB_CUDA(hs.wait_download(b));
        const void* qd = q_dev ? reinterpret_cast<const char*>(q) + (size_t)a * ldq * esz : qbuf[b];
        GroupSpec g;
        group_of(g);
        if (u) g.u = u_dev ? u + a : aux[b];
        if (out_log_ratio) g.out_log_ratio = lr_dev ? out_log_ratio + a : aux[b] + m;
        if (accept) g.accept = ac_dev ? accept + a : fl[b];
        if (in_band) g.in_band = ib_dev ? in_band + a : fl[b] + m;
        DevJob dj[SKB_MAX_GROUP];
        for (int k = 0; k < n_kde; k++) {
            double* lp = nullptr;
            if (out_logp) lp = lp_dev ? out_logp + (size_t)k * M + a : aux[b] + (size_t)(2 + k) * m;
            dj[k] = DevJob{kdes[k], qd, q_dtype == SKB_F32, 0, m, ldq, cols ? cols + (size_t)k * SKB_MAXD : nullptr,
                           lp, nullptr, nullptr};
        }
        {
            Arena ar(st);
            int rc = run_jobs_device(dj, n_kde, true, g, ar, st);
            if (rc) return rc;
        }
        SKB_CUDA(hs.kernels_done(b));
        if (c >= 1) SKB_CUDA(download(c - 1));
    }
    SKB_CUDA(download(nchunks - 1));
    SKB_CUDA(hs.finish());
    return SKB_OK;
}

// ---- draws ----------------------------------------------------------------------------------------
int skb_draw(skb_handle h, int64_t M, const double* u, const double* z, uint64_t seed, uint64_t offset,
             const int64_t* rows, int rcol, double rmin, double rmax, int max_attempts, int64_t* out_idx,
             double* out_x, int64_t ld_out, double* out_z, uint8_t* out_flag, int64_t* n_flagged, void* stream)
{
    if (!h) return fail(SKB_EINVAL, "NULL handle");
    if (M < 0) return fail(SKB_EINVAL, "negative draw count");
    if ((u == nullptr) != (z == nullptr)) return fail(SKB_EINVAL, "u and z streams must be given together");
    if (rcol >= h->hk.D) return fail(SKB_EINVAL, "window column %d out of range", rcol);
    if (out_x && ld_out < h->hk.D) return fail(SKB_EINVAL, "ld_out < d");
    if (n_flagged) *n_flagged = 0;
    if (M == 0) return SKB_OK;
    DeviceGuard guard(h->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int D = h->hk.D;
    // All-or-nothing residency: either every buffer is on the device, or everything is staged.
    const void* ptrs[] = {u, z, rows, out_idx, out_x, out_z, out_flag};
    bool any_host = false, any_dev = false;
    for (const void* p : ptrs) if (p) { if (is_device_ptr(p)) any_dev = true; else any_host = true; }
    if (any_host && any_dev) return fail(SKB_EINVAL, "skb_draw: mix of host and device buffers is not supported");
    Arena ar(st);
    unsigned long long* d_count = nullptr;
    if (n_flagged) {
        void* p = nullptr;
        SKB_CUDA(ar.get(&p, 8));
        d_count = reinterpret_cast<unsigned long long*>(p);
        SKB_CUDA(cudaMemsetAsync(d_count, 0, 8, st));
    }
    if (!any_host) {
        SKB_CUDA(launch_draw(h->dk, D, M, u, z, seed, offset, reinterpret_cast<const long long*>(rows), rcol, rmin,
                             rmax, max_attempts, reinterpret_cast<long long*>(out_idx), out_x, ld_out, out_z, out_flag,
                             n_flagged ? d_count : nullptr, st));
    } else {
        if (rows && (out_x || out_z || out_idx || out_flag)) {
            // scattered destination rows with host outputs: the caller owns the full arrays, so stage the
            // draws densely and scatter on the host.
        }
        void* blk = nullptr;
        {
            const size_t mmax = (size_t)(M < STAGE_ROWS ? M : STAGE_ROWS);
            SKB_CUDA(ar.get(&blk, mmax * (2 + 3 * (size_t)D) * 8 + mmax));
        }
        for (int64_t a = 0; a < M; a += STAGE_ROWS) {
            const int64_t m = (M - a < STAGE_ROWS) ? (M - a) : STAGE_ROWS;
            // layout of the staging block (doubles): u[m] z[m*D] x[m*D] zz[m*D] idx[m] then flags
            double* base = reinterpret_cast<double*>(blk);
            double* du = base; double* dz = du + m; double* dx = dz + (size_t)m * D; double* dzz = dx + (size_t)m * D;
            long long* didx = reinterpret_cast<long long*>(dzz + (size_t)m * D);
            unsigned char* dfl = reinterpret_cast<unsigned char*>(didx + m);
            if (u) {
                SKB_CUDA(cudaMemcpyAsync(du, u + a, (size_t)m * 8, cudaMemcpyHostToDevice, st));
                SKB_CUDA(cudaMemcpyAsync(dz, z + (size_t)a * D, (size_t)m * D * 8, cudaMemcpyHostToDevice, st));
            }
            SKB_CUDA(launch_draw(h->dk, D, m, u ? du : nullptr, u ? dz : nullptr, seed, offset + (uint64_t)a, nullptr,
                                 rcol, rmin, rmax, max_attempts, out_idx ? didx : nullptr, out_x ? dx : nullptr, D,
                                 out_z ? dzz : nullptr, out_flag ? dfl : nullptr, n_flagged ? d_count : nullptr, st));
            if (!rows) {
                if (out_idx) SKB_CUDA(cudaMemcpyAsync(out_idx + a, didx, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
                if (out_z) SKB_CUDA(cudaMemcpyAsync(out_z + (size_t)a * D, dzz, (size_t)m * D * 8, cudaMemcpyDeviceToHost, st));
                if (out_flag) SKB_CUDA(cudaMemcpyAsync(out_flag + a, dfl, (size_t)m, cudaMemcpyDeviceToHost, st));
                if (out_x) {
                    if (ld_out == D) SKB_CUDA(cudaMemcpyAsync(out_x + (size_t)a * D, dx, (size_t)m * D * 8, cudaMemcpyDeviceToHost, st));
                    else SKB_CUDA(cudaMemcpy2DAsync(out_x + (size_t)a * ld_out, (size_t)ld_out * 8, dx, (size_t)D * 8,
                                                    (size_t)D * 8, (size_t)m, cudaMemcpyDeviceToHost, st));
                }
                SKB_CUDA(cudaStreamSynchronize(st));
            } else {
                std::vector<double> hx, hz;
                std::vector<long long> hi;
                std::vector<unsigned char> hf;
                if (out_x) { hx.resize((size_t)m * D); SKB_CUDA(cudaMemcpyAsync(hx.data(), dx, (size_t)m * D * 8, cudaMemcpyDeviceToHost, st)); }
                if (out_z) { hz.resize((size_t)m * D); SKB_CUDA(cudaMemcpyAsync(hz.data(), dzz, (size_t)m * D * 8, cudaMemcpyDeviceToHost, st)); }
                if (out_idx) { hi.resize((size_t)m); SKB_CUDA(cudaMemcpyAsync(hi.data(), didx, (size_t)m * 8, cudaMemcpyDeviceToHost, st)); }
                if (out_flag) { hf.resize((size_t)m); SKB_CUDA(cudaMemcpyAsync(hf.data(), dfl, (size_t)m, cudaMemcpyDeviceToHost, st)); }
                SKB_CUDA(cudaStreamSynchronize(st));
                for (int64_t i = 0; i < m; i++) {
                    const int64_t o = rows[a + i];
                    if (out_x) memcpy(out_x + (size_t)o * ld_out, &hx[(size_t)i * D], (size_t)D * 8);
                    if (out_z) memcpy(out_z + (size_t)o * D, &hz[(size_t)i * D], (size_t)D * 8);
                    if (out_idx) out_idx[o] = hi[(size_t)i];
                    if (out_flag) out_flag[o] = hf[(size_t)i];
                }
            }
        }
    }
    if (n_flagged) {
        unsigned long long c = 0;
        SKB_CUDA(cudaMemcpyAsync(&c, d_count, 8, cudaMemcpyDeviceToHost, st));
        SKB_CUDA(cudaStreamSynchronize(st));
        *n_flagged = (int64_t)c;
    }
    return SKB_OK;
}

// ---- compaction -----------------------------------------------------------------------------------
int skb_compact(const uint8_t* flags, int64_t M, const void* rows, int dtype, int64_t ld, void* out,
                int64_t* out_index, int64_t* n_out, void* stream)
{
    if (M < 0) return fail(SKB_EINVAL, "negative row count");
    const bool n_out_dev = n_out && is_device_ptr(n_out);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n_out && !n_out_dev) *n_out = 0;
    if (M == 0) {
        if (n_out_dev) SKB_CUDA(cudaMemsetAsync(n_out, 0, 8, st));
        return SKB_OK;
    }
    if (!flags) return fail(SKB_EINVAL, "NULL flags");
    if (dtype != SKB_F64 && dtype != SKB_F32) return fail(SKB_EINVAL, "bad dtype %d", dtype);
    if ((rows == nullptr) != (out == nullptr)) return fail(SKB_EINVAL, "rows and out must be given together");
    if (rows && ld < 1) return fail(SKB_EINVAL, "bad row stride");
    const void* ptrs[] = {flags, rows, out, out_index};
    for (const void* p : ptrs)
        if (p && !is_device_ptr(p)) return fail(SKB_EINVAL, "skb_compact works on device buffers only");
    const long long nb = compact_ws_words(M) - 1;                       // ticket + one state word per chunk, then the total
    unsigned long long* ws = nullptr;
    SKB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), (size_t)(nb + 1) * 8, st));
    cudaError_t e = launch_compact(flags, M, rows, dtype == SKB_F32 ? 4 : 8, ld, out,
                                   reinterpret_cast<long long*>(out_index), ws, ws + nb, st);
    if (e == cudaSuccess && n_out) {
        if (n_out_dev) {
            e = cudaMemcpyAsync(n_out, ws + nb, 8, cudaMemcpyDeviceToDevice, st);       // stays asynchronous
        } else {
            unsigned long long c = 0;
            e = cudaMemcpyAsync(&c, ws + nb, 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            *n_out = (int64_t)c;
        }
    }
    cudaFreeAsync(ws, st);
    SKB_CUDA(e);
    return SKB_OK;
}

int skb_features(const skb_feature_spec* specs, int n_feat, int64_t n_rows, double* out, uint8_t* keep, void* stream)
{
    if (!specs || n_feat < 1 || n_feat > SKB_MAXD) return fail(SKB_EINVAL, "feature count must be 1..%d", SKB_MAXD);
    if (n_rows < 0) return fail(SKB_EINVAL, "negative row count");
    if (n_rows == 0) return SKB_OK;
    if (skb_device_count() <= 0) return fail(SKB_ECUDA, "no CUDA device available (this backend has no CPU fallback)");
    if (!out || !keep || !is_device_ptr(out) || !is_device_ptr(keep)) return fail(SKB_EINVAL, "skb_features works on device buffers only");
    for (int f = 0; f < n_feat; f++) {
        const skb_feature_spec& s = specs[f];
        if (s.op < 0 || s.op > 5) return fail(SKB_EINVAL, "feature %d: only + - * / and SQSUM are supported", f);
        if (s.a && (!is_device_ptr(s.a) || (s.a_dtype != SKB_F64 && s.a_dtype != SKB_F32) || s.a_stride < 1))
            return fail(SKB_EINVAL, "feature %d: operand a must be a float32 / float64 device column", f);
        if (s.op != 0 && s.b && (!is_device_ptr(s.b) || (s.b_dtype != SKB_F64 && s.b_dtype != SKB_F32) || s.b_stride < 1))
            return fail(SKB_EINVAL, "feature %d: operand b must be a float32 / float64 device column", f);
    }
    SKB_CUDA(launch_features(specs, n_feat, n_rows, out, keep, reinterpret_cast<cudaStream_t>(stream)));
    return SKB_OK;
}

int skb_set_option(const char* name, int value)
{
    if (!name) return fail(SKB_EINVAL, "NULL option name");
    if (!strcmp(name, "cull")) { g_opt_cull.store(value ? 1 : 0); return SKB_OK; }
    if (!strcmp(name, "force_nsplit")) { g_opt_nsplit.store(value > 0 ? value : 0); return SKB_OK; }
    if (!strcmp(name, "stats")) { g_opt_stats.store(value ? 1 : 0); return SKB_OK; }
    if (!strcmp(name, "kernel")) { g_opt_kernel.store(value); return SKB_OK; }
    if (!strcmp(name, "host_stage_rows")) { g_opt_host_rows.store(value > 0 ? value : 0); return SKB_OK; }
    if (!strcmp(name, "ff2_cells")) { g_opt_ff2_cells.store(value < 0 ? -1 : (value ? 1 : 0)); return SKB_OK; }
    return fail(SKB_EINVAL, "unknown option '%s' (cull, force_nsplit, stats)", name);
}

int skb_get_option(const char* name)
{
    if (!name) return -1;
    if (!strcmp(name, "cull")) return g_opt_cull.load();
    if (!strcmp(name, "force_nsplit")) return g_opt_nsplit.load();
    if (!strcmp(name, "stats")) return g_opt_stats.load();
    if (!strcmp(name, "kernel")) return g_opt_kernel.load();
    if (!strcmp(name, "host_stage_rows")) return (int)g_opt_host_rows.load();
    if (!strcmp(name, "ff2_cells")) return g_opt_ff2_cells.load();
    return -1;
}

int skb_debug_counters(int device, uint64_t* out4, int reset)
{
    if (!out4) return fail(SKB_EINVAL, "NULL output pointer");
    if (skb_device_count() <= 0) return fail(SKB_ECUDA, "no CUDA device available");
    for (int i = 0; i < 4; i++) out4[i] = 0;
    if (device < 0 || device >= 64 || !g_stats_dev[device]) return SKB_OK;      // nothing was counted (skb_set_option("stats", 1) first)
    DeviceGuard guard(device);
    unsigned long long tmp[4];
    SKB_CUDA(cudaDeviceSynchronize());
    SKB_CUDA(cudaMemcpy(tmp, g_stats_dev[device], sizeof(tmp), cudaMemcpyDeviceToHost));
    if (reset) SKB_CUDA(cudaMemset(g_stats_dev[device], 0, sizeof(tmp)));
    for (int i = 0; i < 4; i++) out4[i] = tmp[i];
    return SKB_OK;
}

int skb_measure_pipes(int device, double ms_each, double* fma_lane_per_s, double* ex2_per_s)
{
    if (!fma_lane_per_s || !ex2_per_s) return fail(SKB_EINVAL, "NULL output pointer");
    int ndev = skb_device_count();
    if (ndev <= 0) return fail(SKB_ECUDA, "no CUDA device available");
    if (device < 0 || device >= ndev) return fail(SKB_EINVAL, "device %d out of range", device);
    DeviceGuard guard(device);
    SKB_CUDA(measure_pipes(ms_each > 0 ? ms_each : 50.0, fma_lane_per_s, ex2_per_s));
    return SKB_OK;
}

}  // extern "C"

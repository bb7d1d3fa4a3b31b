// KDE draws, order-preserving compaction and the live pipe-peak probes (sm_100a).
//
//   skb_draw_kernel     KernelDensity.sample (sklearn neighbors/_kde.py:338-349) fused with
//                       pca_inverse_transform (skysampler/emulator.py:329-335) and the radial
//                       window of random_draw (emulator.py:362-375).  HBM/latency bound.
//   compact_*           `samples[inds]` (emulator.py:746-747): flags -> packed rows, input order kept.
//   pipe probes         FFMA2 and MUFU.EX2 saturation loops: the roofline denominators of the
//                       score kernel, measured on the device it runs on.
#include "skb_internal.h"
#include <math.h>

namespace skb {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (draw index lo, hi, attempt, lane), key = seed.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b)   // numpy's 53-bit construction
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1)
{
    const float u1 = (float)a * 2.3283064365386963e-10f + 1.1641532182693481e-10f;  // (a + 0.5) / 2^32
    const float u2 = (float)b * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    n0 = rad * c; n1 = rad * s;
}

// np.searchsorted(cs, target) (side='left') restricted to [lo, hi]; the caller guarantees the answer lies there.
__device__ __forceinline__ long long lower_bound_f64(const double* __restrict__ cs, long long lo, long long hi, double target)
{
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(cs + mid) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Bucket index of the cumulative weights: index[b] = searchsorted(cs, b / B * cs[N-1]), b = 0..B.  A draw with uniform u
// then only has to search cs[index[b-1] .. index[b+2]] with b = floor(u B) (about three entries when B = N) instead of
// taking ~20 dependent probes of the whole array; the result is the same index bit for bit.
__global__ void __launch_bounds__(256)
cs_index_kernel(const double* __restrict__ cs, long long N, int* __restrict__ index, int B)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    const double target = ((double)b / (double)B) * __ldg(cs + N - 1);
    index[b] = (int)lower_bound_f64(cs, 0, N, target);
}

cudaError_t launch_cs_index(const double* cumsum, long long N, int* index, int buckets, cudaStream_t st)
{
    if (!cumsum || buckets <= 0) return cudaSuccess;
    cs_index_kernel<<<(unsigned)(((long long)buckets + 1 + 255) / 256), 256, 0, st>>>(cumsum, N, index, buckets);
    count_launch();
    return cudaGetLastError();
}

// D is a template parameter: the per-draw vectors live in registers (with a run-time D they were local-memory arrays).
template <int D>
__global__ void __launch_bounds__(256)
skb_draw_kernel(const KdeDev* __restrict__ kde, long long M,
                const double* __restrict__ u_in, const double* __restrict__ z_in,
                uint64_t seed, uint64_t offset, const long long* __restrict__ rows,
                int rcol, double rmin, double rmax, int max_attempts,
                long long* __restrict__ out_idx, double* __restrict__ out_x, long long ld_out,
                double* __restrict__ out_z, unsigned char* __restrict__ out_flag,
                unsigned long long* __restrict__ n_flagged)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const long long N = kde->N;
    const double* __restrict__ cs = kde->cumsum_w;
    const int* __restrict__ ci = kde->cs_index;
    const long long nb = kde->cs_buckets;
    const double h = kde->h;
    const bool streams = (u_in != nullptr);
    const bool has_winv = kde->has_winv != 0;
    const uint64_t cidx = (uint64_t)i + offset;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));

    long long idx = 0;
    double zs[D], x[D];
    bool bad = false;
    const int nattempt = streams ? 1 : (max_attempts > 0 ? max_attempts : 1);
    for (int attempt = 0; attempt < nattempt; attempt++) {
        double u;
        if (streams) {
            u = u_in[i];
        } else {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)cidx, (uint32_t)(cidx >> 32), (uint32_t)attempt, 0u), key);
            u = u53(r.x, r.y);
        }
        if (cs) {
            // np.searchsorted(cumsum_weight, u * sum_weight)  (side='left')   sklearn _kde.py:345-347
            const double target = u * __ldg(cs + N - 1);
            long long lo = 0, hi = N;
            if (ci && u >= 0.0 && u < 1.0) {
                // u lies in [(b - 1) / B, (b + 2) / B) whatever way u * B rounded, and the fp64 product with cs[N-1] is
                // monotone in u: the answer is bracketed by the index entries of those two bucket edges
                long long b = (long long)(u * (double)nb);
                b = b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
                lo = __ldg(ci + (b > 0 ? b - 1 : 0));
                hi = __ldg(ci + (b + 2 < nb ? b + 2 : nb));
            }
            lo = lower_bound_f64(cs, lo, hi, target);
            idx = lo < N ? lo : N - 1;
        } else {
            idx = (long long)(u * (double)N);              // (u * N).astype(int64)   sklearn _kde.py:343
            if (idx >= N) idx = N - 1;
        }
        const double* __restrict__ ctr = kde->zt64 + idx * D;
        double cz[D];
        if (D % 2 == 0) {                                  // rows of an even number of doubles are 16-byte aligned in the slab
#pragma unroll
            for (int k = 0; k < D; k += 2) {
                const double2 v = __ldg(reinterpret_cast<const double2*>(ctr + k));
                cz[k] = v.x; cz[k + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; k++) cz[k] = __ldg(ctr + k);
        }
        if (streams) {
#pragma unroll
            for (int k = 0; k < D; k++) zs[k] = __dadd_rn(cz[k], __dmul_rn(h, z_in[i * D + k]));   // rng.normal(data[i], h): loc + scale*g, two roundings like numpy
        } else {
#pragma unroll
            for (int k = 0; k < D; k += 4) {
                const uint4 r = philox4x32_10(make_uint4((uint32_t)cidx, (uint32_t)(cidx >> 32), (uint32_t)attempt,
                                                         1u + (uint32_t)(k >> 2)), key);
                float n[4];
                box_muller(r.x, r.y, n[0], n[1]);
                box_muller(r.z, r.w, n[2], n[3]);
#pragma unroll
                for (int t = 0; t < 4; t++)
                    if (k + t < D) zs[k + t] = cz[k + t] + h * (double)n[t];
            }
        }
        if (has_winv) {
#pragma unroll
            for (int c = 0; c < D; c++) {
                double a = 0.0;
#pragma unroll
                for (int k = 0; k < D; k++) a = fma(zs[k], kde->Winv[k * D + c], a);
                x[c] = a + kde->mu[c];
            }
        } else {
#pragma unroll
            for (int c = 0; c < D; c++) x[c] = zs[c];
        }
        bad = false;
        if (rcol >= 0) {
            double xr = x[0];
#pragma unroll
            for (int c = 1; c < D; c++) xr = (c == rcol) ? x[c] : xr;
            bad = (xr > rmax) || (xr < rmin);                          // emulator.py:370
        }
        if (!bad) break;
    }
    // streaming stores (evict-first): the output passes through L2 once and must not push the centre rows (zt64), the
    // cumulative weights and their index out of it — the gathers of the next draws hit L2 instead of DRAM
    const long long o = rows ? rows[i] : i;
    if (out_idx) __stcs(out_idx + o, idx);
    if (out_x) {
        if ((D % 2 == 0) && ((ld_out & 1) == 0) && ((reinterpret_cast<uintptr_t>(out_x) & 15) == 0)) {
#pragma unroll
            for (int c = 0; c < D; c += 2) __stcs(reinterpret_cast<double2*>(out_x + o * ld_out + c), make_double2(x[c], x[c + 1]));
        } else {
#pragma unroll
            for (int c = 0; c < D; c++) __stcs(out_x + o * ld_out + c, x[c]);
        }
    }
    if (out_z) {
#pragma unroll
        for (int k = 0; k < D; k++) __stcs(out_z + o * D + k, zs[k]);
    }
    if (out_flag) out_flag[o] = bad ? 1 : 0;
    if (n_flagged) {
        const unsigned ballot = __ballot_sync(__activemask(), bad);
        if (bad && (threadIdx.x & 31) == (__ffs(ballot) - 1)) atomicAdd(n_flagged, (unsigned long long)__popc(ballot));
    }
}

cudaError_t launch_draw(const KdeDev* kde, int D, long long M, const double* u, const double* z,
                        uint64_t seed, uint64_t offset, const long long* rows,
                        int rcol, double rmin, double rmax, int max_attempts,
                        long long* out_idx, double* out_x, long long ld_out, double* out_z,
                        unsigned char* out_flag, unsigned long long* n_flagged, cudaStream_t st)
{
    if (M <= 0) return cudaSuccess;
    const int bs = 256;
    const unsigned grid = (unsigned)((M + bs - 1) / bs);
    switch (D) {
#define SKB_CASE(d) case d: skb_draw_kernel<d><<<grid, bs, 0, st>>>(kde, M, u, z, seed, offset, rows, rcol, rmin, rmax, max_attempts, \
                                                                  out_idx, out_x, ld_out, out_z, out_flag, n_flagged); break;
        SKB_CASE(1) SKB_CASE(2) SKB_CASE(3) SKB_CASE(4) SKB_CASE(5) SKB_CASE(6) SKB_CASE(7) SKB_CASE(8)
        SKB_CASE(9) SKB_CASE(10) SKB_CASE(11) SKB_CASE(12) SKB_CASE(13) SKB_CASE(14) SKB_CASE(15) SKB_CASE(16)
#undef SKB_CASE
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Order-preserving compaction in ONE pass over the flags (chained scan with decoupled look-back, Merrill & Garland 2016):
// a CTA takes the next chunk of CP_CHUNK flags by ticket, counts its accepted rows, publishes the count, looks back over
// its predecessors' published counts / prefixes for its output offset, and copies its accepted rows.  The flags are read
// once (the three-kernel count / scan / scatter version read them twice and paid three launches).
// Workspace (zeroed by launch_compact): ws[0] = ticket, ws[1 + c] = state of chunk c: bits 62..63 = 0 nothing yet,
// 1 = the chunk's own count, 2 = inclusive prefix up to and including the chunk; bits 0..61 = the value.
// ---------------------------------------------------------------------------------------------
constexpr int CP_THREADS = 256;
constexpr int CP_PER_THREAD = 16;
constexpr int CP_CHUNK = CP_THREADS * CP_PER_THREAD;
constexpr unsigned long long CP_AGG = 1ull << 62, CP_PFX = 2ull << 62, CP_VAL = (1ull << 62) - 1;

long long compact_nblocks(long long M) { return (M + CP_CHUNK - 1) / CP_CHUNK; }
long long compact_ws_words(long long M) { return compact_nblocks(M) + 2; }           // ticket, chunk states, total

__device__ __forceinline__ unsigned nonzero_bytes4(unsigned v)      // bit t set <=> byte t of v is non-zero
{
    // per byte: (b | (b + 0x7f)) has its top bit set iff b != 0  (no carries across bytes: the low 7 bits are added alone)
    const unsigned t = ((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v;
    const unsigned h = t & 0x80808080u;
    return ((h >> 7) & 1u) | ((h >> 14) & 2u) | ((h >> 21) & 4u) | ((h >> 28) & 8u);
}

__device__ __forceinline__ unsigned load_flags16(const unsigned char* __restrict__ flags, long long base, long long M)
{
    // bit t set <=> flags[base + t] != 0, t = 0..15
    unsigned m = 0;
    if (base + 16 <= M && ((reinterpret_cast<uintptr_t>(flags + base) & 15) == 0)) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(flags + base));
        m = nonzero_bytes4(v.x) | (nonzero_bytes4(v.y) << 4) | (nonzero_bytes4(v.z) << 8) | (nonzero_bytes4(v.w) << 12);
    } else {
        for (int t = 0; t < 16; t++) if (base + t < M && flags[base + t]) m |= 1u << t;
    }
    return m;
}

__global__ void __launch_bounds__(CP_THREADS)
compact_kernel(const unsigned char* __restrict__ flags, long long M, const unsigned char* __restrict__ rows,
               int row_bytes, unsigned char* __restrict__ out, long long* __restrict__ out_index,
               unsigned long long* __restrict__ ws, long long nb, unsigned long long* __restrict__ total_out)
{
    __shared__ int wsum[CP_THREADS / 32];
    __shared__ unsigned short sel[CP_CHUNK];             // accepted rows of this chunk (offset in the chunk), in input order
    __shared__ long long s_chunk;
    __shared__ unsigned long long s_excl;
    volatile unsigned long long* state = ws + 1;
    if (threadIdx.x == 0) s_chunk = (long long)atomicAdd(ws, 1ull);          // chunks start in index order: look-back cannot deadlock
    __syncthreads();
    const long long chunk = s_chunk;
    if (chunk >= nb) return;
    const long long cbase = chunk * CP_CHUNK;
    const long long base = cbase + (long long)threadIdx.x * CP_PER_THREAD;
    const unsigned m = base < M ? load_flags16(flags, base, M) : 0u;
    const int c = __popc(m);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < CP_THREADS / 32; w++) {
        if (w < (int)(threadIdx.x >> 5)) wbase += wsum[w];
        total += wsum[w];
    }
    // publish this chunk's count, then look back (warp 0): sum the predecessors' counts until one carries a full prefix
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        if (lane == 0) {
            state[chunk] = (chunk == 0 ? 2ull << 62 : 1ull << 62) | (unsigned long long)total;
            __threadfence();
        }
        unsigned long long excl = 0;
        long long j = chunk - 1;
        while (j >= 0) {
            const long long jj = j - lane;
            unsigned long long v = 0;
            bool done;
            do {                                            // every lane's predecessor must have published something
                v = jj >= 0 ? state[jj] : (2ull << 62);
                done = (v >> 62) != 0ull;
            } while (!__all_sync(0xffffffffu, done));
            const unsigned pfx = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
            const int stop = pfx ? __ffs(pfx) - 1 : 32;      // nearest predecessor with a full prefix (lane order = distance)
            unsigned long long add = (lane <= stop) ? (v & CP_VAL) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
            excl += add;
            if (pfx) break;
            j -= 32;
        }
        if (lane == 0) {
            if (chunk > 0) { __threadfence(); state[chunk] = CP_PFX | (excl + (unsigned long long)total); }
            s_excl = excl;
            if (chunk == nb - 1) *total_out = excl + (unsigned long long)total;
        }
    }
    int pos = wbase + incl - c;
#pragma unroll
    for (int t = 0; t < CP_PER_THREAD; t++)
        if ((m >> t) & 1u) sel[pos++] = (unsigned short)(threadIdx.x * CP_PER_THREAD + t);
    __syncthreads();
    if (total == 0) return;
    const unsigned long long dst0 = s_excl;
    if (out_index)
        for (int r = threadIdx.x; r < total; r += CP_THREADS) out_index[dst0 + r] = cbase + sel[r];
    if (!out) return;
    // the whole CTA copies the accepted rows: consecutive threads move consecutive 16-byte (or 4- / 1-byte) pieces, so the
    // writes are fully coalesced and the reads are coalesced within each row
    const bool a16 = (row_bytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const bool a4 = (row_bytes & 3) == 0 && ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(out)) & 3) == 0;
    const int unit = a16 ? 16 : (a4 ? 4 : 1);
    const int upr = row_bytes / unit;
    const int nunits = total * upr;
    if (a16 && (upr & (upr - 1)) == 0) {                 // power-of-two pieces per row (8-D fp64 rows: 4): shifts, no divisions
        const int sh = __ffs(upr) - 1;
        for (int i = threadIdx.x; i < nunits; i += CP_THREADS) {
            const int r = i >> sh, part = i & (upr - 1);
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(rows + (size_t)(cbase + sel[r]) * row_bytes) + part);
            reinterpret_cast<uint4*>(out + (size_t)(dst0 + r) * row_bytes)[part] = v;
        }
        return;
    }
    for (int i = threadIdx.x; i < nunits; i += CP_THREADS) {
        const int r = i / upr, part = i % upr;
        const unsigned char* sp = rows + (size_t)(cbase + sel[r]) * row_bytes + (size_t)part * unit;
        unsigned char* dp = out + (size_t)(dst0 + r) * row_bytes + (size_t)part * unit;
        if (a16) *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<const uint4*>(sp);
        else if (a4) *reinterpret_cast<unsigned*>(dp) = *reinterpret_cast<const unsigned*>(sp);
        else *dp = *sp;
    }
}

// ws: compact_ws_words(M) words of scratch; the accepted count lands in ws[compact_ws_words(M) - 1] (= total)
cudaError_t launch_compact(const unsigned char* flags, long long M, const void* rows, int elem_bytes,
                           long long ld, void* out, long long* out_index,
                           unsigned long long* ws, unsigned long long* total, cudaStream_t st)
{
    const long long nb = compact_nblocks(M);
    if (nb == 0) return cudaMemsetAsync(total, 0, sizeof(unsigned long long), st);
    cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)(nb + 1) * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    compact_kernel<<<(unsigned)nb, CP_THREADS, 0, st>>>(flags, M, reinterpret_cast<const unsigned char*>(rows),
                                                        (int)(elem_bytes * ld), reinterpret_cast<unsigned char*>(out),
                                                        out_index, ws, nb, total);
    count_launch(1);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Pipe probes
// ---------------------------------------------------------------------------------------------
constexpr int PROBE_ITERS = 2048;

__global__ void __launch_bounds__(256) probe_ffma2_kernel(float* out, float seed)
{
    u64 p[8];
    float a = seed * 0.999f, b = seed * 1e-3f;
    u64 pa, pb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float v = seed + threadIdx.x * 1e-3f + i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(p[i]) : "f"(v));
    }
    for (int it = 0; it < PROBE_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi; }
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) probe_ex2_kernel(float* out, float seed)
{
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = seed * 1e-3f * (threadIdx.x + i);
    for (int it = 0; it < PROBE_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 123.456f) out[0] = s;
}

cudaError_t measure_pipes(double ms_each, double* fma_lane_per_s, double* ex2_per_s)
{
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* out = nullptr;
    if ((e = cudaMalloc(&out, 4)) != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int nb = sms * 8;
    for (int which = 0; which < 2; which++) {
        // warm up + calibrate the repeat count, then time one long burst
        float ms = 0.f;
        int reps = 4;
        for (int pass = 0; pass < 2; pass++) {
            cudaEventRecord(e0);
            for (int r = 0; r < reps; r++) {
                if (which == 0) probe_ffma2_kernel<<<nb, 256>>>(out, 1.0f);
                else probe_ex2_kernel<<<nb, 256>>>(out, 1.0f);
            }
            cudaEventRecord(e1);
            if ((e = cudaEventSynchronize(e1)) != cudaSuccess) { cudaFree(out); return e; }
            cudaEventElapsedTime(&ms, e0, e1);
            count_launch(reps);
            if (pass == 0) {
                const double per = ms / reps;
                reps = (int)(ms_each / (per > 1e-4 ? per : 1e-4)) + 1;
                if (reps > 20000) reps = 20000;
            }
        }
        const double ops = (double)nb * 256.0 * PROBE_ITERS * reps;
        if (which == 0) *fma_lane_per_s = ops * 16.0 / (ms * 1e-3);
        else *ex2_per_s = ops * 8.0 / (ms * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError();
}

}  // namespace skb

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
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (__ldg(cs + mid) < target) lo = mid + 1; else hi = mid;
            }
            idx = lo < N ? lo : N - 1;
        } else {
            idx = (long long)(u * (double)N);              // (u * N).astype(int64)   sklearn _kde.py:343
            if (idx >= N) idx = N - 1;
        }
        const double* __restrict__ ctr = kde->zt64 + idx * D;
        if (streams) {
#pragma unroll
            for (int k = 0; k < D; k++) zs[k] = __dadd_rn(__ldg(ctr + k), __dmul_rn(h, z_in[i * D + k]));   // rng.normal(data[i], h): loc + scale*g, two roundings like numpy
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
                    if (k + t < D) zs[k + t] = __ldg(ctr + k + t) + h * (double)n[t];
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
    const long long o = rows ? rows[i] : i;
    if (out_idx) out_idx[o] = idx;
    if (out_x) {
        if ((D % 2 == 0) && ((ld_out & 1) == 0) && ((reinterpret_cast<uintptr_t>(out_x) & 15) == 0)) {
#pragma unroll
            for (int c = 0; c < D; c += 2) *reinterpret_cast<double2*>(out_x + o * ld_out + c) = make_double2(x[c], x[c + 1]);
        } else {
#pragma unroll
            for (int c = 0; c < D; c++) out_x[o * ld_out + c] = x[c];
        }
    }
    if (out_z) {
#pragma unroll
        for (int k = 0; k < D; k++) out_z[o * D + k] = zs[k];
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
// Order-preserving compaction: per-chunk counts -> exclusive scan -> scatter.
// ---------------------------------------------------------------------------------------------
constexpr int CP_THREADS = 256;
constexpr int CP_PER_THREAD = 8;
constexpr int CP_CHUNK = CP_THREADS * CP_PER_THREAD;

long long compact_nblocks(long long M) { return (M + CP_CHUNK - 1) / CP_CHUNK; }

__device__ __forceinline__ unsigned load_flags8(const unsigned char* __restrict__ flags, long long base, long long M)
{
    // bit t set <=> flags[base + t] != 0, t = 0..7
    unsigned m = 0;
    if (base + 8 <= M && ((reinterpret_cast<uintptr_t>(flags + base) & 7) == 0)) {
        const uint2 v = *reinterpret_cast<const uint2*>(flags + base);
#pragma unroll
        for (int t = 0; t < 4; t++) {
            if ((v.x >> (8 * t)) & 0xff) m |= 1u << t;
            if ((v.y >> (8 * t)) & 0xff) m |= 1u << (4 + t);
        }
    } else {
        for (int t = 0; t < 8; t++) if (base + t < M && flags[base + t]) m |= 1u << t;
    }
    return m;
}

__global__ void __launch_bounds__(CP_THREADS)
compact_count_kernel(const unsigned char* __restrict__ flags, long long M, unsigned long long* __restrict__ counts)
{
    const long long base = (long long)blockIdx.x * CP_CHUNK + (long long)threadIdx.x * CP_PER_THREAD;
    int c = base < M ? __popc(load_flags8(flags, base, M)) : 0;
    __shared__ int wsum[CP_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < CP_THREADS / 32; w++) t += wsum[w];
        counts[blockIdx.x] = (unsigned long long)t;
    }
}

__global__ void __launch_bounds__(1024)
compact_scan_kernel(unsigned long long* __restrict__ counts, long long nb, unsigned long long* __restrict__ total)
{
    // single CTA, in-place exclusive scan of nb chunk counts
    __shared__ unsigned long long wtot[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long long b0 = 0; b0 < nb; b0 += 1024) {
        const long long i = b0 + threadIdx.x;
        const unsigned long long v = i < nb ? counts[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) wbase += wtot[w];
        const unsigned long long carry = carry_s;
        if (i < nb) counts[i] = carry + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wbase + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(CP_THREADS)
compact_scatter_kernel(const unsigned char* __restrict__ flags, long long M, const unsigned char* __restrict__ rows,
                       int row_bytes, unsigned char* __restrict__ out, long long* __restrict__ out_index,
                       const unsigned long long* __restrict__ offsets)
{
    const long long cbase = (long long)blockIdx.x * CP_CHUNK;
    const long long base = cbase + (long long)threadIdx.x * CP_PER_THREAD;
    const unsigned m = base < M ? load_flags8(flags, base, M) : 0u;
    const int c = __popc(m);
    __shared__ int wsum[CP_THREADS / 32];
    __shared__ unsigned short sel[CP_CHUNK];             // accepted rows of this chunk (offset in the chunk), in input order
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < CP_THREADS / 32; w++) {
        if (w < (int)(threadIdx.x >> 5)) wbase += wsum[w];
        total += wsum[w];
    }
    int pos = wbase + incl - c;
    for (int t = 0; t < 8; t++)
        if ((m >> t) & 1u) sel[pos++] = (unsigned short)(threadIdx.x * CP_PER_THREAD + t);
    __syncthreads();
    const unsigned long long dst0 = offsets[blockIdx.x];
    if (out_index)
        for (int r = threadIdx.x; r < total; r += CP_THREADS) out_index[dst0 + r] = cbase + sel[r];
    if (!out) return;
    // the whole CTA copies the accepted rows: consecutive threads move consecutive 16-byte (or 4- / 1-byte) pieces, so the
    // writes are fully coalesced and the reads are coalesced within each row
    const bool a16 = (row_bytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const bool a4 = (row_bytes & 3) == 0 && ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(out)) & 3) == 0;
    const int unit = a16 ? 16 : (a4 ? 4 : 1);
    const int upr = row_bytes / unit;
    const long long nunits = (long long)total * upr;
    for (long long i = threadIdx.x; i < nunits; i += CP_THREADS) {
        const int r = (int)(i / upr), part = (int)(i % upr);
        const unsigned char* sp = rows + (size_t)(cbase + sel[r]) * row_bytes + (size_t)part * unit;
        unsigned char* dp = out + (size_t)(dst0 + r) * row_bytes + (size_t)part * unit;
        if (a16) *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<const uint4*>(sp);
        else if (a4) *reinterpret_cast<unsigned*>(dp) = *reinterpret_cast<const unsigned*>(sp);
        else *dp = *sp;
    }
}

cudaError_t launch_compact(const unsigned char* flags, long long M, const void* rows, int elem_bytes,
                           long long ld, void* out, long long* out_index,
                           unsigned long long* block_counts, unsigned long long* total, cudaStream_t st)
{
    const long long nb = compact_nblocks(M);
    if (nb == 0) return cudaMemsetAsync(total, 0, sizeof(unsigned long long), st);
    compact_count_kernel<<<(unsigned)nb, CP_THREADS, 0, st>>>(flags, M, block_counts);
    compact_scan_kernel<<<1, 1024, 0, st>>>(block_counts, nb, total);
    compact_scatter_kernel<<<(unsigned)nb, CP_THREADS, 0, st>>>(flags, M, reinterpret_cast<const unsigned char*>(rows),
                                                                (int)(elem_bytes * ld),
                                                                reinterpret_cast<unsigned char*>(out), out_index,
                                                                block_counts);
    count_launch(3);
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

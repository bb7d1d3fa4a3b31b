// Fit-time and per-call ordering kernels (sm_100a).
//
//   skb_whiten_kernel   raw / whitened training rows -> zt64 [N][D] fp64 whitened (original order)
//   skb_pack_kernel     zt64 gathered through the kd-order permutation -> tiles [ntiles][D+2][TN] fp32
//   skb_box_kernel      per-tile bounding boxes of the packed fp32 coordinates (tile header + box table)
//   skb_qkey_kernel     Morton keys of the query rows in the KDE's prescaled space; CUB radix sort
//                       then gives the permutation that makes each warp's queries spatially compact
//
// Why: with training tiles that are kd-tree leaves and queries visited in Morton order, most
// (warp, tile) pairs are so far apart that every pair term is below 2^-126 relative to the query's
// exponent reference, i.e. flushes to exactly 0 in the fp32 arithmetic of the score kernel.  Those
// tiles are skipped with a box test (skb_score.cu) — the result is what the dense evaluation gives.
#include "skb_internal.h"
#include <cub/device/device_radix_sort.cuh>

namespace skb {

__global__ void skb_whiten_kernel(const void* __restrict__ train, int f32, long long N, int D, long long ld,
                                  int is_whitened, const KdeDev* __restrict__ kde, double* __restrict__ zt64)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    double x[SKB_MAXD];
    for (int i = 0; i < D; i++)
        x[i] = f32 ? (double)reinterpret_cast<const float*>(train)[j * ld + i]
                   : reinterpret_cast<const double*>(train)[j * ld + i];
    if (is_whitened) {
        for (int k = 0; k < D; k++) zt64[j * D + k] = x[k];
    } else {
        // z = (x - mu) . W  in fp64;  Wc = W * c  ->  z = ((x - mu) . Wc) / c
        for (int i = 0; i < D; i++) x[i] -= kde->mu[i];
        for (int k = 0; k < D; k++) {
            double a = 0.0;
            for (int i = 0; i < D; i++) a = fma(x[i], kde->Wc[i * D + k], a);
            zt64[j * D + k] = a / kde->c;
        }
    }
}

__global__ void skb_pack_kernel(const double* __restrict__ zt64, const int* __restrict__ perm, int D, double c,
                                const double* __restrict__ w, double inv_wmax, float* __restrict__ tiles, int ntiles)
{
    const int TN = tile_points(D);
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (long long)ntiles * TN) return;
    const long long t = s / TN;
    const int o = (int)(s % TN);
    float* dst = tiles + (size_t)t * stage_floats(D) + o;
    const int src = perm[s];
    if (src < 0) {
        for (int k = 0; k < D; k++) dst[k * TN] = PAD_COORD;
        dst[D * TN] = 0.f;
    } else {
        for (int k = 0; k < D; k++) dst[k * TN] = (float)(zt64[(long long)src * D + k] * c);
        dst[D * TN] = w ? (float)(w[src] * inv_wmax) : 1.0f;
    }
    dst[(D + 1) * TN] = 0.f;     // header row, filled by skb_box_kernel
}

// one warp per tile: min / max of each coordinate over the valid slots
__global__ void skb_box_kernel(float* __restrict__ tiles, const int* __restrict__ perm, int D, int ntiles,
                               float* __restrict__ boxes)
{
    const int TN = tile_points(D);
    const int t = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
    const int lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    float* tile = tiles + (size_t)t * stage_floats(D);
    for (int k = 0; k < D; k++) {
        float lo = INFINITY, hi = -INFINITY;
        for (int o = lane; o < TN; o += 32) {
            if (perm[(long long)t * TN + o] >= 0) {
                const float v = tile[k * TN + o];
                lo = fminf(lo, v); hi = fmaxf(hi, v);
            }
        }
        for (int s = 16; s > 0; s >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s));
        }
        if (lane == 0) {
            tile[(D + 1) * TN + k] = lo;
            tile[(D + 1) * TN + D + k] = hi;
            boxes[(size_t)t * 2 * D + k] = lo;
            boxes[(size_t)t * 2 * D + D + k] = hi;
        }
    }
}

// union boxes of `group` consecutive boxes: one thread per (output box, coordinate)
__global__ void skb_groupbox_kernel(const float* __restrict__ boxes, int D, int nin, int group, int nout, float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nout * D) return;
    const int s = i / D, k = i % D;
    float lo = INFINITY, hi = -INFINITY;
    for (int t = s * group; t < nin && t < (s + 1) * group; t++) {
        const float a = boxes[(size_t)t * 2 * D + k], b = boxes[(size_t)t * 2 * D + D + k];
        if (a <= b) { lo = fminf(lo, a); hi = fmaxf(hi, b); }
    }
    out[(size_t)s * 2 * D + k] = lo;
    out[(size_t)s * 2 * D + D + k] = hi;
}

cudaError_t launch_fit(const void* train, int f32, long long N, int D, long long ld, int is_whitened,
                       const KdeDev* kde_dev, double* zt64, cudaStream_t st)
{
    const int bs = 256;
    skb_whiten_kernel<<<(unsigned)((N + bs - 1) / bs), bs, 0, st>>>(train, f32, N, D, ld, is_whitened, kde_dev, zt64);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_pack(const double* zt64, const int* perm, int D, double c, const double* w, double inv_wmax,
                        float* tiles, int ntiles, float* boxes, float* sboxes, float* cboxes, cudaStream_t st)
{
    const long long total = (long long)ntiles * tile_points(D);
    const int bs = 256;
    skb_pack_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, st>>>(zt64, perm, D, c, w, inv_wmax, tiles, ntiles);
    skb_box_kernel<<<(unsigned)((ntiles + 7) / 8), 256, 0, st>>>(tiles, perm, D, ntiles, boxes);
    const int nsuper = (ntiles + SG - 1) / SG;
    skb_groupbox_kernel<<<(unsigned)((nsuper * D + 127) / 128), 128, 0, st>>>(boxes, D, ntiles, SG, nsuper, sboxes);
    const int nchunk = (nsuper + CG - 1) / CG;
    skb_groupbox_kernel<<<(unsigned)((nchunk * D + 127) / 128), 128, 0, st>>>(sboxes, D, nsuper, CG, nchunk, cboxes);
    count_launch(4);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Query ordering
// ---------------------------------------------------------------------------------------------
__global__ void skb_qkey_kernel(const KdeDev* __restrict__ kde, const void* __restrict__ q, int q_f32, long long M,
                                long long ldq, int q_is_whitened, const signed char* __restrict__ cols /*device*/,
                                unsigned long long* __restrict__ keys, int* __restrict__ idx)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const int D = kde->D;
    double x[SKB_MAXD];
    for (int i = 0; i < D; i++) {
        const int c = cols[i];
        x[i] = q_f32 ? (double)reinterpret_cast<const float*>(q)[r * ldq + c]
                     : reinterpret_cast<const double*>(q)[r * ldq + c];
    }
    float z[SKB_MAXD];
    if (q_is_whitened) {
        for (int k = 0; k < D; k++) z[k] = (float)(x[k] * kde->c);
    } else {
        for (int i = 0; i < D; i++) x[i] -= kde->mu[i];
        for (int k = 0; k < D; k++) {
            double a = 0.0;
            for (int i = 0; i < D; i++) a = fma(x[i], kde->Wc[i * D + k], a);
            z[k] = (float)a;
        }
    }
    // Morton interleave of the leading min(D, 8) coordinates, 63 / nd bits each
    const int nd = D < 8 ? D : 8;
    const int bits = (63 / nd) < 20 ? (63 / nd) : 20;
    const float span = (float)((1u << bits) - 1);
    unsigned cell[8];
    for (int k = 0; k < nd; k++) {
        const float lo = kde->glo[k], hi = kde->ghi[k];
        float f = (z[k] - lo) / fmaxf(hi - lo, 1e-30f);
        f = fminf(fmaxf(f, 0.f), 1.f);
        cell[k] = (unsigned)(f * span);
    }
    unsigned long long key = 0;
    for (int b = bits - 1; b >= 0; b--)
        for (int k = 0; k < nd; k++) key = (key << 1) | ((cell[k] >> b) & 1u);
    keys[r] = key;
    idx[r] = (int)r;
}

size_t sort_temp_bytes(long long M)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, (int)M);
    return bytes;
}

// keys_in/idx_in are scratch; the sorted row ids land in idx_out
cudaError_t launch_query_order(const KdeDev* kde, const void* q, int q_f32, long long M, long long ldq,
                               int q_is_whitened, const signed char* cols_dev, unsigned long long* keys_in,
                               unsigned long long* keys_out, int* idx_in, int* idx_out, void* temp, size_t temp_bytes,
                               cudaStream_t st)
{
    const int bs = 256;
    skb_qkey_kernel<<<(unsigned)((M + bs - 1) / bs), bs, 0, st>>>(kde, q, q_f32, M, ldq, q_is_whitened, cols_dev,
                                                                 keys_in, idx_in);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, idx_out, (int)M, 0, 64, st);
    count_launch(4);
    return e;
}

}  // namespace skb

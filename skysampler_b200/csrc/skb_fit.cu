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
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <vector>
#include <cstring>

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

// D == 1: far-field record of every tile.  With prescaled coordinates a pair term is 2^-(q - x)^2; about the tile's centre c
// (s = q - c, t = x - c):  2^-(s - t)^2 = 2^-(s^2) * e^(2 ln2 s t) * 2^-(t^2), so the whole tile contributes
//     2^-(s^2) * sum_n B_n s^n,   B_n = (2 ln2)^n / n! * sum_j w_j t_j^n 2^-(t_j^2)
// and the series is cut after n = 5 when 2 ln2 |s| r <= 1/4 (r = the tile's half width): relative remainder < 6e-7 of the
// tile's own (positive) contribution.  One warp per tile, fp64 sums.
__global__ void skb_moment_kernel(const float* __restrict__ tiles, const float* __restrict__ boxes, int ntiles,
                                  float* __restrict__ mom)
{
    constexpr int TN = tile_points(1);
    const int t = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
    const int lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    const float* tile = tiles + (size_t)t * stage_floats(1);
    const float lo = boxes[2 * t], hi = boxes[2 * t + 1];
    const float c = 0.5f * lo + 0.5f * hi;
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int o = lane; o < TN; o += 32) {
        const double w = (double)tile[TN + o];
        if (!(w > 0.0)) continue;
        const double tt = (double)tile[o] - (double)c;
        double term = w * exp2(-tt * tt);
#pragma unroll
        for (int n = 0; n < 6; n++) { a[n] += term; term *= tt; }
    }
#pragma unroll
    for (int n = 0; n < 6; n++)
        for (int s = 16; s > 0; s >>= 1) a[n] += __shfl_xor_sync(0xffffffffu, a[n], s);
    if (lane == 0) {
        float* r = mom + (size_t)t * 12;
        const bool valid = lo <= hi;
        r[0] = lo; r[1] = hi; r[2] = valid ? c : 0.f;
        r[3] = valid ? fmaxf(hi - c, c - lo) * 1.000001f : 0.f;
        const double L2 = 2.0 * 0.693147180559945309417232121458;
        double f = 1.0;
        for (int n = 0; n < 6; n++) {
            r[4 + n] = valid ? (float)(a[n] * f) : 0.f;
            f *= L2 / (double)(n + 1);
        }
        r[10] = 0.f; r[11] = 0.f;
    }
}

// D == 2: far-field records (skb_internal.h: FF2_*).  Per tile, about the box centre:
//     B_ij = (2 ln2)^(i+j) / (i! j!) * sum_p w_p tx_p^i ty_p^j 2^-(tx_p^2 + ty_p^2),   i + j <= FF2_P,
// written into the free part of the tile's header row; and the same for each of the tile's four CELLS — the leaf split at
// the median of its wider coordinate, each half at the median of the other coordinate (~32 points, half the radius) — into
// the side table `cells` [tile][4][FF2_CELL].  Cells are sets of points, not ranges: the pair loop does not care about the
// order inside a tile, so nothing is re-ordered.  One warp per tile, lane l owns the coefficients l, l + 32, l + 64; fp64 sums.
__device__ void ff2_record(const float* sx, const float* sy, const float* sw, const unsigned char* cid, int want, int lane,
                           float* __restrict__ out /* cx, cy, rx, ry, B[] */)
{
    constexpr int TN = tile_points(2);
    float lox = INFINITY, loy = INFINITY, hix = -INFINITY, hiy = -INFINITY;
    for (int o = lane; o < TN; o += 32)
        if (sw[o] > 0.f && (want < 0 || cid[o] == want)) {
            lox = fminf(lox, sx[o]); hix = fmaxf(hix, sx[o]);
            loy = fminf(loy, sy[o]); hiy = fmaxf(hiy, sy[o]);
        }
    for (int s = 16; s > 0; s >>= 1) {
        lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, s)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, s));
        loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, s)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, s));
    }
    const bool valid = (lox <= hix) && (loy <= hiy);
    const float cx = valid ? 0.5f * lox + 0.5f * hix : 0.f, cy = valid ? 0.5f * loy + 0.5f * hiy : 0.f;
    const double L2 = 2.0 * 0.693147180559945309417232121458;
    for (int idx = lane; idx < FF2_NCOEF; idx += 32) {
        int i = 0, off = 0;                                   // row i starts at off = i (P + 1) - i (i - 1) / 2
        while (off + (FF2_P + 1 - i) <= idx) { off += FF2_P + 1 - i; i++; }
        const int j = idx - off;
        double a = 0.0;
        for (int o = 0; o < TN; o++) {
            const double w = (double)sw[o];
            if (!(w > 0.0) || !(want < 0 || cid[o] == want)) continue;
            const double tx = (double)sx[o] - (double)cx, ty = (double)sy[o] - (double)cy;
            double term = w * exp2(-(tx * tx + ty * ty));
            for (int k = 0; k < i; k++) term *= tx;
            for (int k = 0; k < j; k++) term *= ty;
            a += term;
        }
        double f = 1.0;
        for (int k = 1; k <= i; k++) f *= L2 / (double)k;
        for (int k = 1; k <= j; k++) f *= L2 / (double)k;
        out[4 + idx] = valid ? (float)(a * f) : 0.f;
    }
    if (lane == 0) {
        out[0] = cx;
        out[1] = cy;
        out[2] = valid ? fmaxf(hix - cx, cx - lox) * 1.000001f : -1.f;
        out[3] = valid ? fmaxf(hiy - cy, cy - loy) * 1.000001f : -1.f;
    }
}

__global__ void __launch_bounds__(256) skb_moment2_kernel(float* __restrict__ tiles, const float* __restrict__ boxes, int ntiles,
                                                          float* __restrict__ cells)
{
    constexpr int TN = tile_points(2);
    __shared__ float s_x[8][TN], s_y[8][TN], s_w[8][TN];
    __shared__ unsigned char s_c[8][TN];
    const int wid = threadIdx.x / 32;
    const int t = blockIdx.x * 8 + wid;
    const int lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    float* tile = tiles + (size_t)t * stage_floats(2);
    float* sx = s_x[wid]; float* sy = s_y[wid]; float* sw = s_w[wid];
    unsigned char* cid = s_c[wid];
    for (int o = lane; o < TN; o += 32) { sx[o] = tile[o]; sy[o] = tile[TN + o]; sw[o] = tile[2 * TN + o]; cid[o] = 255; }
    __syncwarp();
    ff2_record(sx, sy, sw, cid, -1, lane, tile + 3 * TN + FF2_OFF);          // the tile's own record, in its header row
    if (!cells) return;
    // cells: median split along the wider box coordinate, then each half along the other one (ranks by counting)
    const bool xfirst = (boxes[4 * t + 2] - boxes[4 * t]) >= (boxes[4 * t + 3] - boxes[4 * t + 1]);
    const float* pa = xfirst ? sx : sy;
    const float* pb = xfirst ? sy : sx;
    int nv = 0;
    for (int o = 0; o < TN; o++) nv += (sw[o] > 0.f);
    for (int o = lane; o < TN; o += 32) {
        if (!(sw[o] > 0.f)) continue;
        int r = 0;
        for (int k = 0; k < TN; k++) r += (sw[k] > 0.f) && (pa[k] < pa[o] || (pa[k] == pa[o] && k < o));
        cid[o] = (unsigned char)(r >= (nv + 1) / 2 ? 2 : 0);
    }
    __syncwarp();
    int nh[2] = {0, 0};
    for (int o = 0; o < TN; o++) if (cid[o] != 255) nh[cid[o] >> 1]++;
    unsigned char quarter[TN / 32];
    for (int o = lane, u = 0; o < TN; o += 32, u++) {
        quarter[u] = 255;
        if (cid[o] == 255) continue;
        const int h = cid[o] >> 1;
        int r = 0;
        for (int k = 0; k < TN; k++) r += (cid[k] != 255) && ((cid[k] >> 1) == h) && (pb[k] < pb[o] || (pb[k] == pb[o] && k < o));
        quarter[u] = (unsigned char)(2 * h + (r >= (nh[h] + 1) / 2 ? 1 : 0));
    }
    __syncwarp();
    for (int o = lane, u = 0; o < TN; o += 32, u++) cid[o] = quarter[u];
    __syncwarp();
    for (int c = 0; c < 4; c++) {
        float* rec = cells + ((size_t)t * 4 + c) * FF2_CELL;
        ff2_record(sx, sy, sw, cid, c, lane, rec);
        if (lane == 0) rec[FF2_CELL - 1] = 0.f;
    }
}

// mean leaf radius rx + ry (prescaled units) over the tiles that hold points -> KdeDev::ff2_cells_on
__global__ void __launch_bounds__(256) skb_ff2_decide_kernel(const float* __restrict__ tiles, int ntiles, KdeDev* __restrict__ kde)
{
    constexpr int TN = tile_points(2);
    __shared__ double s_sum[256];
    __shared__ int s_cnt[256];
    double sum = 0.0;
    int cnt = 0;
    for (int t = threadIdx.x; t < ntiles; t += 256) {
        const float* hdr = tiles + (size_t)t * stage_floats(2) + 3 * TN + FF2_OFF;
        if (hdr[2] >= 0.f) { sum += (double)hdr[2] + (double)hdr[3]; cnt++; }
    }
    s_sum[threadIdx.x] = sum; s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { s_sum[threadIdx.x] += s_sum[threadIdx.x + s]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) kde->ff2_cells_on = (s_cnt[0] > 0 && s_sum[0] / s_cnt[0] <= (double)FF2_CELLS_MAX_R) ? 1 : 0;
}

cudaError_t launch_moments2(float* tiles, const float* boxes, int ntiles, float* cells, KdeDev* kde_dev, cudaStream_t st)
{
    if (ntiles <= 0) return cudaSuccess;
    skb_moment2_kernel<<<(unsigned)((ntiles + 7) / 8), 256, 0, st>>>(tiles, boxes, ntiles, cells);
    skb_ff2_decide_kernel<<<1, 256, 0, st>>>(tiles, ntiles, kde_dev);
    count_launch(2);
    return cudaGetLastError();
}

cudaError_t launch_moments(const float* tiles, const float* boxes, int ntiles, float* mom, cudaStream_t st)
{
    if (!mom || ntiles <= 0) return cudaSuccess;
    skb_moment_kernel<<<(unsigned)((ntiles + 7) / 8), 256, 0, st>>>(tiles, boxes, ntiles, mom);
    count_launch();
    return cudaGetLastError();
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
// kd ordering of the training points ON THE DEVICE (fit time).
//
// The tiles of a fitted KDE are the leaves of a kd-tree: recursive median split along the widest
// coordinate, the left part always a whole number of tiles (of super-boxes / chunks higher up), so that
// every aligned run of SG tiles / CG super-boxes is a subtree with a tight union box.  The split SIZES
// depend on the point count only, so the host lays out every level's node table up front and the device
// runs, per level and without any host synchronisation:
//   kd_bounds_kernel   per-node bounding boxes (warp-aggregated atomic min / max on order-preserving uints)
//   kd_key_kernel      widest coordinate per node -> 64-bit key (node id << 32 | ordered coordinate)
//   CUB radix sort     (key, row id) pairs: every node's rows end up sorted along its split coordinate,
//                      the first `nl` of them are the left child
//   kd_relabel_kernel  heap-numbered child ids for the next level
// ~15 levels x (4 small kernels + one N-pair sort) = a few milliseconds at N = 1e6.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ord_f32(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// want_flags: out[i] = (w[i] > 0);  else out[i] = i
__global__ void kd_init_kernel(const double* __restrict__ w, long long N, int* __restrict__ out, int want_flags)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    out[i] = want_flags ? ((w[i] > 0.0) ? 1 : 0) : (int)i;
}

__global__ void kd_fill_kernel(int* __restrict__ p, long long n, int v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// bounds[node][0][k] = ordered min (starts at ~0), bounds[node][1][k] = ordered max (starts at 0)
__global__ void kd_boundsinit_kernel(unsigned* __restrict__ bounds, long long n, int D)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bounds[i] = ((i / D) & 1) ? 0u : 0xffffffffu;
}

__global__ void kd_bounds_kernel(const double* __restrict__ zt64, const int* __restrict__ idx, const int* __restrict__ nodeid,
                                 long long nv, int D, double c, int base, unsigned* __restrict__ bounds)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = p < nv;
    const int node = on ? nodeid[p] : -1;
    const int node0 = __shfl_sync(0xffffffffu, node, 0);
    const bool uniform = __all_sync(0xffffffffu, node == node0);
    const double* row = zt64 + (long long)(on ? idx[p] : 0) * D;
    for (int k = 0; k < D; k++) {
        const unsigned o = on ? ord_f32((float)(row[k] * c)) : 0u;
        if (uniform) {                                   // whole warp in one node (top levels): one atomic pair per warp
            unsigned lo = on ? o : 0xffffffffu, hi = o;
            for (int s = 16; s > 0; s >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
            }
            if ((threadIdx.x & 31) == 0 && node0 >= 0) {
                atomicMin(&bounds[(size_t)(node0 - base) * 2 * D + k], lo);
                atomicMax(&bounds[(size_t)(node0 - base) * 2 * D + D + k], hi);
            }
        } else if (on) {
            atomicMin(&bounds[(size_t)(node - base) * 2 * D + k], o);
            atomicMax(&bounds[(size_t)(node - base) * 2 * D + D + k], o);
        }
    }
}

__global__ void kd_key_kernel(const double* __restrict__ zt64, const int* __restrict__ idx, const int* __restrict__ nodeid,
                              long long nv, int D, double c, int base, const unsigned* __restrict__ bounds,
                              unsigned long long* __restrict__ keys)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nv) return;
    const int node = nodeid[p];
    const unsigned* b = bounds + (size_t)(node - base) * 2 * D;
    int kbest = 0;
    float wbest = unord_f32(b[D]) - unord_f32(b[0]);
    for (int k = 1; k < D; k++) {
        const float wk = unord_f32(b[D + k]) - unord_f32(b[k]);
        if (wk > wbest) { wbest = wk; kbest = k; }
    }
    const float v = (float)(zt64[(long long)idx[p] * D + kbest] * c);
    keys[p] = ((unsigned long long)(unsigned)node << 32) | ord_f32(v);
}

// tbl[(node - base) * 2] = first position of the node, [.. + 1] = rows of its left child
__global__ void kd_relabel_kernel(const unsigned long long* __restrict__ keys_sorted, long long nv, int base,
                                  const long long* __restrict__ tbl, int* __restrict__ nodeid)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nv) return;
    const int node = (int)(keys_sorted[p] >> 32);
    const long long begin = tbl[(size_t)(node - base) * 2], nl = tbl[(size_t)(node - base) * 2 + 1];
    nodeid[p] = 2 * node + ((p - begin) >= nl ? 1 : 0);
}

__global__ void kd_finish_kernel(const int* __restrict__ idx, long long nv, long long npad, int* __restrict__ perm,
                                 const unsigned* __restrict__ root_bounds, int D, KdeDev* __restrict__ kde)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npad) perm[p] = p < nv ? idx[p] : -1;
    if (p < D) {
        kde->glo[p] = unord_f32(root_bounds[p]);
        kde->ghi[p] = unord_f32(root_bounds[D + p]);
    }
}

// perm [ceil(nv / TN) * TN] <- kd order of the rows with positive weight (all rows when w == NULL), -1 padded;
// the global box of the packed coordinates is written straight into the device descriptor.  nv = number of
// such rows (counted by the caller on its host copy of the weights).  Everything is stream-ordered; the only
// host work is the level table.
cudaError_t launch_kd_order(const double* zt64, const double* w, long long N, long long nv, int D, double c,
                            int* perm, KdeDev* kde_dev, cudaStream_t st)
{
    const long long TN = tile_points(D);
    const long long npad = ((nv + TN - 1) / TN) * TN;
    const int bs = 256;
    if (nv <= 0 || nv > N || N >= (1LL << 31)) return cudaErrorInvalidValue;
    // ---- host: node table of every level (sizes depend on nv only) ----
    struct Node { long long begin, n; int id; };
    struct Level { size_t tbl_off; int base; int nnodes; bool splits; };
    std::vector<Node> cur{{0, nv, 1}}, nxt;
    std::vector<long long> tbl;
    std::vector<Level> levels;
    for (int L = 0;; L++) {
        if (L >= 30) return cudaErrorInvalidValue;               // node ids are 31-bit
        Level lv{tbl.size(), 1 << L, 1 << L, false};
        tbl.resize(tbl.size() + 2 * (size_t)lv.nnodes, 0);
        nxt.clear();
        for (const Node& nd : cur) {
            long long nl = nd.n;
            if (nd.n > TN) {
                const long long n = nd.n;
                const long long unit = n > (long long)CG * SG * TN ? (long long)CG * SG * TN
                                       : (n > (long long)SG * TN ? (long long)SG * TN : TN);
                const long long nunit = (n + unit - 1) / unit;
                nl = ((nunit + 1) / 2) * unit;                    // 0 < nl < n because nunit >= 2
                lv.splits = true;
                nxt.push_back({nd.begin, nl, 2 * nd.id});
                nxt.push_back({nd.begin + nl, n - nl, 2 * nd.id + 1});
            } else {
                nxt.push_back({nd.begin, nd.n, 2 * nd.id});
            }
            tbl[lv.tbl_off + 2 * (size_t)(nd.id - lv.base)] = nd.begin;
            tbl[lv.tbl_off + 2 * (size_t)(nd.id - lv.base) + 1] = nl;
        }
        levels.push_back(lv);
        if (!lv.splits) break;
        cur.swap(nxt);
    }
    // ---- device scratch from the stream-ordered pool ----
    size_t sort_bytes = 0, sel_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)nullptr,
                                                    (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, (int)nv);
    if (e != cudaSuccess) return e;
    e = cub::DeviceSelect::Flagged(nullptr, sel_bytes, thrust::counting_iterator<int>(0), (const int*)nullptr, (int*)nullptr,
                                   (int*)nullptr, (int)N);
    if (e != cudaSuccess) return e;
    const size_t tmp_bytes = sort_bytes > sel_bytes ? sort_bytes : sel_bytes;
    const size_t max_nodes = (size_t)levels.back().nnodes;
    size_t off = 0;
    auto carve = [&off](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_keys0 = carve((size_t)nv * 8), o_keys1 = carve((size_t)nv * 8);
    const size_t o_idx0 = carve((size_t)N * 4), o_idx1 = carve((size_t)N * 4), o_node = carve((size_t)nv * 4);
    const size_t o_flags = carve((size_t)N * 4), o_cnt = carve(16), o_root = carve(2 * SKB_MAXD * 4);
    const size_t o_bounds = carve(max_nodes * 2 * D * 4), o_tbl = carve(tbl.size() * 8), o_tmp = carve(tmp_bytes);
    char* pool = nullptr;
    if ((e = cudaMallocAsync(reinterpret_cast<void**>(&pool), off, st)) != cudaSuccess) return e;
    auto done = [&](cudaError_t code) { cudaFreeAsync(pool, st); return code; };
    unsigned long long* keys0 = reinterpret_cast<unsigned long long*>(pool + o_keys0);
    unsigned long long* keys1 = reinterpret_cast<unsigned long long*>(pool + o_keys1);
    int* idx_cur = reinterpret_cast<int*>(pool + o_idx0);
    int* idx_alt = reinterpret_cast<int*>(pool + o_idx1);
    int* nodeid = reinterpret_cast<int*>(pool + o_node);
    int* flags = reinterpret_cast<int*>(pool + o_flags);
    unsigned* root = reinterpret_cast<unsigned*>(pool + o_root);
    unsigned* bounds = reinterpret_cast<unsigned*>(pool + o_bounds);
    long long* dtbl = reinterpret_cast<long long*>(pool + o_tbl);
    void* tmp = pool + o_tmp;
    // pageable source: the runtime stages the table before cudaMemcpyAsync returns, so `tbl` may go out of scope
    if ((e = cudaMemcpyAsync(dtbl, tbl.data(), tbl.size() * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return done(e);
    const unsigned gridN = (unsigned)((N + bs - 1) / bs), gridV = (unsigned)((nv + bs - 1) / bs);
    if (w) {                                             // rows with positive weight, original order
        kd_init_kernel<<<gridN, bs, 0, st>>>(w, N, flags, 1);
        size_t tb = tmp_bytes;
        if ((e = cub::DeviceSelect::Flagged(tmp, tb, thrust::counting_iterator<int>(0), flags, idx_cur,
                                            reinterpret_cast<int*>(pool + o_cnt), (int)N, st)) != cudaSuccess) return done(e);
        count_launch(3);
    } else {
        kd_init_kernel<<<gridN, bs, 0, st>>>(nullptr, N, idx_cur, 0);
        count_launch();
    }
    kd_fill_kernel<<<gridV, bs, 0, st>>>(nodeid, nv, 1);
    count_launch();
    for (size_t L = 0; L < levels.size(); L++) {
        const Level& lv = levels[L];
        if (L > 0 && !lv.splits) break;
        unsigned* b = L == 0 ? root : bounds;            // the root box is kept: it is the KDE's global box
        const long long nb = (long long)lv.nnodes * 2 * D;
        kd_boundsinit_kernel<<<(unsigned)((nb + bs - 1) / bs), bs, 0, st>>>(b, nb, D);
        kd_bounds_kernel<<<gridV, bs, 0, st>>>(zt64, idx_cur, nodeid, nv, D, c, lv.base, b);
        count_launch(2);
        if (!lv.splits) break;
        kd_key_kernel<<<gridV, bs, 0, st>>>(zt64, idx_cur, nodeid, nv, D, c, lv.base, b, keys0);
        size_t tb = tmp_bytes;
        if ((e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys0, keys1, idx_cur, idx_alt, (int)nv, 0, 32 + (int)L + 1,
                                                 st)) != cudaSuccess) return done(e);
        kd_relabel_kernel<<<gridV, bs, 0, st>>>(keys1, nv, lv.base, dtbl + lv.tbl_off, nodeid);
        count_launch(6);
        int* t = idx_cur; idx_cur = idx_alt; idx_alt = t;
        if ((e = cudaGetLastError()) != cudaSuccess) return done(e);
    }
    kd_finish_kernel<<<(unsigned)((npad + bs - 1) / bs), bs, 0, st>>>(idx_cur, nv, npad, perm, root, D, kde_dev);
    count_launch();
    return done(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// Feature construction (BaseContainer.construct_features, emulator.py:74-132) — elementwise, HBM bound
// ---------------------------------------------------------------------------------------------
struct FeatureProgram {
    skb_feature_spec spec[SKB_MAXD];
    int n;
};

template <class T>
__device__ __forceinline__ T feature_op(int op, T a, T b)
{
    switch (op) {
        case 1: return a - b;
        case 2: return a + b;
        case 3: return a * b;
        case 4: return a / b;
        default: return a;
    }
}
// np.sqrt(a**2. + b**2.): two roundings for the squares and one for the sum — no FMA contraction
__device__ __forceinline__ float feature_sqsum(float a, float b) { return sqrtf(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b))); }
__device__ __forceinline__ double feature_sqsum(double a, double b) { return sqrt(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b))); }

__global__ void __launch_bounds__(256)
skb_features_kernel(const __grid_constant__ FeatureProgram P, long long n_rows, double* __restrict__ out,
                    unsigned char* __restrict__ keep)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    bool ok = true;
    for (int f = 0; f < P.n; f++) {
        const skb_feature_spec& s = P.spec[f];
        const bool a_arr = s.a != nullptr, b_arr = s.b != nullptr && s.op != 0;
        // numpy promotion: float32 arithmetic only when every ARRAY operand is float32 (python literals are weak)
        const bool f32 = (!a_arr || s.a_dtype == SKB_F32) && (!b_arr || s.b_dtype == SKB_F32) && (a_arr || b_arr);
        double v;
        if (f32) {
            const float a = a_arr ? reinterpret_cast<const float*>(s.a)[r * s.a_stride] : (float)s.a_const;
            const float b = b_arr ? reinterpret_cast<const float*>(s.b)[r * s.b_stride] : (float)s.b_const;
            v = (double)(s.op == 5 ? feature_sqsum(a, b) : feature_op<float>(s.op, a, b));
        } else {
            const double a = !a_arr ? s.a_const
                                    : (s.a_dtype == SKB_F32 ? (double)reinterpret_cast<const float*>(s.a)[r * s.a_stride]
                                                            : reinterpret_cast<const double*>(s.a)[r * s.a_stride]);
            const double b = !b_arr ? s.b_const
                                    : (s.b_dtype == SKB_F32 ? (double)reinterpret_cast<const float*>(s.b)[r * s.b_stride]
                                                            : reinterpret_cast<const double*>(s.b)[r * s.b_stride]);
            v = s.op == 5 ? feature_sqsum(a, b) : feature_op<double>(s.op, a, b);
        }
        ok = ok && (v > s.lo) && (v < s.hi);                // open interval; NaN fails both, as in the reference
        out[r * P.n + f] = s.log10 ? log10(v) : v;
    }
    keep[r] = ok ? 1 : 0;
}

cudaError_t launch_features(const skb_feature_spec* specs, int n_feat, long long n_rows, double* out, unsigned char* keep,
                            cudaStream_t st)
{
    if (n_rows <= 0) return cudaSuccess;
    FeatureProgram P;
    memset(&P, 0, sizeof(P));
    P.n = n_feat;
    for (int f = 0; f < n_feat; f++) P.spec[f] = specs[f];
    skb_features_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(P, n_rows, out, keep);
    count_launch();
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

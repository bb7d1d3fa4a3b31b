"""numpy restatement of the DEVICE arithmetic of the score kernels (test infrastructure, CPU only).

Not the reference's algorithm (that is brute64 / sklearn) but the fp32 scheme skysampler_b200/csrc/skb_score.cu
uses to reach it, written out so that its error budget and the exactness of tile culling can be checked
without a GPU:

  * coordinates are prescaled by c = sqrt(log2(e)/2)/h and rounded to fp32; acc_j = sum_k (zq_k - x_jk)^2
    is an fp32 FMA chain seeded with -m, so a pair term is 2^(m - acc_j) * w_j;
  * m = floor(acc of the nearest training point) - 90 (REF_BIAS): the nearest term is ~2^-90 and every term
    more than ~36 binades farther flushes to exactly zero (ftz below 2^-126);
  * per 64-point tile every "lane" adds its two terms in fp32, lane partials accumulate in fp32 over at most
    2^8 tile indices and are then folded into an fp64 sum; log p = log_norm + log S - m ln 2;
  * a tile whose bounding box is more than 128 binades beyond the reference contributes exactly nothing, so
    skipping it (culling) cannot change the result.

`score()` returns (log p, number of tiles that were skipped).  Slow (pure numpy, one query at a time): use on a
few hundred queries x a few thousand training points.
"""
import numpy as np

REF_BIAS = 90.0
CULL_MARGIN = 128.0
TILE = 64
FOLD_TILES = 256
LN2 = float(np.log(2.0))


def _fma32(a, b, c):
    """fp32 fused multiply-add: the product of two fp32 numbers is exact in fp64, one rounding to fp32."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _exp2_ftz(x):
    """2^x in fp32 with results below 2^-126 flushed to zero (MUFU.EX2 .ftz)."""
    y = np.exp2(x.astype(np.float64))
    y = np.where(y < 2.0 ** -126, 0.0, y)
    return y.astype(np.float32)


def score(train, weights, h, query, cull=True):
    """log p of each query row under the kernels' fp32 scheme; `train` rows are taken in the given order and cut
    into 64-point tiles (the kernels use kd-tree order; the arithmetic does not care)."""
    train = np.ascontiguousarray(train, dtype=np.float64)
    query = np.ascontiguousarray(query, dtype=np.float64)
    n, d = train.shape
    w = np.ones(n) if weights is None else np.asarray(weights, dtype=np.float64)
    keep = w > 0
    train, w = train[keep], w[keep]
    n = len(train)
    wmax = w.max()
    c = np.sqrt(np.log2(np.e) / 2.0) / h
    x32 = (train * c).astype(np.float32)
    w32 = (w / wmax).astype(np.float32)
    ntiles = (n + TILE - 1) // TILE
    pad = ntiles * TILE - n
    if pad:
        x32 = np.vstack([x32, np.full((pad, d), 1.0e18, np.float32)])
        w32 = np.concatenate([w32, np.zeros(pad, np.float32)])
    xt = x32.reshape(ntiles, TILE, d)
    wt = w32.reshape(ntiles, TILE)
    lo = np.where(wt[:, :, None] > 0, xt, np.inf).min(1)
    hi = np.where(wt[:, :, None] > 0, xt, -np.inf).max(1)
    log_norm = (-0.5 * d * np.log(2 * np.pi) - d * np.log(h)) - np.log(w.sum()) + np.log(wmax)

    out = np.empty(len(query))
    skipped = 0
    for iq, row in enumerate(query):
        zq = (row * c).astype(np.float32)
        # reference: exact nearest point (the kernels start from the home tile and recentre; same end state)
        diff = zq[None, :] - x32[:n]
        acc0 = np.zeros(n, np.float32)
        for k in range(d):
            acc0 = _fma32(diff[:, k], diff[:, k], acc0)
        m = np.float32(np.floor(acc0.min()) - REF_BIAS)
        # box lower bounds, fp32 FMA chain like the kernels
        gap = np.maximum(np.maximum(lo - zq[None, :], zq[None, :] - hi), 0).astype(np.float32)
        lb = np.zeros(ntiles, np.float32)
        for k in range(d):
            lb = _fma32(gap[:, k], gap[:, k], lb)
        S = 0.0
        part = np.zeros(TILE // 2, np.float32)
        for t in range(ntiles):
            if t % FOLD_TILES == 0 and t:
                S += float(part.astype(np.float64).sum()); part[:] = 0
            if cull and (lb[t] - m > CULL_MARGIN):
                skipped += 1
                continue
            dz = zq[None, :] - xt[t]
            acc = np.full(TILE, -m, np.float32)
            for k in range(d):
                acc = _fma32(dz[:, k], dz[:, k], acc)
            e = _exp2_ftz(-acc)
            # lane l holds points 2l, 2l+1:  t = fma(w1, e1, w0 * e0)
            t0 = (wt[t, 0::2] * e[0::2]).astype(np.float32)
            term = _fma32(wt[t, 1::2], e[1::2], t0)
            part = (part + term).astype(np.float32)
        S += float(part.astype(np.float64).sum())
        with np.errstate(divide="ignore"):
            out[iq] = log_norm + np.log(S) - float(m) * LN2
    return out, skipped

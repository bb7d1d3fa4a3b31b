"""fp64 brute-force Gaussian-KDE log-density — ground truth in the tails (test infrastructure).

Follows the definition sklearn's KD-tree implements (and sklearn's own test-suite pins with
``compute_kernel_slow``, ``sklearn/neighbors/tests/test_kde.py:16-37``):

    log K(d)   = -0.5 d^2 / h^2                              (_binary_tree.pxi.tp:376-378)
    log_knorm  = -0.5 D log(2 pi) - D log(h)                 (_binary_tree.pxi.tp:448-478)
    log p(z)   = log_knorm + log sum_j w_j K(|z - z_j|) - log(sum_j w_j)
                                                             (_binary_tree.pxi.tp:1527-1529, _kde.py:273-287)

Direct differences, fp64 everywhere, max-shifted log-sum-exp.  Unlike sklearn's tree walk
(`logsubexp` cancellation, SURVEY.md finding 5) this is accurate arbitrarily far into the tails.
"""
import numpy as np

LOG_2PI = float(np.log(2.0 * np.pi))


def log_knorm(ndim, h):
    """Gaussian kernel normalisation, sklearn ``_log_kernel_norm`` (_binary_tree.pxi.tp:448-478)."""
    return -0.5 * ndim * LOG_2PI - ndim * float(np.log(h))


def partial_lse(train, weights, h, query, block=2048):
    """Per-query ``(m, s)`` with ``sum_j w_j exp(-|z-z_j|^2/2h^2) = s * exp(m)`` (natural log units).

    ``m`` is the largest single exponent ``max_j (log w_j - 0.5 d_j^2/h^2)`` over w_j > 0.
    Used directly by the training-sharded combine tests (additivity over shards).
    """
    train = np.ascontiguousarray(train, dtype=np.float64)
    query = np.ascontiguousarray(query, dtype=np.float64)
    n, d = train.shape
    if weights is None:
        logw = np.zeros(n)
    else:
        w = np.asarray(weights, dtype=np.float64)
        with np.errstate(divide="ignore"):
            logw = np.log(w)
    m_out = np.empty(len(query))
    s_out = np.empty(len(query))
    inv2h2 = 0.5 / (h * h)
    for a in range(0, len(query), block):
        q = query[a:a + block]
        d2 = np.zeros((len(q), n))
        for k in range(d):
            diff = q[:, k:k + 1] - train[None, :, k]
            d2 += diff * diff
        e = logw[None, :] - d2 * inv2h2
        m = e.max(axis=1)
        m_safe = np.where(np.isfinite(m), m, 0.0)
        s = np.exp(e - m_safe[:, None]).sum(axis=1)
        m_out[a:a + block] = m
        s_out[a:a + block] = s
    return m_out, s_out


def score(train, weights, h, query, block=2048):
    """log p(query) under the weighted Gaussian KDE fitted on ``train`` (both already whitened)."""
    train = np.asarray(train, dtype=np.float64)
    n, d = train.shape
    m, s = partial_lse(train, weights, h, query, block=block)
    sum_w = float(n) if weights is None else float(np.sum(np.asarray(weights, dtype=np.float64)))
    with np.errstate(divide="ignore"):
        return log_knorm(d, h) + m + np.log(s) - np.log(sum_w)


def combine_partials(ms, ss):
    """Merge per-shard ``(m, s)`` pairs (lists of arrays) into one — the training-sharded combine.

    Mirrors what the NCCL path does: all-reduce(max) on m, rescale s, all-reduce(sum) on s.
    """
    ms = np.stack(ms)
    ss = np.stack(ss)
    m = ms.max(axis=0)
    m_safe = np.where(np.isfinite(m), m, 0.0)
    s = (ss * np.exp(np.where(np.isfinite(ms), ms, -np.inf) - m_safe[None, :])).sum(axis=0)
    return m, s

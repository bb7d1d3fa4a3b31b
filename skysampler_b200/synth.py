"""Seeded synthetic inputs of the BASELINE.json shapes — plain numpy, no I/O.

The reference ships no data and its catalogue loaders need DES files; benchmarks and tests use
these stand-ins: a K-component Gaussian mixture in d dimensions ("galaxy feature space") with
"deep" (unit) or "wide" (piecewise-constant positive) sample weights.
"""
import numpy as np
import pandas as pd


def mixture(d, n, seed, k=8, cov_scale=1.0, sample_seed=None):
    """G(d, n, seed): K=8 Gaussians, means ~ N(0, 2^2 I), cov = A A^T with A = I + 0.3 N(0,1).

    ``sample_seed`` draws a different sample from the SAME mixture (optionally with ``cov_scale``): the
    cluster / field pair of the ratio workload must overlap, like a cluster line of sight and its field."""
    rng = np.random.RandomState(seed)
    means = rng.normal(0.0, 2.0, size=(k, d))
    mats = np.eye(d)[None, :, :] + 0.3 * rng.normal(size=(k, d, d))
    pis = rng.dirichlet(np.ones(k))
    if sample_seed is not None:
        rng = np.random.RandomState(sample_seed)
    comp = rng.choice(k, size=n, p=pis)
    z = rng.normal(size=(n, d))
    x = means[comp] + np.einsum("nij,nj->ni", mats[comp], z) * np.sqrt(cov_scale)
    return np.ascontiguousarray(x)


def mixture_big(d, n, seed, k=8, cov_scale=1.0):
    """The same mixture family as ``mixture`` for multi-million-row sets: rows are generated component by
    component (no [n, d, d] temporary) and then shuffled, so memory stays at a few copies of [n, d]."""
    rng = np.random.RandomState(seed)
    means = rng.normal(0.0, 2.0, size=(k, d))
    mats = np.eye(d)[None, :, :] + 0.3 * rng.normal(size=(k, d, d))
    pis = rng.dirichlet(np.ones(k))
    counts = rng.multinomial(n, pis)
    x = np.empty((n, d))
    a = 0
    for c in range(k):
        z = rng.normal(size=(counts[c], d))
        x[a:a + counts[c]] = means[c] + (z @ mats[c].T) * np.sqrt(cov_scale)
        a += counts[c]
    rng.shuffle(x)
    return x


def wide_weights(n, seed, nblocks=100):
    """'wide' weights: U(0.5, 2) per 1% block of rows (mimics indexer.py:925 x emulator.py:198)."""
    rng = np.random.RandomState(seed)
    vals = rng.uniform(0.5, 2.0, size=nblocks)
    idx = np.minimum((np.arange(n) * nblocks) // max(n, 1), nblocks - 1)
    return vals[idx]


def column_names(d, rcol_last=True):
    names = ["F%d" % i for i in range(d)]
    if rcol_last:
        names[-1] = "LOGR"
    return names


def frame(d, n, seed, weighted=True, cov_scale=1.0):
    """(DataFrame[n x d], Series weights or None) ready for KDEContainer(raw_data, weights)."""
    x = mixture(d, n, seed, cov_scale=cov_scale)
    df = pd.DataFrame(x, columns=column_names(d))
    w = pd.Series(wide_weights(n, seed + 7919)) if weighted else None
    return df, w


def whiten_params(x, w=None):
    """Deterministic whitening (weighted mean, eigen-rotation of the weighted covariance, unit scale):
    same structure as KDEContainer.fit_pca (emulator.py:286-320) without its stochastic resample.
    Returns mean1, pca_mean, components, std2."""
    x = np.asarray(x, dtype=np.float64)
    w = np.ones(len(x)) if w is None else np.asarray(w, dtype=np.float64)
    mean1 = np.average(x, axis=0, weights=w)
    xc = x - mean1
    cov = (xc * w[:, None]).T @ xc / w.sum()
    evals, evecs = np.linalg.eigh(cov)
    order = np.argsort(evals)[::-1]
    comps = evecs[:, order].T
    rot = xc @ comps.T
    std2 = np.sqrt(np.average(rot ** 2, axis=0, weights=w))
    return mean1, np.zeros(x.shape[1]), comps, std2


def queries(train_z, h, m, seed, tail_frac=0.1, tail_scale=2.0):
    """Whitened-space queries: KDE draws from ``train_z`` (unweighted centres) with a fraction inflated
    about the origin for tail coverage."""
    rng = np.random.RandomState(seed)
    idx = rng.randint(0, len(train_z), size=m)
    q = train_z[idx] + h * rng.normal(size=(m, train_z.shape[1]))
    ntail = int(m * tail_frac)
    if ntail:
        q[:ntail] *= tail_scale
    return np.ascontiguousarray(q)

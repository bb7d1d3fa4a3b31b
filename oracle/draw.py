"""Restatement of ``KernelDensity.sample`` for the Gaussian kernel (sklearn neighbors/_kde.py:338-349)
from explicit streams.  TEST INFRASTRUCTURE only.

    u = rng.uniform(0, 1, n)
    weighted:    i = searchsorted(cumsum(w), u * cumsum(w)[-1])          (side='left')
    unweighted:  i = (u * N).astype(int64)
    sample      = rng.normal(data[i], h)   ==   data[i] + h * z,   z = the next n*d legacy normals
"""
import numpy as np


def select_index(u, n, weights=None):
    u = np.asarray(u, dtype=np.float64)
    if weights is None:
        return (u * n).astype(np.int64)
    cumsum_weight = np.cumsum(np.asarray(weights, dtype=np.float64))
    sum_weight = cumsum_weight[-1]
    return np.searchsorted(cumsum_weight, u * sum_weight)


def sample_from_streams(data, h, u, z, weights=None):
    """Whitened-space draws and their centre indices from uniform stream u[n] and normal stream z[n, d]."""
    data = np.asarray(data, dtype=np.float64)
    i = select_index(u, data.shape[0], weights)
    return data[i] + h * np.asarray(z, dtype=np.float64), i


def inverse_whiten(zs, mean1, pca_mean, components, std2):
    """pca_inverse_transform (emulator.py:329-335): (z * std2) @ C + pca.mean_ + mean1."""
    return (np.asarray(zs) * std2) @ np.asarray(components) + pca_mean + mean1


def whiten(x, mean1, pca_mean, components, std2):
    """pca_transform (emulator.py:322-327) with sklearn's PCA.transform (decomposition/_base.py:151-159)."""
    x = np.asarray(x, dtype=np.float64) - mean1
    z = x @ np.asarray(components).T - np.asarray(pca_mean) @ np.asarray(components).T
    return z / std2

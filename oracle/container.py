"""Restatement of ``skysampler.emulator.KDEContainer`` (emulator.py:234-394) on top of scikit-learn,
exactly as the reference calls it.  TEST INFRASTRUCTURE / CPU baseline only (see oracle/__init__.py).

This is the "port" that ``bench.py --impl reference`` and the ``cpu_baseline`` leg time on the host
cores: sklearn's KD-tree ``KernelDensity`` is the reference's own engine for this path
(emulator.py:354-356, 360, 372, 384), scikit-learn 1.9.0 in this image.
"""
import copy

import numpy as np
import pandas as pd
import sklearn.decomposition as decomp
import sklearn.neighbors as neighbors


def weighted_mean(values, weights):                      # emulator.py:52-54
    return np.average(values, axis=0, weights=weights)


def weighted_std(values, weights):                       # emulator.py:57-64
    average = np.average(values, axis=0, weights=weights)
    variance = np.average((values - average) ** 2, axis=0, weights=weights)
    return np.sqrt(variance)


class RefKDEContainer(object):
    """Same state and call order as the reference container; ``atol``/``rtol`` are instance
    attributes so the tests can run both the shipped setting (1e-6/1e-6) and the exact one (0/0)."""

    def __init__(self, raw_data, weights=None, seed=None, atol=1e-6, rtol=1e-6, breadth_first=False):
        self.rng = np.random.RandomState(seed) if seed is not None else np.random.RandomState()   # :245-248
        self.data = raw_data
        self.columns = raw_data.columns
        self.weights = np.ones(len(raw_data), dtype=float) if weights is None else weights.astype(float)  # :252-255
        self.ndim = self.data.shape[1]
        self._atol, self._rtol, self._breadth_first = atol, rtol, breadth_first

    # ---- whitening: emulator.py:286-340 ----
    def select_subset(self, data, weights, nsample=10000):
        indexes = np.arange(len(data))
        ww = weights / weights.sum()
        inds = self.rng.choice(indexes, size=int(nsample), p=ww, replace=True)
        return data.iloc[inds]

    def fit_pca(self):
        self.mean1 = weighted_mean(self.data, self.weights)
        _data = self.data - self.mean1
        subset = self.select_subset(_data, self.weights, nsample=100000)
        self.pca = decomp.PCA()
        self.pca.fit(subset)
        _data = self.pca.transform(_data)
        self.std2 = weighted_std(_data, self.weights)
        self._jacobian_matrix = np.dot(np.diag(1. / self.std2), self.pca.components_)
        self._jacobian_matrix_inv = np.linalg.inv(self._jacobian_matrix)
        self._jacobian_det = np.linalg.det(self._jacobian_matrix_inv)

    def set_whitening(self, mean1, pca_mean, components, std2):
        """Install fixed whitening parameters (fit_pca is stochastic)."""
        components = np.array(components, dtype=np.float64)
        self.mean1 = np.array(mean1, dtype=np.float64)
        self.std2 = np.array(std2, dtype=np.float64)
        pca = decomp.PCA()
        pca.mean_ = np.array(pca_mean, dtype=np.float64)
        pca.components_ = components
        pca.n_components_ = components.shape[0]
        pca.n_features_in_ = components.shape[1]
        pca.explained_variance_ = np.ones(components.shape[0])
        pca.whiten = False
        self.pca = pca
        self._jacobian_matrix = np.dot(np.diag(1. / self.std2), components)
        self._jacobian_matrix_inv = np.linalg.inv(self._jacobian_matrix)
        self._jacobian_det = np.linalg.det(self._jacobian_matrix_inv)
        self._data = self.pca_transform(self.data)

    def pca_transform(self, data):
        _data = data - self.mean1
        _data = self.pca.transform(_data)
        _data /= self.std2
        return _data

    def pca_inverse_transform(self, data):
        _data = data * self.std2
        _data = self.pca.inverse_transform(_data)
        _data = _data + self.mean1
        return pd.DataFrame(_data, columns=self.columns)

    def standardize_data(self):
        self.fit_pca()
        self._data = self.pca_transform(self.data)

    # ---- the hot path: emulator.py:351-385 ----
    def construct_kde(self, bandwidth):
        self.bandwidth = bandwidth
        self.kde = neighbors.KernelDensity(bandwidth=self.bandwidth, kernel="gaussian", atol=self._atol,
                                           rtol=self._rtol, breadth_first=self._breadth_first)
        self.kde.fit(self._data, sample_weight=self.weights)

    def random_draw(self, num, rmin=None, rmax=None, rcol="LOGR"):
        _res = self.kde.sample(n_samples=int(num), random_state=self.rng)
        self.res = self.pca_inverse_transform(_res)
        if (rmin is not None) or (rmax is not None):
            if rmin is None:
                rmin = self.data[rcol].min()
            if rmax is None:
                rmax = self.data[rcol].max()
            inds = np.asarray((self.res[rcol] > rmax) | (self.res[rcol] < rmin))
            while inds.sum():
                vals = self.kde.sample(n_samples=int(inds.sum()), random_state=self.rng)
                _res[inds, :] = vals
                self.res = self.pca_inverse_transform(_res)
                inds = np.asarray((self.res[rcol] > rmax) | (self.res[rcol] < rmin))
        self._res = _res
        self.res = self.pca_inverse_transform(_res)
        return self.res

    def score_samples(self, arr):
        arr = self.pca_transform(arr)
        res = self.kde.score_samples(arr)
        return res, self._jacobian_det

    def drop_kde(self):
        self.pca = None
        self.kde = None


def score_worker(args):
    """One worker of the reference's one-process-per-chunk pool (emulator.py:698-712): fit its own
    tree on the whitened rows and score its chunk.  Returns the number of rows scored."""
    z, w, h, q, atol, rtol = args
    kde = neighbors.KernelDensity(bandwidth=h, kernel="gaussian", atol=atol, rtol=rtol, breadth_first=False)
    kde.fit(z, sample_weight=w)
    out = kde.score_samples(q)
    return len(out), float(out[0])

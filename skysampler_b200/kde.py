"""The KDE container of skysampler on the B200 backend.

``KDEContainer`` mirrors ``skysampler/emulator.py:234-394`` name for name — constructor, attributes,
``standardize_data / construct_kde / random_draw / score_samples / drop_kde`` — so that
``calc_scores*`` / ``make_*_infodicts`` (emulator.py:525-712) and the ``bin/*_sampler*.py`` drivers
run unchanged against it.  What changed is underneath: ``construct_kde`` uploads the whitened
training rows to the GPU instead of building sklearn's KD-tree, ``score_samples`` evaluates every
(query, training) pair exactly (the reference's ``atol = rtol = 1e-6`` becomes 0) with the whitening
fused into the kernel, and ``random_draw`` runs the centre selection / noise / inverse whitening /
radial window on the device.  The PCA fit itself stays on the host with sklearn's ``PCA`` so the
whitening parameters are the reference's numbers.

``GaussianKDE`` is the thin object playing the role of ``sklearn.neighbors.KernelDensity`` for
this path (fit / score_samples / sample), usable with numpy arrays or torch CUDA tensors.
"""
import copy
import ctypes as C

import numpy as np
import pandas as pd

from . import _cabi
from ._cabi import check, ptr, dtype_code, stream_ptr, cols_array


def weighted_mean(values, weights):
    """emulator.py:52-54."""
    return np.average(values, axis=0, weights=weights)


def weighted_std(values, weights):
    """emulator.py:57-64."""
    average = np.average(values, axis=0, weights=weights)
    variance = np.average((values - average) ** 2, axis=0, weights=weights)
    return np.sqrt(variance)


def _as_matrix(x, allow_torch=True, check_finite=True):
    """DataFrame / ndarray / torch tensor -> C-contiguous 2-D float matrix (float64 unless float32 given)."""
    if allow_torch and _cabi._is_torch(x):
        if x.dim() == 1:
            x = x.reshape(-1, 1)
        if str(x.dtype) not in ("torch.float64", "torch.float32"):
            x = x.double()
        return x.contiguous()
    if isinstance(x, (pd.DataFrame, pd.Series)):
        x = x.values
    x = np.asarray(x)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    if x.ndim != 2:
        raise ValueError("Expected 2D array, got %dD array instead" % x.ndim)
    if x.dtype != np.float32:
        x = x.astype(np.float64, copy=False)
    if check_finite:
        # sklearn's validate_data raises the same way (KernelDensity.fit / score_samples, _kde.py:233,272) and uses the same
        # shortcut (utils/validation.py: _assert_all_finite): one pass for the sum, the element-wise test only if that
        # sum is not finite (an overflowing sum of finite values is rare and merely costs the slow path)
        with np.errstate(over="ignore", invalid="ignore"):
            total = x.sum(dtype=np.float64)
        if not np.isfinite(total) and not np.isfinite(x).all():
            raise ValueError("Input X contains NaN or infinity.")
    return np.ascontiguousarray(x)


class GaussianKDE(object):
    """Gaussian KDE on one B200; the KernelDensity stand-in (sklearn _kde.py:198-349) of this path.

    Parameters mirror what the reference passes (emulator.py:354-356); ``atol``/``rtol``/
    ``breadth_first`` are accepted and ignored: every pair is evaluated.
    """

    def __init__(self, bandwidth=1.0, kernel="gaussian", atol=0, rtol=0, breadth_first=True, device=None):
        if kernel != "gaussian":
            raise ValueError("the B200 backend implements kernel='gaussian' only, got %r" % (kernel,))
        self.bandwidth = float(bandwidth)
        self.kernel = kernel
        self.atol, self.rtol, self.breadth_first = atol, rtol, breadth_first
        self.device = device
        self._h = None

    # -- lifetime ---------------------------------------------------------------------------------
    def fit(self, X, y=None, sample_weight=None, whitening=None, is_whitened=True):
        """Upload the training rows.  ``whitening = (mu, W, Winv)`` enables raw-space queries/draws."""
        self.close()
        lib = _cabi.lib()
        X = _as_matrix(X)
        n, d = X.shape
        w = None
        if sample_weight is not None:
            w = np.ascontiguousarray(np.asarray(sample_weight, dtype=np.float64))
            if w.shape != (n,):
                raise ValueError("sample_weight.shape == %s, expected (%d,)!" % (w.shape, n))
        cs = np.cumsum(w) if w is not None else None          # sklearn _kde.py:345 (sequential fp64)
        mu = W = Winv = None
        if whitening is not None:
            mu, W, Winv = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in whitening]
        dev = self.device
        if dev is None:
            dev = X.device.index if (_cabi._is_torch(X) and X.is_cuda) else _current_device()
        h = C.c_void_p()
        check(lib.skb_create(ptr(X), dtype_code(X), n, d, d, 1 if is_whitened else 0, ptr(w), ptr(cs),
                             self.bandwidth, ptr(mu), ptr(W), ptr(Winv), int(dev), C.byref(h)))
        self._h = h
        self.n_train_, self.ndim_, self.device_ = n, d, int(dev)
        self.sum_weight_ = float(lib.skb_sum_weight(h))
        self.fit_ms_ = float(lib.skb_fit_ms(h))           # device time of the fit (whitening, kd ordering, packing)
        return self

    def close(self):
        if getattr(self, "_h", None) is not None:
            _cabi.lib().skb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getstate__(self):
        raise TypeError("a fitted GaussianKDE holds device memory and cannot be pickled; "
                        "call KDEContainer.drop_kde() first (emulator.py:387-389)")

    def _handle(self):
        if self._h is None:
            raise RuntimeError("This GaussianKDE instance is not fitted yet (or was dropped).")
        return self._h

    # -- scoring ----------------------------------------------------------------------------------
    def score_samples(self, X, cols=None, is_whitened=True, out=None, stream=None):
        """log p(X): float64 ndarray for host input, torch CUDA tensor for CUDA input (zero copy)."""
        h = self._handle()
        X = _as_matrix(X)
        m, ld = X.shape
        if cols is None and ld != self.ndim_:
            raise ValueError("X has %d features, but GaussianKDE is expecting %d features as input." % (ld, self.ndim_))
        cols_a = cols_array(cols, self.ndim_, ld)
        if _cabi._is_torch(X) and X.is_cuda:
            import torch
            if out is None:
                out = torch.empty(m, dtype=torch.float64, device=X.device)
            st = stream if stream is not None else torch.cuda.current_stream(X.device)
            check(_cabi.lib().skb_score(h, ptr(X), dtype_code(X), m, ld, ptr(cols_a), 1 if is_whitened else 0,
                                        ptr(out), stream_ptr(st)))
            return out
        if _cabi._is_torch(X):
            X = X.numpy()
        if out is None:
            out = np.empty(m, dtype=np.float64)
        check(_cabi.lib().skb_score(h, ptr(X), dtype_code(X), m, ld, ptr(cols_a), 1 if is_whitened else 0,
                                    ptr(out), None))
        return out

    def score_partial(self, X, cols=None, is_whitened=True, stream=None):
        """Per-query ``(m, s)`` over this handle's rows: ``sum_j w_j K_j = s * exp(m)`` (training-sharded mode)."""
        h = self._handle()
        X = _as_matrix(X)
        m, ld = X.shape
        cols_a = cols_array(cols, self.ndim_, ld)
        if _cabi._is_torch(X) and X.is_cuda:
            import torch
            om = torch.empty(m, dtype=torch.float64, device=X.device)
            os_ = torch.empty(m, dtype=torch.float64, device=X.device)
            st = stream if stream is not None else torch.cuda.current_stream(X.device)
            check(_cabi.lib().skb_score_partial(h, ptr(X), dtype_code(X), m, ld, ptr(cols_a),
                                                1 if is_whitened else 0, ptr(om), ptr(os_), stream_ptr(st)))
            return om, os_
        if _cabi._is_torch(X):
            X = X.numpy()
        om = np.empty(m, dtype=np.float64)
        os_ = np.empty(m, dtype=np.float64)
        check(_cabi.lib().skb_score_partial(h, ptr(X), dtype_code(X), m, ld, ptr(cols_a),
                                            1 if is_whitened else 0, ptr(om), ptr(os_), None))
        return om, os_

    # -- draws ------------------------------------------------------------------------------------
    def sample(self, n_samples=1, random_state=None, raw=False, return_index=False):
        """``KernelDensity.sample`` with the reference's stream order (sklearn _kde.py:338-349):
        n uniforms, then n*d normals, both from ``random_state``; the arithmetic runs on the GPU.
        ``raw=True`` additionally applies the inverse whitening map."""
        h = self._handle()
        n = int(n_samples)
        rng = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
        u = rng.uniform(0, 1, size=n)
        z = rng.normal(0.0, 1.0, size=(n, self.ndim_))
        out_z = np.empty((n, self.ndim_), dtype=np.float64)
        out_x = np.empty((n, self.ndim_), dtype=np.float64) if raw else None
        idx = np.empty(n, dtype=np.int64) if return_index else None
        check(_cabi.lib().skb_draw(h, n, ptr(u), ptr(z), 0, 0, None, -1, 0.0, 0.0, 1, ptr(idx), ptr(out_x),
                                   self.ndim_, ptr(out_z), None, None, None))
        res = out_x if raw else out_z
        return (res, idx) if return_index else res

    def draw_streams(self, u, z, out_x, out_z, rows=None, out_idx=None):
        """Draws from caller-supplied streams (``u`` [n] uniforms, ``z`` [n, d] normals), consumed exactly as
        sklearn does (_kde.py:338-349): draw i lands in row ``rows[i]`` (default i) of ``out_x`` (raw space)
        and ``out_z`` (whitened space) — the building block of ``KDEContainer.random_draw``'s redraw loop
        (emulator.py:371-375)."""
        h = self._handle()
        n = len(u)
        check(_cabi.lib().skb_draw(h, n, ptr(u), ptr(z), 0, 0, ptr(rows), -1, 0.0, 0.0, 1, ptr(out_idx), ptr(out_x),
                                   self.ndim_, ptr(out_z), None, None, None))

    def sample_philox(self, n_samples, seed, offset=0, rcol=-1, rmin=0.0, rmax=0.0, max_attempts=1000,
                      device_out=False, stream=None):
        """Device-side draws from the Philox counter RNG, optional in-kernel radial window.
        Returns (x, flags, n_flagged); x is raw-space when the KDE was fitted with an inverse map."""
        h = self._handle()
        n = int(n_samples)
        nfl = C.c_int64(0)
        if device_out:
            import torch
            dev = torch.device("cuda", self.device_)
            x = torch.empty((n, self.ndim_), dtype=torch.float64, device=dev)
            fl = torch.empty(n, dtype=torch.uint8, device=dev)
            st = stream if stream is not None else torch.cuda.current_stream(dev)
            check(_cabi.lib().skb_draw(h, n, None, None, int(seed), int(offset), None, int(rcol), float(rmin),
                                       float(rmax), int(max_attempts), None, ptr(x), self.ndim_, None, ptr(fl),
                                       None, stream_ptr(st)))
            return x, fl, None
        x = np.empty((n, self.ndim_), dtype=np.float64)
        fl = np.empty(n, dtype=np.uint8)
        check(_cabi.lib().skb_draw(h, n, None, None, int(seed), int(offset), None, int(rcol), float(rmin),
                                   float(rmax), int(max_attempts), None, ptr(x), self.ndim_, None, ptr(fl),
                                   C.byref(nfl), None))
        return x, fl, int(nfl.value)


def _current_device():
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


class KDEContainer(object):
    """Drop-in for ``skysampler.emulator.KDEContainer`` (emulator.py:234-394) on the B200 backend."""
    _default_subset_sizes = (2000, 5000, 10000)
    _kernel = "gaussian"
    _atol = 1e-6              # kept for attribute compatibility; the backend is exact (0 / 0)
    _rtol = 1e-6
    _breadth_first = False
    _jacobian_matrix = None
    _jacobian_det = None
    # "host": draws consume self.rng exactly as the reference does (stream-for-stream replay);
    # "philox": device counter RNG seeded from self.rng, window redraws inside the kernel.
    _rng_mode = "host"
    _device = None

    def __init__(self, raw_data, weights=None, transform_params=None, seed=None):
        """emulator.py:244-256."""
        if seed is not None:
            self.rng = np.random.RandomState(seed)
        else:
            self.rng = np.random.RandomState()
        self.data = raw_data
        self.columns = raw_data.columns
        if weights is None:
            self.weights = np.ones(len(raw_data), dtype=float)
        else:
            self.weights = weights.astype(float)
        self.ndim = self.data.shape[1]
        self.pca = None
        self.kde = None

    def set_seed(self, seed):
        """emulator.py:258-259."""
        self.rng.seed(seed)

    # -- row subsampling (not a KDE draw) -----------------------------------------------------------
    def shuffle(self):
        """emulator.py:271-272."""
        self.sample(n=None, frac=1.)

    def sample(self, n=None, frac=1.):
        """emulator.py:274-284.  The reference compares ``None > len`` (a Python-2 idiom that raises on
        Python 3); the intended behaviour — n larger than the table falls back to ``frac`` — is kept."""
        inds = np.arange(len(self.data))
        tab = pd.DataFrame()
        tab["IND"] = inds
        if n is not None and n > len(tab):
            n = None
        if n is not None:
            inds = tab.sample(n=n)["IND"].values
        else:
            inds = tab.sample(frac=frac)["IND"].values
        self.data = self.data.iloc[inds].copy().reset_index(drop=True)
        if isinstance(self.weights, pd.Series):
            self.weights = self.weights.iloc[inds].copy().reset_index(drop=True)
        else:
            self.weights = np.asarray(self.weights)[inds].copy()

    # -- whitening (host, sklearn PCA: the reference's own numbers) ---------------------------------
    def fit_pca(self):
        """Standardize -> PCA -> Standardize (emulator.py:286-320)."""
        import sklearn.decomposition as decomp
        self.mean1 = weighted_mean(self.data, self.weights)
        _data = self.data - self.mean1
        subset = self.select_subset(_data, self.weights, nsample=100000)
        self.pca = decomp.PCA()
        self.pca.fit(subset)
        _data = self.pca.transform(_data)
        self.std2 = weighted_std(_data, self.weights)
        rotation_matrix = self.pca.components_
        scale_matrix = np.diag(1. / self.std2)
        self._jacobian_matrix = np.dot(scale_matrix, rotation_matrix)
        self._jacobian_matrix_inv = np.linalg.inv(self._jacobian_matrix)
        self._jacobian_det = np.linalg.det(self._jacobian_matrix_inv)
        self.pca_params = {
            "mean1": self.mean1.copy(),
            "std2": self.std2.copy(),
            "pca": copy.deepcopy(self.pca),
        }

    def set_whitening(self, mean1, pca_mean, components, std2):
        """Install whitening parameters fitted elsewhere (e.g. snapshotted from a reference container:
        ``fit_pca`` is stochastic) instead of re-fitting."""
        import sklearn.decomposition as decomp
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
        self.pca_params = {"mean1": self.mean1.copy(), "std2": self.std2.copy(), "pca": copy.deepcopy(self.pca)}
        self._data = self.pca_transform(self.data)

    def pca_transform(self, data):
        """emulator.py:322-327."""
        _data = data - self.mean1
        _data = self.pca.transform(_data)
        _data /= self.std2
        return _data

    def pca_inverse_transform(self, data):
        """emulator.py:329-335."""
        _data = data * self.std2
        _data = self.pca.inverse_transform(_data)
        _data = _data + self.mean1
        res = pd.DataFrame(_data, columns=self.columns)
        return res

    def standardize_data(self):
        """emulator.py:337-340."""
        self.fit_pca()
        self._data = self.pca_transform(self.data)

    def select_subset(self, data, weights, nsample=10000):
        """emulator.py:342-349."""
        indexes = np.arange(len(data))
        ww = weights / weights.sum()
        inds = self.rng.choice(indexes, size=int(nsample), p=ww, replace=True)
        subset = data.iloc[inds]
        return subset

    def whitening_maps(self):
        """(mu, W, Winv) with z = (x - mu) . W and x = z . Winv + mu — the two affine maps of
        pca_transform / pca_inverse_transform folded for the kernels."""
        comp = np.asarray(self.pca.components_, dtype=np.float64)
        std2 = np.asarray(self.std2, dtype=np.float64)
        mu = np.asarray(self.mean1, dtype=np.float64) + np.asarray(self.pca.mean_, dtype=np.float64)
        W = comp.T / std2[None, :]
        Winv = std2[:, None] * comp
        return mu, W, Winv

    # -- the hot path -------------------------------------------------------------------------------
    def construct_kde(self, bandwidth):
        """emulator.py:351-356: fit = upload (whitened rows, weights, whitening maps)."""
        self.bandwidth = bandwidth
        if getattr(self, "kde", None) is not None:          # re-fit (calc_scores does, chunk after chunk): free the old device copy now
            self.kde.close()
        self.kde = GaussianKDE(bandwidth=self.bandwidth, kernel=self._kernel, atol=self._atol, rtol=self._rtol,
                               breadth_first=self._breadth_first, device=self._device)
        self.kde.fit(np.asarray(self._data), sample_weight=np.asarray(self.weights, dtype=np.float64),
                     whitening=self.whitening_maps(), is_whitened=True)

    def random_draw(self, num, rmin=None, rmax=None, rcol="LOGR"):
        """draws random samples from KDE maximum radius (emulator.py:358-378)."""
        num = int(num)
        windowed = (rmin is not None) or (rmax is not None)
        icol = -1
        if windowed:
            if rmin is None:
                rmin = self.data[rcol].min()
            if rmax is None:
                rmax = self.data[rcol].max()
            icol = list(self.columns).index(rcol)
        if self._rng_mode == "philox":
            seed = int(self.rng.randint(0, 2 ** 31 - 1)) | (int(self.rng.randint(0, 2 ** 31 - 1)) << 31)
            x, fl, nfl = self.kde.sample_philox(num, seed, rcol=icol, rmin=rmin if windowed else 0.0,
                                                rmax=rmax if windowed else 0.0)
            if nfl:
                raise RuntimeError("random_draw: %d draws never entered the window [%g, %g]" % (nfl, rmin, rmax))
            self.res = pd.DataFrame(x, columns=self.columns)
            return self.res
        if self._rng_mode != "host":
            raise ValueError("unknown _rng_mode %r" % (self._rng_mode,))
        # stream-for-stream replay of the reference: n uniforms then n*d normals per sample() call,
        # redraw calls consume n_bad then n_bad*d (emulator.py:371-375).
        d = self.ndim
        _res = np.empty((num, d), dtype=np.float64)
        xraw = np.empty((num, d), dtype=np.float64)
        u = self.rng.uniform(0, 1, size=num)
        z = self.rng.normal(0.0, 1.0, size=(num, d))
        self.kde.draw_streams(u, z, xraw, _res)
        if windowed:
            inds = (xraw[:, icol] > rmax) | (xraw[:, icol] < rmin)
            while inds.sum():
                rows = np.ascontiguousarray(np.nonzero(inds)[0].astype(np.int64))
                nb = len(rows)
                u = self.rng.uniform(0, 1, size=nb)
                z = self.rng.normal(0.0, 1.0, size=(nb, d))
                self.kde.draw_streams(u, z, xraw, _res, rows=rows)
                inds = (xraw[:, icol] > rmax) | (xraw[:, icol] < rmin)
        self._res = _res
        self.res = pd.DataFrame(xraw, columns=self.columns)
        return self.res

    def score_samples(self, arr):
        """Assuming that arr is in the data format (emulator.py:380-385): columns are positional."""
        # (the finiteness check runs once, inside the KDE's own score_samples)
        if isinstance(arr, (pd.DataFrame, pd.Series)) or isinstance(arr, np.ndarray) or _cabi._is_torch(arr):
            X = _as_matrix(arr, check_finite=False)
        else:
            X = _as_matrix(np.asarray(arr), check_finite=False)
        if X.shape[1] != self.ndim:
            raise ValueError("X has %d features, but KDEContainer is expecting %d features as input."
                             % (X.shape[1], self.ndim))
        res = self.kde.score_samples(X, is_whitened=False)
        return res, self._jacobian_det

    def drop_kde(self):
        """emulator.py:387-389 (also frees the device copy)."""
        if getattr(self, "kde", None) is not None:
            self.kde.close()
        self.pca = None
        self.kde = None

    def drop_col(self, colname):
        """emulator.py:391-394."""
        self.data = self.data.drop(columns=colname)
        self.columns = self.data.columns
        self.ndim = len(self.columns)

    def __getstate__(self):
        state = self.__dict__.copy()
        if state.get("kde", None) is not None:
            raise TypeError("KDEContainer holds a device KDE; call drop_kde() before pickling "
                            "(the reference has the same rule for its KD-tree, emulator.py:387-389)")
        return state

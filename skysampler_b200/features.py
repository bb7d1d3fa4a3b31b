"""Feature tables in front of the KDE path — host-side mirrors of ``skysampler/emulator.py:67-231, 495-521``.

``BaseContainer.construct_features`` (column algebra, limits, log10), ``DeepFeatureContainer``,
``FeatureSpaceContainer`` (radially balanced, weight-preserving downsample) and the two
``construct_*_container`` helpers the drivers call before ``make_*_infodicts``.  Same names, arguments and
results as the reference; differences are noted inline.  ``construct_features(..., device=k)`` runs the column
algebra, the limit cuts and the log10 on GPU k (``skb_features`` + ``skb_compact``: one elementwise kernel over the
catalogue rows and an order-preserving compaction) and brings the same DataFrame back; ``features_on_device`` keeps the
result there (torch CUDA tensors ready for ``GaussianKDE.fit``).  The radially balanced downsample stays on the host:
its draws are defined by numpy's ``RandomState.choice(replace=False, p=...)`` stream.
"""
import copy

import numpy as np
import pandas as pd

from .kde import KDEContainer

# binary column operators of a feature spec ("A", "B", op), emulator.py:98-111
_OPS = {
    "-": lambda a, b: a - b,
    "+": lambda a, b: a + b,
    "*": lambda a, b: a * b,
    "/": lambda a, b: a / b,
    "SQSUM": lambda a, b: np.sqrt(a ** 2. + b ** 2.),
}


class BaseContainer(object):
    def __init__(self):
        self.alldata = None
        self.features = None
        self.weights = None

    def _operand(self, spec):
        """A column name, a (name, index) pair into a vector column, or a literal (emulator.py:84-96)."""
        if isinstance(spec, str):
            return self.alldata[spec]
        if isinstance(spec, (list, tuple)):
            return self.alldata[spec[0]][:, spec[1]]
        return spec

    _OPCODES = {"-": 1, "+": 2, "*": 3, "/": 4, "SQSUM": 5}

    def features_on_device(self, columns, limits=None, logs=None, device=0):
        """The feature table on GPU ``device``: (features [n_keep, d] float64 CUDA tensor, kept row ids int64 CUDA tensor).
        Same column algebra / cuts / log10 as ``construct_features`` (emulator.py:74-132), evaluated by ``skb_features``."""
        import ctypes as C
        import torch
        from . import _cabi
        dev = torch.device("cuda", device)
        n = len(self.alldata)
        cache, keepalive = {}, []

        def operand(spec):
            """-> (device pointer or None, dtype code, stride, constant)"""
            if isinstance(spec, str) or isinstance(spec, (list, tuple)):
                name, comp = (spec, None) if isinstance(spec, str) else (spec[0], spec[1])
                if name not in cache:
                    arr = np.asarray(self.alldata[name])
                    if arr.dtype not in (np.float32, np.float64):
                        arr = arr.astype(np.float64)
                    cache[name] = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
                t = cache[name]
                stride = 1 if t.dim() == 1 else t.shape[1]
                off = 0 if comp is None else int(comp)
                return t.data_ptr() + off * t.element_size(), _cabi.dtype_code(t), stride, 0.0
            return None, _cabi.SKB_F64, 1, float(spec)

        specs = (_cabi.SkbFeatureSpec * len(columns))()
        for i, (name, src) in enumerate(columns):
            sp = specs[i]
            if isinstance(src, str):
                a, b, op = src, 0.0, 0
            elif len(src) == 3:
                if src[2] not in self._OPCODES:
                    raise KeyError("only + - * / are supported at the moment")
                a, b, op = src[0], src[1], self._OPCODES[src[2]]
            elif len(src) == 2:
                a, b, op = (src[0], src[1]), 0.0, 0
            else:
                raise KeyError
            pa, da, sa, ca = operand(a)
            pb, db, sb, cb = operand(b)
            sp.a, sp.a_dtype, sp.a_stride, sp.a_const = pa, da, sa, ca
            sp.b, sp.b_dtype, sp.b_stride, sp.b_const = pb, db, sb, cb
            sp.op = op
            sp.lo = float(limits[i][0]) if limits is not None else -np.inf
            sp.hi = float(limits[i][1]) if limits is not None else np.inf
            sp.log10 = 1 if (logs is not None and logs[i]) else 0
        d = len(columns)
        out = torch.empty((n, d), dtype=torch.float64, device=dev)
        keep = torch.empty(n, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev)
        lib = _cabi.lib()
        _cabi.check(lib.skb_features(specs, d, n, _cabi.ptr(out), _cabi.ptr(keep), _cabi.stream_ptr(st)))
        packed = torch.empty_like(out)
        index = torch.empty(n, dtype=torch.int64, device=dev)
        cnt = C.c_int64(0)
        _cabi.check(lib.skb_compact(_cabi.ptr(keep), n, _cabi.ptr(out), _cabi.SKB_F64, d, _cabi.ptr(packed), _cabi.ptr(index),
                                    C.byref(cnt), _cabi.stream_ptr(st)))
        k = int(cnt.value)
        return packed[:k], index[:k]

    def construct_features(self, columns, limits=None, logs=None, device=None, **kwargs):
        """emulator.py:74-132: build the feature columns, cut on ``limits`` (open intervals), then log10.
        ``device=k``: the same table computed on GPU k (see ``features_on_device``) and copied back."""
        self.columns = columns
        self.limits = limits
        self.logs = logs
        if device is not None:
            feats, index = self.features_on_device(columns, limits=limits, logs=logs, device=device)
            rows = index.cpu().numpy()
            self.inds = np.zeros(len(self.alldata), dtype=bool)
            self.inds[rows] = True
            self.features = pd.DataFrame(feats.cpu().numpy(), columns=[c[0] for c in columns], index=rows)
            try:
                self.weights = self.alldata["WEIGHT"][self.inds]
            except Exception:
                self.weights = pd.Series(data=np.ones(len(self.features)), name="WEIGHT")
            return
        self.features = pd.DataFrame()
        self.inds = np.ones(len(self.alldata), dtype=bool)
        for i, (name, src) in enumerate(columns):
            if isinstance(src, str):
                res = self.alldata[src]
            elif len(src) == 3:
                if src[2] not in _OPS:
                    raise KeyError("only + - * / are supported at the moment")
                res = _OPS[src[2]](self._operand(src[0]), self._operand(src[1]))
            elif len(src) == 2:
                res = self.alldata[src[0]][:, src[1]]
            else:
                raise KeyError
            self.features[name] = np.asarray(res).astype("float64")
            if limits is not None:
                self.inds &= np.asarray((self.features[name] > limits[i][0]) & (self.features[name] < limits[i][1]))
        self.features = self.features[self.inds]
        try:
            self.weights = self.alldata["WEIGHT"][self.inds]
        except Exception:
            self.weights = pd.Series(data=np.ones(len(self.features)), name="WEIGHT")
        for i, (name, src) in enumerate(columns):
            if logs is not None and logs[i]:
                self.features[name] = np.log10(self.features[name])

    def to_kde(self, **kwargs):
        """emulator.py:134-136."""
        return KDEContainer(self.features, weights=self.weights)


class FeatureSpaceContainer(BaseContainer):
    """Wide-field (indexed survey) features, emulator.py:139-209."""

    def __init__(self, info):
        BaseContainer.__init__(self)
        self.rcens = info.rcens
        self.redges = info.redges
        self.rareas = info.rareas
        self.survey = info.survey
        self.target = info.target
        self.numprof = info.numprof
        self.samples = info.samples
        parts = [tmp for tmp in self.samples if len(tmp) > 0]
        self.alldata = pd.concat(parts).reset_index(drop=True)
        self.nobj = self.target.nrow

    def surfdens(self, icol=0, scaler=1):
        """emulator.py:164-170."""
        arr = self.features.values[:, icol]
        if self.logs[icol]:
            arr = 10 ** arr
        return np.histogram(arr, bins=self.redges, weights=self.weights)[0] / self.nobj / self.rareas * scaler

    def downsample(self, nmax=10000, r_key="LOGR", nbins=40, rng=None, **kwargs):
        """Radially balanced downsampling (emulator.py:172-209): at most ``nmax`` rows per radial bin, drawn
        without replacement with probability proportional to weight, kept weights rescaled by len/nmax."""
        if rng is None:
            rng = np.random.RandomState()
        rarr = self.features[r_key]
        rbins = np.linspace(rarr.min(), rarr.max(), nbins + 1)
        keep_f, keep_w = [], []
        for lo, hi in zip(rbins[:-1], rbins[1:]):
            sel = (self.features[r_key] > lo) & (self.features[r_key] < hi)
            vals = self.features.loc[sel]
            ww = self.weights.loc[sel]
            if len(vals) >= nmax:
                pick = rng.choice(np.arange(len(vals)), size=nmax, replace=False, p=ww / ww.sum())
                vals, ww = vals.iloc[pick], ww.iloc[pick] * len(ww) / nmax
            keep_f.append(vals)
            keep_w.append(ww)
        return KDEContainer(pd.concat(keep_f), weights=pd.concat(keep_w))


class DeepFeatureContainer(BaseContainer):
    """Deep-field table (a record array), emulator.py:212-231."""

    def __init__(self, data):
        BaseContainer.__init__(self)
        self.alldata = data
        self.weights = pd.Series(data=np.ones(len(self.alldata)), name="WEIGHT")

    @classmethod
    def from_file(cls, fname, flagsel=True):
        """emulator.py:218-231 (FITS through fitsio when installed, else this package's reader)."""
        if ".fit" in fname:
            from .emulator import _fits
            _deep = _fits().read(fname)
        else:
            _deep = pd.read_hdf(fname, key="data").to_records()
        deep = _deep[_deep["flags"] == 0] if flagsel else _deep
        return cls(deep)


def construct_wide_container(dataloader, settings, nbins=100, nmax=5000, seed=None, drop=None, **kwargs):
    """emulator.py:495-507.  (The reference's ``shuffle()`` raises on Python 3, emulator.py:279; the
    container here implements the intended behaviour.)"""
    fsc = FeatureSpaceContainer(dataloader)
    fsc.construct_features(**settings)
    cont_small = fsc.downsample(nbins=nbins, nmax=nmax, kwargs=kwargs)
    cont_small.set_seed(seed)
    cont_small.shuffle()
    if drop is not None:
        cont_small.drop_col(drop)
    settings = copy.copy(settings)
    settings.update({"container": cont_small})
    return settings


def construct_deep_container(data, settings, seed=None, frac=1., drop=None):
    """emulator.py:510-521."""
    fsc = DeepFeatureContainer(data)
    fsc.construct_features(**settings)
    cont = fsc.to_kde()
    if drop is not None:
        cont.drop_col(drop)
    cont.set_seed(seed)
    cont.sample(frac=frac)
    settings = copy.copy(settings)
    settings.update({"container": cont})
    return settings

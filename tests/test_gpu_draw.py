"""KDE draws through the C ABI: stream-for-stream replay of the reference, Philox mode, compaction."""
import ctypes as C

import numpy as np
import pandas as pd
import pytest

from conftest import load_golden
from oracle import draw

pytestmark = pytest.mark.gpu


def _container(g, data, weights, cols, seed=None):
    from skysampler_b200 import KDEContainer
    c = KDEContainer(pd.DataFrame(data, columns=cols), pd.Series(weights) if bool(g["weighted"]) else None, seed=seed)
    return c


@pytest.mark.parametrize("case", ["draw_w", "draw_u"])
def test_centre_indices_bit_exact_from_host_streams(case, c1_train):
    data, weights, cols = c1_train
    g = load_golden(case)
    c = _container(g, data, weights, cols)
    c.set_whitening(g["mean1"], g["pca_mean"], g["components"], g["std2"])
    c.construct_kde(float(g["h"]))
    from skysampler_b200 import _cabi
    n, d = g["z"].shape
    idx = np.empty(n, dtype=np.int64); xz = np.empty((n, d)); xr = np.empty((n, d))
    u = np.ascontiguousarray(g["u"]); z = np.ascontiguousarray(g["z"])
    _cabi.check(_cabi.lib().skb_draw(c.kde._handle(), n, _cabi.ptr(u), _cabi.ptr(z), 0, 0, None, -1, 0.0, 0.0, 1,
                                     _cabi.ptr(idx), _cabi.ptr(xr), d, _cabi.ptr(xz), None, None, None))
    assert np.array_equal(idx, g["idx"])                                   # bit-exact centre selection
    np.testing.assert_allclose(xz, g["res_whitened"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(xr, g["res_plain"], rtol=0, atol=1e-11)
    c.drop_kde()


@pytest.mark.parametrize("case", ["draw_w", "draw_u"])
def test_random_draw_replays_the_reference_including_the_window_loop(case, c1_train):
    """Same seed -> same whitening resample, same stream consumption, same redraw loop
    (emulator.py:358-378): the drop-in returns the reference's DataFrame."""
    data, weights, cols = c1_train
    g = load_golden(case)
    c = _container(g, data, weights, cols, seed=int(g["seed"]))
    c.standardize_data()
    c.construct_kde(float(g["h"]))
    res = c.random_draw(4000)
    assert isinstance(res, pd.DataFrame) and list(res.columns) == cols
    np.testing.assert_allclose(res.values, g["res_plain"], rtol=0, atol=1e-10)
    lo, hi = g["window"]
    res = c.random_draw(4000, rmin=lo, rmax=hi)
    np.testing.assert_allclose(res.values, g["res_window"], rtol=0, atol=1e-10)
    assert res["LOGR"].min() >= lo and res["LOGR"].max() <= hi
    c.drop_kde()


def test_unweighted_handle_uses_truncation_rule():
    from skysampler_b200 import GaussianKDE
    rs = np.random.RandomState(0)
    z = rs.normal(size=(1000, 2))
    kde = GaussianKDE(0.1).fit(z)                    # no weights at all -> (u*N).astype(int64)
    u = rs.uniform(size=5000); u[:4] = [0.0, 0.25, 0.5, 0.999999999999]
    zz = rs.normal(size=(5000, 2))

    class RS(np.random.RandomState):
        pass
    streams = iter([u, zz])
    fake = RS(0)
    fake.uniform = lambda *a, **k: next(streams)
    fake.normal = lambda *a, **k: next(streams)
    res, idx = kde.sample(5000, random_state=fake, return_index=True)
    assert np.array_equal(idx, (u * 1000).astype(np.int64))
    np.testing.assert_allclose(res, z[idx] + 0.1 * zz, rtol=0, atol=1e-14)
    kde.close()


def test_philox_draws_statistics_window_and_determinism(c1_train):
    data, weights, cols = c1_train
    g = load_golden("draw_w")
    c = _container(g, data, weights, cols, seed=3)
    c.set_whitening(g["mean1"], g["pca_mean"], g["components"], g["std2"])
    c.construct_kde(0.1)
    n = 400000
    x1, fl, nfl = c.kde.sample_philox(n, seed=1234)
    x2, _, _ = c.kde.sample_philox(n, seed=1234)
    x3, _, _ = c.kde.sample_philox(n, seed=1235)
    assert nfl == 0 and np.array_equal(x1, x2) and not np.array_equal(x1, x3)
    # chunked generation with an offset is the same stream
    xa, _, _ = c.kde.sample_philox(1000, seed=1234, offset=5000)
    assert np.array_equal(xa, x1[5000:6000])
    # moments: draws are training rows (weighted) + h * N(0, I) in whitened space
    z = draw.whiten(x1, g["mean1"], g["pca_mean"], g["components"], g["std2"])
    zt = np.asarray(c._data); w = weights / weights.sum()
    mean_t = (zt * w[:, None]).sum(0)
    cov_t = (zt * w[:, None]).T @ zt - np.outer(mean_t, mean_t) + 0.01 * np.eye(4)
    assert np.abs(z.mean(0) - mean_t).max() < 6 * np.sqrt(np.diag(cov_t).max() / n)
    assert np.abs(np.cov(z.T) - cov_t).max() < 0.02
    # window: in-kernel rejection, every row inside, marginal matches the truncated reference draws
    lo, hi = g["window"]
    icol = cols.index("LOGR")
    xw, fl, nfl = c.kde.sample_philox(100000, seed=7, rcol=icol, rmin=lo, rmax=hi)
    assert nfl == 0 and fl.sum() == 0
    assert xw[:, icol].min() >= lo and xw[:, icol].max() <= hi
    ref = g["res_window"][:, icol]
    assert abs(xw[:, icol].mean() - ref.mean()) < 5 * ref.std() / np.sqrt(len(ref))
    # an impossible window is reported, not looped forever
    _, fl, nfl = c.kde.sample_philox(1000, seed=7, rcol=icol, rmin=1e6, rmax=2e6, max_attempts=5)
    assert nfl == 1000 and fl.all()
    # container-level Philox mode
    c._rng_mode = "philox"
    res = c.random_draw(5000, rmin=lo, rmax=hi)
    assert res.shape == (5000, 4) and res["LOGR"].between(lo, hi).all()
    c.drop_kde()


def test_philox_index_distribution_follows_the_weights():
    from skysampler_b200 import GaussianKDE, _cabi
    rs = np.random.RandomState(5)
    n = 64
    z = rs.normal(size=(n, 2)); w = rs.uniform(0.0, 3.0, n); w[::7] = 0.0
    kde = GaussianKDE(0.01).fit(z, sample_weight=w)
    m = 2000000
    import torch
    idx = torch.empty(m, dtype=torch.int64, device="cuda")
    x = torch.empty((m, 2), dtype=torch.float64, device="cuda")
    _cabi.check(_cabi.lib().skb_draw(kde._handle(), m, None, None, 99, 0, None, -1, 0.0, 0.0, 1, _cabi.ptr(idx),
                                     _cabi.ptr(x), 2, None, None, None, None))
    counts = np.bincount(idx.cpu().numpy(), minlength=n).astype(np.float64)
    p = w / w.sum()
    assert np.all(counts[w == 0] == 0)
    ok = p > 0
    chi2 = ((counts[ok] - m * p[ok]) ** 2 / (m * p[ok])).sum()
    assert chi2 < 2.0 * (p > 0).sum()                    # ~dof, generous
    kde.close()


@pytest.mark.parametrize("m,frac", [(0, 0.5), (1, 1.0), (7, 0.0), (2047, 0.3), (2048, 1.0), (100003, 0.057)])
def test_compaction_preserves_order(m, frac):
    import torch
    from skysampler_b200 import _cabi
    rs = np.random.RandomState(m + 1)
    flags = (rs.rand(m) < frac).astype(np.uint8)
    rows = rs.normal(size=(m, 5))
    ft = torch.from_numpy(flags).cuda(); rt = torch.from_numpy(rows).cuda()
    out = torch.full((max(m, 1), 5), -1.0, dtype=torch.float64, device="cuda")
    oi = torch.full((max(m, 1),), -1, dtype=torch.int64, device="cuda")
    cnt = C.c_int64(-1)
    _cabi.check(_cabi.lib().skb_compact(_cabi.ptr(ft), m, _cabi.ptr(rt), 0, 5, _cabi.ptr(out), _cabi.ptr(oi),
                                        C.byref(cnt), None))
    k = int(flags.sum())
    assert cnt.value == k
    assert np.array_equal(out[:k].cpu().numpy(), rows[flags.astype(bool)])
    assert np.array_equal(oi[:k].cpu().numpy(), np.nonzero(flags)[0])


@pytest.mark.parametrize("m,frac,width,dtype,shift", [
    (4096, 0.5, 8, np.float64, 0),            # exactly one chunk
    (4097, 1.0, 8, np.float64, 0),            # one row into the second chunk
    (40 * 4096 + 13, 0.02, 8, np.float64, 0),  # look-back past one 32-chunk window
    (300 * 4096 + 5, 0.0, 4, np.float64, 0),   # nothing accepted over many chunks
    (70 * 4096, 0.93, 3, np.float32, 0),       # fp32 rows of 12 bytes (4-byte copy path)
    (50001, 0.4, 2, np.float64, 3),            # flags not 16-byte aligned (byte-wise flag loads)
])
def test_single_pass_compaction_chunks_lookback_and_alignment(m, frac, width, dtype, shift):
    """The chained-scan compaction (one pass over the flags, decoupled look-back) against numpy's samples[inds]
    (emulator.py:746-747): chunk boundaries, long look-back chains, odd row widths, unaligned flags, device count."""
    import torch
    from skysampler_b200 import _cabi
    rs = np.random.RandomState(m % 1000 + width)
    flags = (rs.rand(m) < frac).astype(np.uint8)
    flags[flags > 0] = rs.randint(1, 256, size=int(flags.sum())).astype(np.uint8)      # any non-zero byte is "accepted"
    rows = rs.normal(size=(m, width)).astype(dtype)
    fbuf = torch.zeros(m + 16, dtype=torch.uint8, device="cuda")
    ft = fbuf[shift:shift + m]
    ft.copy_(torch.from_numpy(flags))
    rt = torch.from_numpy(rows).cuda()
    out = torch.full((m, width), -1.0, dtype=rt.dtype, device="cuda")
    oi = torch.full((m,), -1, dtype=torch.int64, device="cuda")
    cnt_dev = torch.full((1,), -5, dtype=torch.int64, device="cuda")
    for _ in range(2):                                            # twice: the workspace is recycled by the stream-ordered pool
        _cabi.check(_cabi.lib().skb_compact(_cabi.ptr(ft), m, _cabi.ptr(rt), 1 if dtype == np.float32 else 0, width,
                                            _cabi.ptr(out), _cabi.ptr(oi),
                                            C.cast(C.c_void_p(cnt_dev.data_ptr()), C.POINTER(C.c_int64)), None))
    k = int((flags != 0).sum())
    assert int(cnt_dev.item()) == k
    assert np.array_equal(out[:k].cpu().numpy(), rows[flags != 0])
    assert np.array_equal(oi[:k].cpu().numpy(), np.nonzero(flags)[0])
    assert bool((out[k:] == -1.0).all())                          # nothing written past the accepted rows


def test_bracketed_centre_search_equals_searchsorted_on_hard_weights():
    """The draw brackets its search of the cumulative weights with a bucket index (skb_sample.cu: cs_index_kernel);
    the selected centre must still be np.searchsorted(cumsum, u * cumsum[-1]) (sklearn _kde.py:345-347) bit for bit:
    zero weights, 12 decades of dynamic range, uniforms on and next to every bucket edge, u = 0 and u -> 1."""
    from skysampler_b200 import GaussianKDE, _cabi
    rs = np.random.RandomState(5)
    n, d = 20011, 3
    x = rs.normal(size=(n, d))
    w = 10.0 ** rs.uniform(-9, 3, size=n)
    w[rs.rand(n) < 0.3] = 0.0
    w[:50] = 0.0; w[-70:] = 0.0; w[5000:5600] = 0.0
    kde = GaussianKDE(0.2).fit(x, sample_weight=w)
    cs = np.cumsum(w)
    edges = np.arange(n + 1) / float(n)
    u = np.concatenate([edges[:-1], np.nextafter(edges[1:], 0), np.nextafter(edges[:-1], 1), rs.uniform(size=30000),
                        [0.0, np.nextafter(1.0, 0), 0.5]])
    m = len(u)
    z = np.zeros((m, d))
    idx = np.empty(m, dtype=np.int64); xz = np.empty((m, d)); xr = np.empty((m, d))
    _cabi.check(_cabi.lib().skb_draw(kde._handle(), m, _cabi.ptr(u), _cabi.ptr(z), 0, 0, None, -1, 0.0, 0.0, 1,
                                     _cabi.ptr(idx), _cabi.ptr(xr), d, _cabi.ptr(xz), None, None, None))
    want = np.minimum(np.searchsorted(cs, u * cs[-1]), n - 1)
    assert np.array_equal(idx, want)
    assert np.all(w[idx[u > 0]] > 0)
    kde.close()

"""Host-side mirror of the reference surface (no GPU needed): whitening, partition, pickling, columns."""
import pickle

import numpy as np
import pandas as pd
import pytest

from conftest import load_golden
from skysampler_b200 import KDEContainer, emulator, synth
from oracle import draw


def test_fit_pca_reproduces_reference_whitening(c1_train):
    """Same seed, same rng consumption (emulator.py:286-320) -> same mean1 / components / std2 / jac."""
    data, weights, cols = c1_train
    for case, weighted in (("c1_score_w_h0p1", True), ("c1_score_u_h0p1", False)):
        g = load_golden(case)
        c = KDEContainer(pd.DataFrame(data, columns=cols), pd.Series(weights) if weighted else None, seed=11)
        c.standardize_data()
        np.testing.assert_allclose(np.asarray(c.mean1), g["mean1"], rtol=1e-13)
        np.testing.assert_allclose(c.pca.components_, g["components"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(c.std2, g["std2"], rtol=1e-12)
        assert abs(c._jacobian_det - float(g["jac"])) <= 1e-11 * abs(float(g["jac"]))
        np.testing.assert_allclose(np.asarray(c._data)[:64], g["whitened_train_head"], rtol=0, atol=1e-11)
        # the folded affine maps handed to the kernels are the same transform
        mu, W, Winv = c.whitening_maps()
        z = (data - mu) @ W
        np.testing.assert_allclose(z, np.asarray(c._data), rtol=0, atol=1e-11)
        np.testing.assert_allclose(z @ Winv + mu, data, rtol=0, atol=1e-11)
        np.testing.assert_allclose(c.pca_inverse_transform(np.asarray(c._data)).values, data, rtol=0, atol=1e-11)
        assert abs(abs(np.linalg.det(Winv)) - abs(c._jacobian_det)) < 1e-9 * abs(c._jacobian_det)


def test_set_whitening_equals_oracle_transform(c1_train):
    data, weights, cols = c1_train
    g = load_golden("c1_score_w_h0p05")
    c = KDEContainer(pd.DataFrame(data, columns=cols), pd.Series(weights))
    c.set_whitening(g["mean1"], g["pca_mean"], g["components"], g["std2"])
    np.testing.assert_allclose(np.asarray(c._data),
                               draw.whiten(data, g["mean1"], g["pca_mean"], g["components"], g["std2"]),
                               rtol=0, atol=1e-12)


def test_constructor_defaults_and_surface():
    df, w = synth.frame(3, 50, 1)
    c = KDEContainer(df)                              # weights=None -> ones (emulator.py:252-253)
    assert c.ndim == 3 and list(c.columns) == list(df.columns)
    assert np.array_equal(c.weights, np.ones(50))
    for name in ("set_seed", "shuffle", "sample", "fit_pca", "pca_transform", "pca_inverse_transform",
                 "standardize_data", "select_subset", "construct_kde", "random_draw", "score_samples",
                 "drop_kde", "drop_col"):
        assert callable(getattr(c, name))
    assert c._kernel == "gaussian" and c._breadth_first is False
    c.drop_col("F0")
    assert c.ndim == 2
    c2 = KDEContainer(df, w, seed=3)
    a = c2.rng.uniform()
    c2.set_seed(3)
    assert c2.rng.uniform() == a


def test_row_sample_and_shuffle_keep_rows_aligned():
    df, w = synth.frame(2, 200, 5)
    c = KDEContainer(df.copy(), w.copy(), seed=1)
    key = dict(zip(map(tuple, df.values), w.values))
    c.shuffle()
    assert len(c.data) == 200
    assert all(key[tuple(r)] == ww for r, ww in zip(c.data.values, np.asarray(c.weights)))
    c.sample(n=50)
    assert len(c.data) == 50 and len(c.weights) == 50
    c.sample(n=10 ** 6)                                # n > len -> falls back to frac (emulator.py:279-280)
    assert len(c.data) == 50


def test_container_pickles_after_drop_kde():
    df, w = synth.frame(3, 300, 9)
    c = KDEContainer(df, w, seed=2)
    c.standardize_data()
    c.drop_kde()
    c2 = pickle.loads(pickle.dumps(c))
    assert c2.pca is None and c2.kde is None
    np.testing.assert_array_equal(c2._data, c._data)
    assert c2.rng.uniform() == c.rng.uniform()


def test_partition_follows_reference_rounding():
    lst = np.arange(20)                                # utils.py:148-166 examples
    assert [list(p) for p in emulator.partition(lst, 5)] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11],
                                                             [12, 13, 14, 15], [16, 17, 18, 19]]
    assert [len(p) for p in emulator.partition(lst, 6)] == [3, 4, 3, 3, 4, 3]
    parts = emulator.partition(list(range(1501)), 150)
    assert sum(len(p) for p in parts) == 1501 and [x for p in parts for x in p] == list(range(1501))


def test_accept_concentric_matches_golden_masks():
    g = load_golden("driver_naive")
    sc = pd.DataFrame(g["scores_exact"], columns=[str(c) for c in g["score_columns"]])
    for m in (20, 2):
        assert np.array_equal(emulator.accept_concentric(sc, m_factor=m, seed=6), g["accept_m%d" % m])


def test_gaussian_kernel_only():
    from skysampler_b200 import GaussianKDE
    with pytest.raises(ValueError):
        GaussianKDE(0.1, kernel="tophat")


def test_synth_generators_are_seeded():
    a = synth.mixture(4, 100, 7); b = synth.mixture(4, 100, 7)
    assert np.array_equal(a, b) and a.shape == (100, 4)
    w = synth.wide_weights(1000, 1)
    assert w.min() >= 0.5 and w.max() <= 2.0 and len(np.unique(w)) <= 100
    m1, pm, comps, std2 = synth.whiten_params(a)
    z = (a - m1) @ comps.T / std2
    np.testing.assert_allclose(z.std(axis=0), 1.0, rtol=1e-10)


def test_fits_tables_round_trip_and_read_concentric(tmp_path):
    """_samples.fits / _scores.fits layout (one BINTABLE, big-endian scalar columns, 2880-byte blocks) and
    the file-based accept step of read_concentric (emulator.py:717-748) against the golden masks."""
    from skysampler_b200 import fitsio_lite
    g = load_golden("driver_naive")
    samples = pd.DataFrame(g["samples_exact"], columns=[str(c) for c in g["sample_columns"]])
    scores = pd.DataFrame(g["scores_exact"], columns=[str(c) for c in g["score_columns"]])
    out = str(tmp_path / "run_7_rbin0")
    emulator.write_tables(out, samples, scores)
    for suffix, df in (("_samples.fits", samples), ("_scores.fits", scores)):
        raw = open(out + suffix, "rb").read()
        assert len(raw) % 2880 == 0 and raw[:30].decode() == "SIMPLE  =                    T"
        assert b"XTENSION= 'BINTABLE'" in raw[2880:5760]
        rec = fitsio_lite.read(out + suffix)
        assert rec.dtype.names[0] == "index" and list(rec.dtype.names[1:]) == list(df.columns)   # to_records() keeps the index
        assert np.array_equal(rec["index"], np.arange(len(df)))
        for c in df.columns:
            assert np.array_equal(rec[c], df[c].values)                                            # bit exact
    for m in (20, 2):
        res = emulator.read_concentric(str(tmp_path / "run_*_scores.fits"), m_factor=m, seed=6)
        assert np.array_equal(res[samples.columns].values, g["accepted_rows_m%d" % m])
    with pytest.raises(IOError):
        fitsio_lite.write(out + "_samples.fits", samples.to_records())                            # clobber=False
    # mixed dtypes
    rec = np.rec.fromarrays([np.arange(5, dtype=np.int64), np.arange(5, dtype=np.float32), np.arange(5, dtype=np.int32)],
                            names="a,b,c")
    fitsio_lite.write(str(tmp_path / "t.fits"), rec, clobber=True)
    back = fitsio_lite.read(str(tmp_path / "t.fits"))
    assert back.dtype["a"] == np.int64 and back.dtype["b"] == np.float32 and back.dtype["c"] == np.int32
    assert np.array_equal(back["b"], rec["b"])


def test_feature_tables_match_reference():
    """Feature construction, limits, log10, surface density and the radially balanced downsample
    (emulator.py:74-209) against the reference's own outputs."""
    from skysampler_b200 import features
    g = load_golden("features")
    nd = len(g["deep_MAG_G"])
    deep = np.zeros(nd, dtype=[("MAG_G", "f8"), ("MAG_R", "f8"), ("MAG_I", "f8"), ("MAG_Z", "f8"), ("SIZE", "f8"), ("E", "f8", (2,))])
    for n in ("MAG_G", "MAG_R", "MAG_I", "MAG_Z", "SIZE", "E"):
        deep[n] = g["deep_" + n]
    deep_settings = {
        "columns": [("MAG_I", "MAG_I"), ("COLOR_G_R", ("MAG_G", "MAG_R", "-")), ("COLOR_R_I", ("MAG_R", "MAG_I", "-")),
                    ("LOGSIZE", "SIZE"), ("ELL", (("E", 0), ("E", 1), "SQSUM")), ("E1", ("E", 1))],
        "logs": [False, False, False, True, False, False],
        "limits": [(19, 24), (-1, 3), (-1, 3), (1e-3, 2.0), (0, 10), (-5, 5)],
    }
    dfc = features.DeepFeatureContainer(deep)
    dfc.construct_features(**deep_settings)
    assert [c for c in dfc.features.columns] == [c[0] for c in deep_settings["columns"]]
    np.testing.assert_array_equal(np.asarray(dfc.features.index), g["deep_feature_index"])
    np.testing.assert_allclose(dfc.features.values, g["deep_features"], rtol=0, atol=0)
    np.testing.assert_array_equal(np.asarray(dfc.weights), g["deep_weights"])
    kc = dfc.to_kde()
    assert kc.ndim == 6 and len(kc.data) == len(g["deep_features"])
    out = features.construct_deep_container(deep, deep_settings, seed=3, frac=0.5, drop="E1")
    assert out["container"].ndim == 5 and abs(len(out["container"].data) - len(g["deep_features"]) / 2) <= 1

    class Info(object):
        pass
    info = Info()
    sizes = [int(s) for s in g["wide_part_sizes"]]
    parts, a = [], 0
    for sz in sizes:
        parts.append(pd.DataFrame(g["wide_parts"][a:a + sz], columns=["MAG_R", "MAG_I", "DIST", "WEIGHT"])); a += sz
    info.samples = parts
    info.redges = g["wide_redges"]; info.rcens = np.sqrt(info.redges[1:] * info.redges[:-1]); info.rareas = g["wide_rareas"]
    info.survey = None; info.numprof = None
    info.target = Info(); info.target.nrow = 50
    wide_settings = {"columns": [("COLOR_R_I", ("MAG_R", "MAG_I", "-")), ("LOGR", "DIST")], "logs": [False, True],
                     "limits": [(-1, 3), (1e-3, 16.0)]}
    fsc = features.FeatureSpaceContainer(info)
    fsc.construct_features(**wide_settings)
    np.testing.assert_allclose(fsc.features.values, g["wide_features"], rtol=0, atol=0)
    np.testing.assert_array_equal(np.asarray(fsc.weights), g["wide_weights"])
    np.testing.assert_allclose(fsc.surfdens(icol=1), g["wide_surfdens"], rtol=1e-14)
    small = fsc.downsample(nmax=150, nbins=10, rng=np.random.RandomState(5))
    np.testing.assert_array_equal(small.data.values, g["wide_down_data"])
    np.testing.assert_array_equal(np.asarray(small.weights), g["wide_down_weights"])
    # an empty radial bin (the reference trips over it under numpy >= 1.24, emulator.py:156-158) is simply skipped
    info.samples = [parts[0], pd.DataFrame(columns=parts[0].columns), parts[1], parts[2]]
    fsc2 = features.FeatureSpaceContainer(info)
    fsc2.construct_features(**wide_settings)
    np.testing.assert_allclose(fsc2.features.values.astype(float), g["wide_features"], rtol=0, atol=0)
    wc = features.construct_wide_container(info, wide_settings, nbins=10, nmax=150, seed=4)
    assert len(wc["container"].data) == 1500 and wc["container"].ndim == 2


def test_query_validation_matches_sklearn_and_takes_the_one_pass_shortcut():
    """`KernelDensity.score_samples` rejects NaN / inf rows with a ValueError (validate_data, sklearn _kde.py:272); the
    host mirror does the same with sklearn's own shortcut (sum first, element-wise only if the sum is not finite)."""
    from skysampler_b200.kde import _as_matrix
    rs = np.random.RandomState(3)
    x = rs.normal(size=(1000, 3))
    assert _as_matrix(x) is x or np.array_equal(_as_matrix(x), x)
    for bad in (np.nan, np.inf, -np.inf):
        y = x.copy(); y[517, 1] = bad
        with pytest.raises(ValueError, match="NaN or infinity"):
            _as_matrix(y)
        assert np.array_equal(_as_matrix(y, check_finite=False), y, equal_nan=True)
    big = np.full((4, 2), 1.0e308)                      # finite values whose sum overflows: slow path, still accepted
    assert np.array_equal(_as_matrix(big), big)
    f = x.astype(np.float32)
    assert _as_matrix(f).dtype == np.float32
    assert _as_matrix(np.asfortranarray(x)).flags["C_CONTIGUOUS"]

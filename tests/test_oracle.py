"""The oracle against the reference's own outputs (tests/golden, made by oracle/make_golden.py)."""
import numpy as np
import pandas as pd
import pytest

from conftest import load_golden
from oracle import brute64, draw, accept
from oracle.container import RefKDEContainer

CASES = ["c1_score_w_h0p05", "c1_score_w_h0p1", "c1_score_u_h0p05", "c1_score_u_h0p1"]


def _container(g, data, weights, cols, atol, rtol):
    df = pd.DataFrame(data, columns=cols)
    w = pd.Series(weights) if bool(g["weighted"]) else None
    c = RefKDEContainer(df, w, seed=0, atol=atol, rtol=rtol)
    c.set_whitening(g["mean1"], g["pca_mean"], g["components"], g["std2"])
    return c


@pytest.mark.parametrize("case", CASES)
def test_container_restatement_reproduces_reference_scores(case, c1_train):
    data, weights, cols = c1_train
    g = load_golden(case)
    for name, tol in (("exact", 0.0), ("ref", 1e-6)):
        c = _container(g, data, weights, cols, tol, tol)
        np.testing.assert_allclose(np.asarray(c._data)[:64], g["whitened_train_head"], rtol=0, atol=1e-12)
        c.construct_kde(float(g["h"]))
        lp, jac = c.score_samples(pd.DataFrame(g["query"], columns=cols))
        np.testing.assert_allclose(lp, g["logp_" + name], rtol=0, atol=1e-9)
        assert abs(jac - float(g["jac"])) <= 1e-12 * abs(float(g["jac"]))


@pytest.mark.parametrize("case", CASES)
def test_brute64_matches_sklearn_exact_on_the_bulk(case, c1_train):
    """sklearn's atol=rtol=0 walk is exact only near the peak (SURVEY finding 5): compare there,
    and document how far it drifts in the tails."""
    data, weights, cols = c1_train
    g = load_golden(case)
    w = weights if bool(g["weighted"]) else None
    z = draw.whiten(data, g["mean1"], g["pca_mean"], g["components"], g["std2"])
    zq = draw.whiten(g["query"], g["mean1"], g["pca_mean"], g["components"], g["std2"])
    b = brute64.score(z, w, float(g["h"]), zq)
    ref = g["logp_exact"]
    bulk = b >= b.max() - 11.0
    assert bulk.sum() > 500
    assert np.abs(b - ref)[bulk].max() <= 1e-6
    mid = (b >= b.max() - 17.0)
    assert np.abs(b - ref)[mid].max() <= 1e-3
    assert np.all(np.isfinite(b))


def test_brute64_known_answer():
    # one training point: log p = log N(z; x, h^2 I)
    z = np.array([[0.3, -0.2, 0.1]])
    q = np.array([[0.0, 0.0, 0.0], [0.3, -0.2, 0.1]])
    h = 0.25
    expect = -0.5 * ((q - z) ** 2).sum(1) / h ** 2 - 1.5 * np.log(2 * np.pi) - 3 * np.log(h)
    np.testing.assert_allclose(brute64.score(z, None, h, q), expect, rtol=0, atol=1e-13)
    # weights only matter through their ratio
    zz = np.vstack([z, z + 1.0])
    a = brute64.score(zz, np.array([1.0, 3.0]), h, q)
    b = brute64.score(zz, np.array([10.0, 30.0]), h, q)
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-13)


def test_partial_lse_is_additive_over_training_shards():
    rs = np.random.RandomState(3)
    z = rs.normal(size=(3000, 5)); w = rs.uniform(0.1, 2, 3000); q = 2 * rs.normal(size=(200, 5))
    full = brute64.score(z, w, 0.2, q)
    parts = [brute64.partial_lse(z[a:b], w[a:b], 0.2, q) for a, b in ((0, 700), (700, 701), (701, 3000))]
    m, s = brute64.combine_partials([p[0] for p in parts], [p[1] for p in parts])
    lp = brute64.log_knorm(5, 0.2) + m + np.log(s) - np.log(w.sum())
    np.testing.assert_allclose(lp, full, rtol=0, atol=1e-11)


@pytest.mark.parametrize("case", ["draw_w", "draw_u"])
def test_draw_restatement_bit_exact(case, c1_train):
    data, weights, cols = c1_train
    g = load_golden(case)
    w = weights if bool(g["weighted"]) else np.ones(len(data))       # the reference always passes weights (:252-255)
    z = draw.whiten(data, g["mean1"], g["pca_mean"], g["components"], g["std2"])
    zs, idx = draw.sample_from_streams(z, float(g["h"]), g["u"], g["z"], weights=w)
    assert np.array_equal(idx, g["idx"])
    np.testing.assert_allclose(zs, g["res_whitened"], rtol=0, atol=1e-12)
    x = draw.inverse_whiten(zs, g["mean1"], g["pca_mean"], g["components"], g["std2"])
    np.testing.assert_allclose(x, g["res_plain"], rtol=0, atol=1e-11)


def test_unweighted_index_rule():
    u = np.array([0.0, 0.2499999, 0.25, 0.999999999])
    assert list(draw.select_index(u, 4)) == [0, 0, 1, 3]
    # weights of ones take the searchsorted branch (side='left'): differs exactly at integer u*N
    assert list(draw.select_index(u, 4, np.ones(4))) == [0, 0, 0, 3]


@pytest.mark.parametrize("case", ["draw_w", "draw_u"])
def test_container_restatement_replays_reference_draws(case, c1_train):
    """Same seed -> same PCA resample, same streams, same redraw loop (emulator.py:358-378)."""
    data, weights, cols = c1_train
    g = load_golden(case)
    df = pd.DataFrame(data, columns=cols)
    w = pd.Series(weights) if bool(g["weighted"]) else None
    c = RefKDEContainer(df, w, seed=int(g["seed"]))
    c.standardize_data()
    np.testing.assert_allclose(c.std2, g["std2"], rtol=1e-12)
    c.construct_kde(float(g["h"]))
    np.testing.assert_allclose(c.random_draw(4000).values, g["res_plain"], rtol=0, atol=1e-11)
    lo, hi = g["window"]
    res = c.random_draw(4000, rmin=lo, rmax=hi)
    np.testing.assert_allclose(res.values, g["res_window"], rtol=0, atol=1e-11)
    assert res["LOGR"].min() >= lo and res["LOGR"].max() <= hi


def test_accept_restatement_matches_read_concentric():
    g = load_golden("driver_naive")
    sc = pd.DataFrame(g["scores_exact"], columns=[str(c) for c in g["score_columns"]])
    for m_factor in (20, 2):
        mask, u = accept.accept(sc["dc"], sc["dc_jac"], sc["wr"], sc["wr_jac"], sc["wcr"], sc["wcr_jac"],
                                m_factor=m_factor, seed=6)
        assert np.array_equal(np.asarray(mask), g["accept_m%d" % m_factor])
        # log-space form used by the fused epilogue agrees away from the knife edge
        lr = accept.log_ratio([sc["wcr"], sc["dc"], sc["wr"]],
                              [np.log(abs(sc["wcr_jac"][0])), np.log(abs(sc["dc_jac"][0])), np.log(abs(sc["wr_jac"][0]))],
                              [1, -1, -1], np.log(m_factor))
        band = np.abs(lr - np.log(u)) <= 1e-9
        assert np.array_equal((np.log(u) < lr)[~band], np.asarray(mask)[~band])
        samples = g["samples_exact"]
        np.testing.assert_array_equal(samples[np.asarray(mask)], g["accepted_rows_m%d" % m_factor])


def test_fp32_window_model_meets_the_parity_bar_and_culling_is_exact():
    """The kernels' fp32 scheme restated in numpy (oracle/fp32_model.py): exponent window 90 binades below the
    nearest point, ftz below 2^-126, fp32 lane partials, fp64 folds.  It must (1) meet the 1e-4 parity bar
    against the fp64 oracle within 100 nats of the peak, for clustered data and for tail queries, and
    (2) give the same BITS with and without tile culling — a culled tile contributes exactly nothing."""
    from oracle import brute64, fp32_model
    rs = np.random.RandomState(12)
    for d, h in ((2, 0.1), (5, 0.1), (8, 0.15)):
        centres = 3.0 * rs.normal(size=(4, d))
        z = np.vstack([c + 0.4 * rs.normal(size=(700, d)) for c in centres])
        order = np.lexsort(z.T[::-1])                       # any spatially coherent order makes boxes that cull
        z = z[order]
        w = rs.uniform(0.2, 3.0, len(z))
        q = np.vstack([z[rs.randint(0, len(z), 120)] + h * rs.normal(size=(120, d)),
                       2.0 * z[rs.randint(0, len(z), 20)],
                       40.0 * rs.normal(size=(10, d))])
        ref = brute64.score(z, w, h, q)
        got, skipped = fp32_model.score(z, w, h, q, cull=True)
        dense, none = fp32_model.score(z, w, h, q, cull=False)
        assert none == 0 and skipped > 0.3 * len(q) * (len(z) // 64), (d, skipped)
        assert np.array_equal(got, dense)                     # culling is exact
        assert np.all(np.isfinite(got))
        near = ref >= ref.max() - 100.0
        assert np.abs(got - ref)[near].max() <= 1e-4, (d, np.abs(got - ref)[near].max())
        far = ~near
        assert (np.abs(got - ref)[far] / np.abs(ref[far])).max() <= 1e-5


def test_c_brute_force_twin_matches_the_numpy_oracle():
    """oracle/brute64.c (used at the multi-million-row configs) is the same function as oracle/brute64.py."""
    from oracle import cbrute
    rs = np.random.RandomState(3)
    for d, n, h, weighted in ((1, 3000, 0.05, True), (4, 5000, 0.1, True), (8, 4000, 0.1, False), (16, 3000, 0.3, True)):
        z = rs.normal(size=(n, d))
        w = rs.uniform(0.5, 2.0, n) if weighted else None
        if weighted:
            w[::11] = 0.0
        q = np.vstack([1.5 * rs.normal(size=(300, d)), 40.0 * rs.normal(size=(20, d))])
        a = brute64.score(z, w, h, q)
        b = cbrute.score(z, w, h, q)
        assert np.all(np.isfinite(b)) and np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(a).max())
        ma, sa = brute64.partial_lse(z, w, h, q)
        mb, sb = cbrute.partial_lse(z, w, h, q)
        assert np.array_equal(ma, mb) and np.allclose(sa, sb, rtol=1e-12, atol=0)


def test_far_field_record_remainder_bound_and_fp32_evaluation():
    """The admission rules of the d <= 2 far-field paths (1-D: P = 5, z <= 1/4; 2-D: P = 12, z <= 1.7) rest on the bound
    z^(P+1)/(P+1)! e^(2z) for the record's relative error.  Checked here in fp64 on random tiles (it must hold for EVERY
    point layout), together with the fp32 Horner evaluation the kernels run (error << the 1e-4 contract)."""
    from oracle import farfield_model as ff
    rs = np.random.RandomState(12)
    assert ff.remainder_bound(0.25, 5) < 6e-7 and ff.remainder_bound(1.7, 12) < 5e-6
    worst = {1: 0.0, 2: 0.0}
    worst32 = {1: 0.0, 2: 0.0}
    for d, P, zmax in ((1, 5, 0.25), (2, 12, 1.7)):
        for trial in range(300):
            n = int(rs.randint(1, 129))
            r = 10.0 ** rs.uniform(-3, 0.3, size=d)                       # half widths of the tile
            layout = trial % 3
            if layout == 0:
                t = rs.uniform(-1, 1, size=(n, d)) * r
            elif layout == 1:                                             # everything in the corners: the bound's worst case
                t = rs.choice([-1.0, 1.0], size=(n, d)) * r
            else:                                                         # one-sided
                t = np.abs(rs.uniform(0.5, 1, size=(n, d))) * r
            w = 10.0 ** rs.uniform(-6, 0, size=n)
            B = ff.coefficients(t, w, P)
            # offsets on the admission boundary z = 2 ln2 sum |s_k| r_k = zmax and inside it, all sign patterns
            for frac in (1.0, 0.6, 0.2):
                share = rs.dirichlet(np.ones(d))
                s = share * (frac * zmax / (2 * ff.LN2)) / r * rs.choice([-1.0, 1.0], size=d)
                z = 2 * ff.LN2 * float((np.abs(s) * r).sum())
                assert z <= zmax * (1 + 1e-12)
                want = ff.exact(t, w, s)
                got = float(ff.evaluate(B, s))
                if want > 1e-290:
                    rel = abs(got - want) / want
                    assert rel <= ff.remainder_bound(z, P) + 1e-11, (d, trial, frac, rel, ff.remainder_bound(z, P))   # + fp64 rounding of the model itself
                    worst[d] = max(worst[d], rel)
                    got32 = float(ff.evaluate({k: np.float32(v) for k, v in B.items()}, s.astype(np.float32), dtype=np.float32))
                    if want > 1e-30:
                        worst32[d] = max(worst32[d], abs(got32 - want) / want)
    assert worst[1] <= 6e-7 and worst[2] <= 5e-6
    # fp32: the exponent -|s|^2 alone carries ~|s|^2 2^-24 (the pair loop has the same term), the Horner chain the rest
    assert worst32[1] <= 2e-5 and worst32[2] <= 3e-5, worst32

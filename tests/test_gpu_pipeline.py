"""The callers of the path (emulator.py:525-748) on the B200 backend vs the reference's golden outputs."""
import numpy as np
import pandas as pd
import pytest

from conftest import load_golden
from oracle import accept as oaccept

pytestmark = pytest.mark.gpu


def _containers(g):
    from skysampler_b200 import KDEContainer
    seeds = [int(s) for s in g["seeds"]]
    deep_c = {"container": KDEContainer(pd.DataFrame(g["deep_c"], columns=["C0", "C1", "C2"]), None, seed=seeds[0])}
    deep_smc = {"container": KDEContainer(pd.DataFrame(g["deep_smc"], columns=["MAG", "C0", "C1", "C2"]), None, seed=seeds[1])}
    wide_cr = {"container": KDEContainer(pd.DataFrame(g["wide_cr"], columns=["C0", "C1", "C2", "LOGR"]), pd.Series(g["w_cr"]), seed=seeds[2])}
    wide_r = {"container": KDEContainer(pd.DataFrame(g["wide_r"], columns=["LOGR"]), pd.Series(g["w_r"]), seed=seeds[3])}
    columns = {"cols_dc": ["C0", "C1", "C2"], "cols_wr": ["LOGR"], "cols_wcr": ["C0", "C1", "C2", "LOGR"]}
    return wide_cr, wide_r, deep_c, deep_smc, columns


def test_rejection_sampler_driver_slice_matches_reference():
    """bin/delta_rejection_sampler_v01.py:95-110: make_naive_infodicts -> run_scores, same seeds."""
    from skysampler_b200 import emulator
    g = load_golden("driver_naive")
    wide_cr, wide_r, deep_c, deep_smc, columns = _containers(g)
    infodicts, samples = emulator.make_naive_infodicts(wide_cr, wide_r, deep_c, deep_smc, columns,
                                                       nsamples=int(g["nsamples"]), nchunks=int(g["nchunks"]),
                                                       bandwidth=float(g["bandwidth"]), rmin=None, rmax=float(g["rmax"]))
    assert list(samples.columns) == [str(c) for c in g["sample_columns"]]
    np.testing.assert_allclose(samples.values, g["samples_exact"], rtol=0, atol=1e-9)       # proposals replayed
    assert len(infodicts) == 4 and sum(len(i["sample"]) for i in infodicts) == len(samples)
    assert wide_r["container"].kde is None                                                  # dropped, picklable
    result = emulator.run_scores(infodicts)
    assert list(result.columns) == [str(c) for c in g["score_columns"]]
    ref = pd.DataFrame(g["scores_exact"], columns=result.columns)
    for col in ("dc_jac", "wr_jac", "wcr_jac"):
        np.testing.assert_allclose(result[col].values, ref[col].values, rtol=1e-10)
    loose = pd.DataFrame(g["scores_ref"], columns=result.columns)
    for col in ("dc", "wr", "wcr"):
        err = np.abs(result[col].values - ref[col].values)
        bulk = ref[col].values >= ref[col].values.max() - 11.0          # where sklearn-exact is exact
        assert err[bulk].max() <= 1e-4, (col, err[bulk].max())
        # (the shipped atol=rtol=1e-6 run is looser than this backend — SURVEY.md finding 6: its own
        #  error reaches ~0.1 nats in the bulk of this data — so only its median is compared)
        assert np.median(np.abs(result[col].values - loose[col].values)[bulk]) <= 1e-3
    # accept decisions: identical to the reference's outside the tolerance band
    for m in (20, 2):
        mine = emulator.accept_concentric(result.reset_index(drop=True), m_factor=m, seed=6)
        gold = g["accept_m%d" % m]
        u = np.random.RandomState(6).uniform(0, 1, len(ref))
        lr = oaccept.log_ratio([ref["wcr"], ref["dc"], ref["wr"]],
                               [np.log(abs(ref["wcr_jac"][0])), np.log(abs(ref["dc_jac"][0])), np.log(abs(ref["wr_jac"][0]))],
                               [1, -1, -1], np.log(m))
        trust = np.ones(len(ref), dtype=bool)
        for col in ("dc", "wr", "wcr"):
            trust &= ref[col].values >= ref[col].values.max() - 11.0
        band = np.abs(lr - np.log(u)) <= 3e-4
        assert np.array_equal(mine[trust & ~band], gold[trust & ~band])
        assert (mine != gold).mean() < 0.02


def test_calc_scores_keeps_per_chunk_semantics():
    from skysampler_b200 import emulator
    g = load_golden("driver_naive")
    wide_cr, wide_r, deep_c, deep_smc, columns = _containers(g)
    infodicts, samples = emulator.make_naive_infodicts(wide_cr, wide_r, deep_c, deep_smc, columns, nsamples=600,
                                                       nchunks=2, bandwidth=0.1, rmax=0.4)
    sc = emulator.calc_scores(infodicts[0])
    assert list(sc.columns) == ["dc", "dc_jac", "wr", "wr_jac", "wcr", "wcr_jac"] and len(sc) == 300
    assert np.all(np.isfinite(sc.values))


def test_classifier_layout_columns():
    from skysampler_b200 import emulator, KDEContainer
    g = load_golden("driver_naive")
    wide_cr, wide_r, deep_c, deep_smc, columns = _containers(g)
    rands = {"container": KDEContainer(pd.DataFrame(g["wide_cr"][::-1].copy(), columns=["C0", "C1", "C2", "LOGR"]),
                                       pd.Series(g["w_cr"][::-1].copy()), seed=31)}
    infodicts, samples = emulator.make_classifier_infodicts(wide_cr, wide_r, rands, deep_c, deep_smc, columns,
                                                            nsamples=800, nchunks=3, bandwidth=0.1, rmax=0.4)
    res = emulator.run_scores2(infodicts)
    assert list(res.columns) == ["dc", "dc_jac", "wr", "wr_jac", "wcr_clust", "wcr_clust_jac", "wcr_rands", "wcr_rands_jac"]
    assert len(res) == 800 and np.all(np.isfinite(res.values))
    one = emulator.calc_scores2(infodicts[1])
    assert list(one.columns) == list(res.columns)


def test_fused_ratio_accept_equals_oracle_decisions():
    """skb_ratio_accept: K KDEs of different dimension, column subsets of one table, accept in the epilogue."""
    from skysampler_b200 import emulator, KDEContainer, synth
    from oracle import brute64, draw
    rs = np.random.RandomState(3)
    n = 4000
    tab = pd.DataFrame(np.column_stack([synth.mixture(3, n, 41), 0.3 * rs.normal(size=n)]), columns=["C0", "C1", "C2", "LOGR"])
    conts, colsets = [], [["C0", "C1", "C2", "LOGR"], ["C0", "C1", "C2"], ["LOGR"]]
    for k, cs in enumerate(colsets):
        x = tab[cs].values + 0.1 * rs.normal(size=(n, len(cs)))
        c = KDEContainer(pd.DataFrame(x, columns=cs), pd.Series(synth.wide_weights(n, 50 + k)), seed=60 + k)
        c.standardize_data(); c.construct_kde(0.1)
        conts.append(c)
    m = 6000
    prop = pd.DataFrame(tab.values[rs.randint(0, n, m)] + 0.15 * rs.normal(size=(m, 4)), columns=tab.columns)
    u = rs.uniform(size=m)
    out = emulator.score_group(conts, colsets, prop, sign=[1, -1, -1], log_m=np.log(20.0), u=u, band=2e-4)
    refs = []
    for c, cs in zip(conts, colsets):
        z = np.asarray(c._data)
        zq = np.asarray(c.pca_transform(prop[cs]))
        refs.append(brute64.score(z, np.asarray(c.weights), 0.1, zq))
    for lp, ref in zip(out["logp"], refs):
        near = ref >= ref.max() - 100
        assert np.abs(lp - ref)[near].max() <= 1e-4
    lr = oaccept.log_ratio(refs, [np.log(abs(c._jacobian_det)) for c in conts], [1, -1, -1], np.log(20.0))
    assert np.abs(out["log_ratio"] - lr).max() <= 3e-4
    gold = np.log(u) < lr
    band = np.abs(lr - np.log(u)) <= 4e-4
    assert np.array_equal(out["accept"][~band], gold[~band])
    assert out["in_band"].sum() <= band.sum() + 5 and 0 < gold.sum() < m
    # and the probability-space form the reference uses gives the same decisions
    ref_mask, _ = oaccept.accept(refs[1], conts[1]._jacobian_det, refs[2], conts[2]._jacobian_det,
                                 refs[0], conts[0]._jacobian_det, m_factor=20, uniform=u)
    assert np.array_equal(out["accept"][~band], np.asarray(ref_mask)[~band])
    for c in conts:
        c.drop_kde()


def test_concentric_shells_batched_equals_per_shell_reference_flow():
    """Config-3 shape scaled down: several shells, 4 KDEs each, one batched launch per dimension; the
    scores must equal what calc_scores2 gives shell by shell and sit within 1e-4 of brute64."""
    from skysampler_b200 import KDEContainer, concentric, synth
    from oracle import brute64
    rs = np.random.RandomState(7)
    columns = {"cols_dc": ["C0", "C1", "C2"], "cols_wr": ["LOGR"], "cols_wcr": ["C0", "C1", "C2", "LOGR"]}
    shells = []
    for s in range(3):
        n = 3000 + 500 * s
        smc = synth.mixture(4, n, 700 + s)
        wcr = synth.mixture(4, n, 710 + s); wcr[:, 3] = 0.3 * wcr[:, 3] - 0.2 + 0.1 * s
        mk = lambda arr, cols, w, seed: {"container": KDEContainer(pd.DataFrame(arr, columns=cols), w, seed=seed)}
        shells.append({
            "deep_smc": mk(smc, ["MAG", "C0", "C1", "C2"], None, 1 + s),
            "deep_c": mk(synth.mixture(4, n, 720 + s)[:, 1:], ["C0", "C1", "C2"], None, 11 + s),
            "wide_cr_clust": mk(wcr, ["C0", "C1", "C2", "LOGR"], pd.Series(synth.wide_weights(n, 730 + s)), 21 + s),
            "wide_cr_rands": mk(wcr[::-1].copy(), ["C0", "C1", "C2", "LOGR"], pd.Series(synth.wide_weights(n, 740 + s)), 31 + s),
            "wide_r_ref": mk(0.3 * synth.mixture(1, n, 750 + s) - 0.2 + 0.1 * s, ["LOGR"], pd.Series(synth.wide_weights(n, 760 + s)), 41 + s),
            "rmin": -0.5 + 0.1 * s, "rmax": 0.1 + 0.1 * s,
        })
    out = concentric.run_concentric(shells, columns, nsamples=2000, bandwidth=0.1, philox=True)
    assert len(out) == 3
    for sh, (samples, scores) in zip(shells, out):
        assert list(scores.columns) == ["dc", "dc_jac", "wr", "wr_jac", "wcr_clust", "wcr_clust_jac", "wcr_rands", "wcr_rands_jac"]
        assert len(samples) == 2000 and samples["LOGR"].between(sh["rmin"], sh["rmax"]).all()
        assert np.all(np.isfinite(scores.values))
        # oracle check of one KDE per shell with the whitening the run used (containers keep it)
        c = sh["wide_cr_clust"]["container"]
        mu, W, Winv = None, None, None
        zq = (samples[columns["cols_wcr"]].values - (c.mean1 + c.pca_params["pca"].mean_)) @ (c.pca_params["pca"].components_.T / c.std2[None, :])
        ref = brute64.score(np.asarray(c._data), np.asarray(c.weights), 0.1, zq)
        near = ref >= ref.max() - 100
        assert np.abs(scores["wcr_clust"].values - ref)[near].max() <= 1e-4


def test_kfold_cv_bandwidth_scores():
    """emulator.py:397-490 on the backend: CV score per bandwidth equals the oracle's weighted mean log-likelihood."""
    from skysampler_b200 import KDEContainer, emulator, synth
    from oracle import brute64
    df, w = synth.frame(3, 4000, 77)
    cont = KDEContainer(df, w, seed=5)
    cv = emulator.KFoldCV(cont, nfold=3, nprocess=7, seed=9)
    res = cv._loop_bandwidths([0.1, 0.3])
    assert np.shape(res) == (2, 3)
    z = np.asarray(cont._data); ww = np.asarray(cont.weights)
    for ib, bw in enumerate((0.1, 0.3)):
        for fold in range(3):
            tr = cv.labels != fold; te = cv.labels == fold
            lp = brute64.score(z[tr], ww[tr], bw, z[te])
            ref = np.sum(ww[te] * (lp + np.log(abs(cont._jacobian_det)))) / np.sum(ww[te])
            assert abs(res[ib][fold] - ref) <= 1e-4 * max(1.0, abs(ref))
    # the reference's iparts[0] behaviour is reproducible on request
    cvb = emulator.KFoldCV(KDEContainer(df, w, seed=5), nfold=3, nprocess=7, seed=9, reference_bug_compat=True)
    rb = cvb._loop_cv(0.3)
    te = np.nonzero(cvb.labels == 0)[0]
    first = emulator.partition(np.arange(len(te)), 7)[0]
    zz = np.asarray(cvb.cont._data); tr = cvb.labels != 0
    lp = brute64.score(zz[tr], ww[tr], 0.3, zz[te][first])
    ref = np.sum(ww[te][first] * (lp + np.log(abs(cvb.cont._jacobian_det)))) / np.sum(ww[te][first])
    assert abs(rb[0] - ref) <= 1e-4 * max(1.0, abs(ref))


def test_group_and_single_paths_give_the_same_bits():
    """score_samples (register epilogue) and the fused group path (partial state + last-arriver
    epilogue) must agree bit for bit — smoke() relies on it."""
    from skysampler_b200 import KDEContainer, emulator, synth
    df, w = synth.frame(4, 10000, 101)
    c = KDEContainer(df, w, seed=1)
    c.standardize_data(); c.construct_kde(0.1)
    prop = c.random_draw(20000)
    lp, _ = c.score_samples(prop)
    out = emulator.score_group([c], [list(df.columns)], prop, sign=[1], log_m=0.0, u=np.random.RandomState(1).uniform(size=len(prop)))
    assert np.array_equal(out["logp"][0], lp)
    c.drop_kde()


def test_rejection_sampler_script_end_to_end(tmp_path):
    """bin/rejection_sampler_b200.py: feature tables -> containers -> proposals -> scores -> FITS -> accept."""
    import json, os, subprocess, sys
    from conftest import ROOT
    from skysampler_b200 import fitsio_lite
    out = str(tmp_path / "delta")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "rejection_sampler_b200.py"), "--ndeep", "20000", "--nwide", "40000",
                        "--nsamples", "30000", "--nchunks", "4", "--out", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["proposals"] == 30000 and 0 < info["accepted"] < 30000
    sc = fitsio_lite.read(out + "_scores.fits")
    sm = fitsio_lite.read(out + "_samples.fits")
    assert list(sc.dtype.names)[1:] == ["dc", "dc_jac", "wr", "wr_jac", "wcr", "wcr_jac"] and len(sm) == 30000
    assert sm["LOGR"].max() <= np.log10(5.0) + 1e-12
    assert np.all(np.isfinite(sc["wcr"])) and np.all(np.isfinite(sc["dc"])) and np.all(np.isfinite(sc["wr"]))


def test_pipelined_host_batches_equal_resident_steps():
    """RatioSampler.run_host (H2D on a side stream, two slots) gives the flags and counts of step()."""
    import torch
    from skysampler_b200 import GaussianKDE
    from skysampler_b200.pipeline import RatioSampler
    rs = np.random.RandomState(61)
    za = rs.normal(size=(30000, 4)); zb = 1.2 * rs.normal(size=(30000, 4))
    ka = GaussianKDE(0.2).fit(za); kb = GaussianKDE(0.2).fit(zb)
    s = RatioSampler([ka, kb], [1, -1], log_m=0.5, cols=None)
    s.kdes[0]  # keep
    M = 20000
    qs = [torch.from_numpy(1.2 * rs.normal(size=(M, 4))).pin_memory() for _ in range(5)]
    us = [torch.from_numpy(rs.uniform(size=M)).pin_memory() for _ in range(5)]
    # handles were fitted on whitened rows without a map: score in whitened space through the group API needs W
    ka.close(); kb.close()
    I = np.eye(4)
    ka = GaussianKDE(0.2).fit(za, whitening=(np.zeros(4), I, I)); kb = GaussianKDE(0.2).fit(zb, whitening=(np.zeros(4), I, I))
    s = RatioSampler([ka, kb], [1, -1], log_m=0.5)
    flags, counts = s.run_host(zip(qs, us))
    for i in range(5):
        acc, _, _, _ = s.score_accept(qs[i].cuda(), us[i].cuda())
        assert np.array_equal(flags[i].numpy(), acc.cpu().numpy())
        assert counts[i] == int(acc.sum().item())
        rows, n = s.step(qs[i].cuda(), us[i].cuda())
        assert n == counts[i] and np.array_equal(rows.cpu().numpy(), qs[i].numpy()[flags[i].numpy().astype(bool)])
    ka.close(); kb.close()


def test_fused_ratio_accept_mixes_both_score_kernels():
    """A KDE group with a 6-D member (per-query-culling kernel) and 2-D / 1-D members (lanes = queries kernel):
    the launches share the visit order and the arrival counters; whichever CTA arrives last finalises."""
    from skysampler_b200 import emulator, KDEContainer, synth
    from oracle import brute64
    rs = np.random.RandomState(5)
    n = 6000
    names = ["C0", "C1", "C2", "C3", "C4", "LOGR"]
    tab = pd.DataFrame(np.column_stack([synth.mixture(5, n, 43), 0.3 * rs.normal(size=n)]), columns=names)
    conts, colsets = [], [names, ["C0", "C1"], ["LOGR"]]
    for k, cs in enumerate(colsets):
        x = tab[cs].values + 0.1 * rs.normal(size=(n, len(cs)))
        c = KDEContainer(pd.DataFrame(x, columns=cs), pd.Series(synth.wide_weights(n, 70 + k)), seed=80 + k)
        c.standardize_data(); c.construct_kde(0.15)
        conts.append(c)
    m = 7001
    prop = pd.DataFrame(tab.values[rs.randint(0, n, m)] + 0.15 * rs.normal(size=(m, 6)), columns=names)
    u = rs.uniform(size=m)
    out = emulator.score_group(conts, colsets, prop, sign=[1, -1, -1], log_m=np.log(20.0), u=u, band=2e-4)
    refs, mismatch = [], []
    for c, cs, lp in zip(conts, colsets, out["logp"]):
        ref = brute64.score(np.asarray(c._data), np.asarray(c.weights), 0.15, np.asarray(c.pca_transform(prop[cs])))
        near = ref >= ref.max() - 100
        assert np.abs(lp - ref)[near].max() <= 1e-4
        single, _ = c.score_samples(prop[cs])
        if not np.array_equal(single, lp):                     # group and single paths: same bits
            mismatch.append((len(cs), int((single != lp).sum()), float(np.abs(single - lp).max())))
        refs.append(ref)
    lr = oaccept.log_ratio(refs, [np.log(abs(c._jacobian_det)) for c in conts], [1, -1, -1], np.log(20.0))
    assert np.abs(out["log_ratio"] - lr).max() <= 3e-4
    band = np.abs(lr - np.log(u)) <= 4e-4
    assert np.array_equal(out["accept"][~band], (np.log(u) < lr)[~band])
    again = emulator.score_group(conts, colsets, prop, sign=[1, -1, -1], log_m=np.log(20.0), u=u, band=2e-4)
    assert np.array_equal(again["log_ratio"], out["log_ratio"])    # counters self-reset
    for c in conts:
        c.drop_kde()
    assert not mismatch, "(dimension, rows that differ, max |d log p|) per member: %r" % (mismatch,)


def test_feature_construction_on_the_device_matches_the_reference_tables():
    """construct_features(device=0): column algebra, limit cuts and log10 in ONE kernel + the order-preserving compaction,
    against the reference's own feature tables (tests/golden/features.npz, emulator.py:74-132) and against the host
    mirror for float32 catalogues (numpy's promotion rules: float32 arithmetic, float64 storage)."""
    from conftest import load_golden
    from skysampler_b200 import features
    g = load_golden("features")
    nd = len(g["deep_MAG_G"])
    deep = np.zeros(nd, dtype=[("MAG_G", "f8"), ("MAG_R", "f8"), ("MAG_I", "f8"), ("MAG_Z", "f8"), ("SIZE", "f8"), ("E", "f8", (2,))])
    for n in ("MAG_G", "MAG_R", "MAG_I", "MAG_Z", "SIZE", "E"):
        deep[n] = g["deep_" + n]
    settings = {
        "columns": [("MAG_I", "MAG_I"), ("COLOR_G_R", ("MAG_G", "MAG_R", "-")), ("COLOR_R_I", ("MAG_R", "MAG_I", "-")),
                    ("LOGSIZE", "SIZE"), ("ELL", (("E", 0), ("E", 1), "SQSUM")), ("E1", ("E", 1))],
        "logs": [False, False, False, True, False, False],
        "limits": [(19, 24), (-1, 3), (-1, 3), (1e-3, 2.0), (0, 10), (-5, 5)],
    }
    dfc = features.DeepFeatureContainer(deep)
    dfc.construct_features(device=0, **settings)
    np.testing.assert_array_equal(np.asarray(dfc.features.index), g["deep_feature_index"])
    got, ref = dfc.features.values, g["deep_features"]
    logcol = 3
    for c in range(got.shape[1]):
        if c == logcol:
            np.testing.assert_allclose(got[:, c], ref[:, c], rtol=4e-16, atol=0)       # log10: within 2 ulp of numpy's
        else:
            np.testing.assert_array_equal(got[:, c], ref[:, c])
    np.testing.assert_array_equal(np.asarray(dfc.weights), g["deep_weights"])
    # float32 catalogue + literals + every operator: device == host mirror
    rs = np.random.RandomState(4)
    cat = np.zeros(20000, dtype=[("A", "f4"), ("B", "f4"), ("V", "f4", (3,)), ("W8", "f8")])
    cat["A"] = rs.uniform(0.5, 3, 20000); cat["B"] = rs.uniform(0.5, 3, 20000); cat["V"] = rs.normal(size=(20000, 3)); cat["W8"] = rs.uniform(1, 2, 20000)
    cols = [("S", ("A", "B", "+")), ("D", ("A", "B", "-")), ("P", ("A", "W8", "*")), ("Q", ("A", "B", "/")),
            ("R", (("V", 0), ("V", 2), "SQSUM")), ("K", ("A", 2.5, "*")), ("V1", ("V", 1)), ("LA", "A")]
    lims = [(1.2, 5.5), (-2, 2), (0.6, 5.5), (0.2, 5), (0.05, 3), (1.3, 7.4), (-2.5, 2.5), (0.5, 3)]
    logs = [False, False, True, False, True, False, False, True]
    host = features.DeepFeatureContainer(cat); host.construct_features(cols, limits=lims, logs=logs)
    devc = features.DeepFeatureContainer(cat); devc.construct_features(cols, limits=lims, logs=logs, device=0)
    np.testing.assert_array_equal(np.asarray(devc.features.index), np.asarray(host.features.index))
    assert 2000 < len(host.features) < 19000
    for c, lg in enumerate(logs):
        np.testing.assert_allclose(devc.features.values[:, c], host.features.values[:, c], rtol=4e-16 if lg else 0, atol=0)
    # and it can stay on the device for the fit
    feats, index = devc.features_on_device(cols, limits=lims, logs=logs, device=0)
    assert feats.is_cuda and feats.shape == (len(host.features), 8) and index.shape[0] == len(host.features)

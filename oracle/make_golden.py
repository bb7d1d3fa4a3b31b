"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python oracle/make_golden.py

``skysampler/emulator.py`` is imported from /root/reference with ``fitsio`` and ``matplotlib``
stubbed (both are imported at module level, emulator.py:11,34-39, and unused on the KDE path).
Every array written here is an OUTPUT OF THE REFERENCE'S OWN CODE (KDEContainer, make_naive_infodicts,
run_scores, read_concentric) or an input needed to replay it.
"""
import os
import sys
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def load_reference(path="/root/reference"):
    for m in ("fitsio", "matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, path)
    from skysampler import emulator  # noqa
    return emulator


def snapshot(c):
    return dict(mean1=np.asarray(c.mean1, dtype=np.float64), pca_mean=np.asarray(c.pca.mean_),
                components=np.asarray(c.pca.components_), std2=np.asarray(c.std2),
                jac=np.float64(c._jacobian_det))


def main():
    from skysampler_b200 import synth
    emu = load_reference()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "features":
        return features(emu)
    kde_path(emu, synth)
    features(emu)


def kde_path(emu, synth):

    # ---------------------------------------------------------------- C1-shaped scoring vectors
    d, n = 4, 10000
    df, w = synth.frame(d, n, 101, weighted=True)
    np.savez_compressed(os.path.join(OUT, "c1_train.npz"), data=df.values, weights=w.values,
                        columns=np.array(list(df.columns)))
    for weighted in (True, False):
        for h in (0.05, 0.1):
            c = emu.KDEContainer(df.copy(), w.copy() if weighted else None, seed=11)
            c.standardize_data()
            snap = snapshot(c)
            c.construct_kde(h)
            rng_state = c.rng.get_state()
            draws = c.random_draw(1800)                       # reference draws (bulk queries)
            rs = np.random.RandomState(77)
            tail = df.values[rs.randint(0, n, 200)]
            center = np.asarray(c.mean1)
            tail = center + 2.0 * (tail - center) + 0.3 * rs.normal(size=tail.shape)
            q = np.vstack([draws.values, tail])
            qdf = pd.DataFrame(q, columns=df.columns)
            out = {}
            for name, (atol, rtol) in (("ref", (1e-6, 1e-6)), ("exact", (0.0, 0.0))):
                c._atol, c._rtol = atol, rtol
                c.construct_kde(h)
                lp, jac = c.score_samples(qdf)
                out["logp_" + name] = lp
            tag = "c1_score_%s_h%s" % ("w" if weighted else "u", str(h).replace(".", "p"))
            np.savez_compressed(os.path.join(OUT, tag + ".npz"), h=np.float64(h), weighted=np.bool_(weighted),
                                query=q, whitened_train_head=np.asarray(c._data)[:64], **snap, **out)
            print(tag, out["logp_exact"][:3], "jac", snap["jac"])

    # ---------------------------------------------------------------- draw stream vectors
    for weighted in (True, False):
        c = emu.KDEContainer(df.copy(), w.copy() if weighted else None, seed=7)
        c.standardize_data()
        snap = snapshot(c)
        h = 0.1
        c.construct_kde(h)
        st0 = c.rng.get_state()
        res_plain = c.random_draw(4000)
        st1 = c.rng.get_state()
        lo, hi = np.percentile(df["LOGR"].values, [20, 70])
        res_win = c.random_draw(4000, rmin=lo, rmax=hi)
        # replay the first call's streams to expose u, z and the centre indices the reference selected
        rs = np.random.RandomState()
        rs.set_state(st0)
        u = rs.uniform(0, 1, size=4000)
        z = rs.normal(0.0, 1.0, size=(4000, d))
        rs2 = np.random.RandomState()
        rs2.set_state(st0)
        _res = c.kde.sample(n_samples=4000, random_state=rs2)       # the reference's whitened draws
        tdata = np.asarray(c.kde.tree_.data)
        centres = _res - h * z
        # recover the index the reference picked: the training row equal to the centre
        from scipy.spatial import cKDTree
        dist, idx = cKDTree(tdata).query(centres)
        assert dist.max() < 1e-9, dist.max()
        tag = "draw_%s" % ("w" if weighted else "u")
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), h=np.float64(h), weighted=np.bool_(weighted),
                            u=u, z=z, idx=idx.astype(np.int64), res_whitened=_res,
                            res_plain=res_plain.values, res_window=res_win.values,
                            window=np.array([lo, hi]), seed=np.int64(7), **snap)
        print(tag, "idx[:5]", idx[:5], "window", lo, hi, res_win["LOGR"].min(), res_win["LOGR"].max())

    # ---------------------------------------------------------------- end-to-end driver slice
    # bin/delta_rejection_sampler_v01.py:95-110 on synthetic containers: make_naive_infodicts -> run_scores
    nn = 5000
    smc = synth.mixture(4, nn, 601)
    deep_smc_df = pd.DataFrame(smc, columns=["MAG", "C0", "C1", "C2"])
    deep_c_df = pd.DataFrame(synth.mixture(4, nn, 602)[:, 1:], columns=["C0", "C1", "C2"])
    wcr = synth.mixture(4, nn, 603)
    wcr[:, 3] = 0.35 * wcr[:, 3] - 0.2
    wide_cr_df = pd.DataFrame(wcr, columns=["C0", "C1", "C2", "LOGR"])
    wide_r_df = pd.DataFrame(0.35 * synth.mixture(1, nn, 604) - 0.2, columns=["LOGR"])
    w_cr = pd.Series(synth.wide_weights(nn, 605))
    w_r = pd.Series(synth.wide_weights(nn, 606))
    columns = {"cols_dc": ["C0", "C1", "C2"], "cols_wr": ["LOGR"], "cols_wcr": ["C0", "C1", "C2", "LOGR"]}
    seeds = [21, 22, 23, 24]
    gold = dict(deep_smc=deep_smc_df.values, deep_c=deep_c_df.values, wide_cr=wide_cr_df.values,
                wide_r=wide_r_df.values, w_cr=w_cr.values, w_r=w_r.values, seeds=np.array(seeds),
                nsamples=np.int64(3000), nchunks=np.int64(4), bandwidth=np.float64(0.1),
                rmax=np.float64(0.4))
    for name, tol in (("ref", 1e-6), ("exact", 0.0)):
        emu.KDEContainer._atol = tol
        emu.KDEContainer._rtol = tol
        deep_c = {"container": emu.KDEContainer(deep_c_df.copy(), None, seed=seeds[0])}
        deep_smc = {"container": emu.KDEContainer(deep_smc_df.copy(), None, seed=seeds[1])}
        wide_cr = {"container": emu.KDEContainer(wide_cr_df.copy(), w_cr.copy(), seed=seeds[2])}
        wide_r = {"container": emu.KDEContainer(wide_r_df.copy(), w_r.copy(), seed=seeds[3])}
        infodicts, samples = emu.make_naive_infodicts(wide_cr, wide_r, deep_c, deep_smc, columns, nsamples=3000,
                                                      nchunks=4, bandwidth=0.1, rmin=None, rmax=0.4)
        result = emu.run_scores(infodicts)
        gold["samples_" + name] = samples.values
        gold["sample_columns"] = np.array(list(samples.columns))
        gold["scores_" + name] = result.values
        gold["score_columns"] = np.array(list(result.columns))
        print("driver", name, result.iloc[:2].values)
    emu.KDEContainer._atol = 1e-6
    emu.KDEContainer._rtol = 1e-6

    # ---------------------------------------------------------------- accept test (read_concentric :717-748)
    scores = pd.DataFrame(gold["scores_exact"], columns=list(gold["score_columns"]))
    samples = pd.DataFrame(gold["samples_exact"], columns=list(gold["sample_columns"]))
    tables = {"x_scores.fits": scores.to_records(index=False), "x_samples.fits": samples.to_records(index=False)}
    emu.glob.glob = lambda expr: ["x_scores.fits"]
    emu.fio.read = lambda fname: tables[fname]
    for m_factor in (20, 2):
        resamples = emu.read_concentric("x_scores.fits", m_factor=m_factor, seed=6)
        mask = np.zeros(len(samples), dtype=bool)
        mask[np.asarray(resamples.index)] = True
        gold["accept_m%d" % m_factor] = mask
        gold["accepted_rows_m%d" % m_factor] = resamples.values
        print("accept m=%d:" % m_factor, mask.sum(), "of", len(mask))
    np.savez_compressed(os.path.join(OUT, "driver_naive.npz"), **gold)


def features(emu):
    # ---------------------------------------------------------------- feature tables (emulator.py:67-231)
    rs = np.random.RandomState(77)
    nd = 3000
    deep = np.zeros(nd, dtype=[("MAG_G", "f8"), ("MAG_R", "f8"), ("MAG_I", "f8"), ("MAG_Z", "f8"), ("SIZE", "f8"),
                               ("E", "f8", (2,))])
    for name, mu in (("MAG_G", 22.0), ("MAG_R", 21.5), ("MAG_I", 21.2), ("MAG_Z", 21.0)):
        deep[name] = rs.normal(mu, 1, nd)
    deep["SIZE"] = np.abs(rs.normal(0.5, 0.3, nd)) + 0.01
    deep["E"] = rs.normal(size=(nd, 2))
    deep_settings = {
        "columns": [("MAG_I", "MAG_I"), ("COLOR_G_R", ("MAG_G", "MAG_R", "-")), ("COLOR_R_I", ("MAG_R", "MAG_I", "-")),
                    ("LOGSIZE", "SIZE"), ("ELL", (("E", 0), ("E", 1), "SQSUM")), ("E1", ("E", 1))],
        "logs": [False, False, False, True, False, False],
        "limits": [(19, 24), (-1, 3), (-1, 3), (1e-3, 2.0), (0, 10), (-5, 5)],
    }
    dfc = emu.DeepFeatureContainer(deep)
    dfc.construct_features(**deep_settings)
    feat = dict(deep_MAG_G=deep["MAG_G"], deep_MAG_R=deep["MAG_R"], deep_MAG_I=deep["MAG_I"], deep_MAG_Z=deep["MAG_Z"],
                deep_SIZE=deep["SIZE"], deep_E=deep["E"], deep_features=dfc.features.values,
                deep_feature_index=np.asarray(dfc.features.index), deep_weights=np.asarray(dfc.weights))

    class Info(object):
        pass
    info = Info()
    nw = 6000
    parts = []
    for b in range(3):
        parts.append(pd.DataFrame({"MAG_R": rs.normal(21, 1, nw // 3), "MAG_I": rs.normal(20.6, 1, nw // 3),
                                   "DIST": 10 ** rs.uniform(-2.5, 1.0, nw // 3), "WEIGHT": rs.uniform(0.5, 2.0, nw // 3)}))
    # (an EMPTY radial bin would take the reference through np.array(self.samples)[valid_elements],
    #  emulator.py:156-158, which raises under numpy >= 1.24; the golden therefore has none)
    info.samples = parts
    info.redges = np.logspace(-2.5, 1.0, 8); info.rcens = np.sqrt(info.redges[1:] * info.redges[:-1])
    info.rareas = np.pi * (info.redges[1:] ** 2 - info.redges[:-1] ** 2)
    info.survey = None; info.numprof = None
    info.target = Info(); info.target.nrow = 50
    wide_settings = {"columns": [("COLOR_R_I", ("MAG_R", "MAG_I", "-")), ("LOGR", "DIST")], "logs": [False, True],
                     "limits": [(-1, 3), (1e-3, 16.0)]}
    fsc = emu.FeatureSpaceContainer(info)
    fsc.construct_features(**wide_settings)
    small = fsc.downsample(nmax=150, nbins=10, rng=np.random.RandomState(5))
    feat.update(wide_parts=np.vstack([p.values.astype(float) for p in parts if len(p)]), wide_part_sizes=np.array([len(p) for p in parts]),
                wide_redges=info.redges, wide_rareas=info.rareas, wide_features=fsc.features.values,
                wide_weights=np.asarray(fsc.weights), wide_surfdens=fsc.surfdens(icol=1),
                wide_down_data=small.data.values, wide_down_weights=np.asarray(small.weights))
    np.savez_compressed(os.path.join(OUT, "features.npz"), **feat)
    print("features: deep", dfc.features.shape, "wide", fsc.features.shape, "downsampled", small.data.shape)


if __name__ == "__main__":
    main()

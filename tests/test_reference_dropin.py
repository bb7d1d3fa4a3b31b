"""INTEGRATION.md "Option A": the reference's OWN module, with its KDEContainer swapped for the B200 one.

``/root/reference/skysampler/emulator.py`` is imported unmodified (``fitsio`` / ``matplotlib`` stubbed — unused on
this path) and its functions ``make_naive_infodicts`` / ``calc_scores`` / ``make_classifier_infodicts`` /
``calc_scores2`` are run from the reference's code against ``skysampler_b200.KDEContainer``, side by side with the
reference's own container, on the same seeds.  The whole file skips where ``/root/reference`` is absent (the GPU box).

* CPU variant: the device KDE behind the container is replaced by a recording stand-in built on the oracle
  (sklearn ``KernelDensity`` / ``oracle.draw``), so what is tested is the SURFACE — constructor, attributes, RNG
  consumption order, argument types the reference passes, DataFrame columns it gets back.
* ``-m gpu`` variant: the real backend; scores must agree with the reference's to the parity bar where the
  reference's own tolerance (atol = rtol = 1e-6, depth-first) allows a comparison.
"""
import os
import sys
import types

import numpy as np
import pandas as pd
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "skysampler")),
                                reason="the reference tree only exists in the build container")


def load_reference():
    for m in ("fitsio", "matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from skysampler import emulator
    return emulator


class RecordingKDE(object):
    """Plays ``skysampler_b200.kde.GaussianKDE`` without a GPU: same methods, oracle arithmetic, call log."""
    calls = []

    def __init__(self, bandwidth=1.0, kernel="gaussian", atol=0, rtol=0, breadth_first=True, device=None):
        self.bandwidth = float(bandwidth)
        RecordingKDE.calls.append(("init", kernel, atol, rtol, breadth_first))

    def fit(self, X, y=None, sample_weight=None, whitening=None, is_whitened=True):
        from sklearn.neighbors import KernelDensity
        X = np.asarray(X, dtype=np.float64)
        self.ndim_ = X.shape[1]
        self.z = X
        self.w = None if sample_weight is None else np.asarray(sample_weight, dtype=np.float64)
        self.mu, self.W, self.Winv = whitening
        self.sk = KernelDensity(bandwidth=self.bandwidth, atol=0, rtol=0, breadth_first=False).fit(X, sample_weight=self.w)
        RecordingKDE.calls.append(("fit", X.shape, sample_weight is not None, bool(is_whitened)))
        return self

    def score_samples(self, X, cols=None, is_whitened=True, out=None, stream=None):
        X = np.asarray(X, dtype=np.float64)
        RecordingKDE.calls.append(("score_samples", X.shape, bool(is_whitened)))
        z = X if is_whitened else (X - self.mu) @ self.W
        return self.sk.score_samples(z)

    def draw_streams(self, u, z, out_x, out_z, rows=None, out_idx=None):
        from oracle import draw
        RecordingKDE.calls.append(("draw_streams", len(u), rows is not None))
        zz, idx = draw.sample_from_streams(self.z, self.bandwidth, u, z, self.w)   # rng.normal(data[i], h): loc + scale * g
        x = zz @ self.Winv + self.mu
        r = np.arange(len(u)) if rows is None else rows
        out_z[r] = zz
        out_x[r] = x

    def close(self):
        RecordingKDE.calls.append(("close",))


def _tables(seed=0):
    """Small stand-ins for the catalogues of bin/delta_rejection_sampler_v01.py: deep_smc (4-D incl. colours),
    wide_r (1-D LOGR), deep_c (2-D), wide_cr (3-D)."""
    rs = np.random.RandomState(seed)
    n = 4000
    base = rs.normal(size=(n, 4)) @ (np.eye(4) + 0.3 * rs.normal(size=(4, 4)))
    logr = 0.4 * rs.normal(size=n)
    return {
        "deep_smc": (pd.DataFrame(base, columns=["MAG_I", "COLOR_G_R", "COLOR_R_I", "SIZE"]), None),
        "wide_r": (pd.DataFrame({"LOGR": logr}), pd.Series(rs.uniform(0.5, 2.0, n))),
        "deep_c": (pd.DataFrame(base[:, 1:3] + 0.05 * rs.normal(size=(n, 2)), columns=["COLOR_G_R", "COLOR_R_I"]), None),
        "wide_cr": (pd.DataFrame(np.column_stack([base[:, 1:3] + 0.1 * rs.normal(size=(n, 2)), logr]),
                                 columns=["COLOR_G_R", "COLOR_R_I", "LOGR"]), pd.Series(rs.uniform(0.5, 2.0, n))),
    }


COLUMNS = {"cols_dc": ["COLOR_G_R", "COLOR_R_I"], "cols_wr": ["LOGR"], "cols_wcr": ["COLOR_G_R", "COLOR_R_I", "LOGR"]}


def _settings(emu_module, cls):
    """Containers built THROUGH the reference module's name ``KDEContainer`` (as its drivers do)."""
    tabs = _tables()
    out = {}
    for k, (seed, key) in enumerate(((11, "deep_smc"), (12, "wide_r"), (13, "deep_c"), (14, "wide_cr"))):
        df, w = tabs[key]
        out[key] = {"container": cls(df.copy(), None if w is None else w.copy(), seed=seed)}
    return out


def _run_reference_flow(emu, cls):
    s = _settings(emu, cls)
    infodicts, samples = emu.make_naive_infodicts(s["wide_cr"], s["wide_r"], s["deep_c"], s["deep_smc"], COLUMNS,
                                                  nsamples=1500, nchunks=3, bandwidth=0.2, rmin=-0.5, rmax=0.6)
    scores = pd.concat([emu.calc_scores(info) for info in infodicts])
    return infodicts, samples, scores


def test_reference_functions_run_on_the_swapped_container_cpu(monkeypatch):
    emu = load_reference()
    import skysampler_b200.kde as bkde
    # the reference's own flow, untouched
    ref_infos, ref_samples, ref_scores = _run_reference_flow(emu, emu.KDEContainer)
    # the same flow with the class swapped in the reference module and the device KDE replaced by the recorder
    RecordingKDE.calls = []
    monkeypatch.setattr(bkde, "GaussianKDE", RecordingKDE)
    monkeypatch.setattr(emu, "KDEContainer", bkde.KDEContainer)
    infos, samples, scores = _run_reference_flow(emu, emu.KDEContainer)
    # RNG streams are consumed in the reference's order: the proposal table replays exactly
    assert list(samples.columns) == list(ref_samples.columns)
    assert np.abs(samples.values - ref_samples.values).max() <= 1e-9
    assert [len(i["sample"]) for i in infos] == [len(i["sample"]) for i in ref_infos]
    # same DataFrame surface: columns, jacobians (sign kept), scores where the reference's tolerance is tight
    assert list(scores.columns) == list(ref_scores.columns) == ["dc", "dc_jac", "wr", "wr_jac", "wcr", "wcr_jac"]
    for c in ("dc_jac", "wr_jac", "wcr_jac"):
        assert np.allclose(scores[c].values, ref_scores[c].values, rtol=1e-9, atol=0)
    for c in ("dc", "wr", "wcr"):
        bulk = ref_scores[c].values >= ref_scores[c].values.max() - 6.0
        assert np.abs(scores[c].values - ref_scores[c].values)[bulk].max() <= 2e-2     # the reference runs at 1e-6/1e-6
    kinds = [c[0] for c in RecordingKDE.calls]
    # make_naive_infodicts: 2 fits + draws; calc_scores: 3 fits + 3 score calls per chunk; every KDE dropped
    assert kinds.count("fit") == 2 + 3 * 3 and kinds.count("score_samples") == 3 * 3
    # make_naive_infodicts drops its two KDEs; calc_scores never drops (emulator.py:660-695: the worker just exits), so
    # a re-fit frees the previous device copy and the last three stay attached until drop_kde()
    assert kinds.count("close") == 2 + 3 * 2
    assert any(c[0] == "draw_streams" and c[2] for c in RecordingKDE.calls), "the window redraw loop was not exercised"
    for c in RecordingKDE.calls:
        if c[0] == "score_samples":
            assert c[2] is False                                   # raw-space rows: whitening is fused below the surface
    # the containers pickle after the flow (mp.Pool ships them to workers, emulator.py:698-712)
    import pickle
    for key in ("wide_cr", "wide_r", "deep_c"):
        infos[0][key]["container"].drop_kde()
        pickle.dumps(infos[0][key]["container"])


@pytest.mark.gpu
def test_reference_functions_run_on_the_b200_backend(monkeypatch):
    emu = load_reference()
    import skysampler_b200.kde as bkde
    ref_infos, ref_samples, ref_scores = _run_reference_flow(emu, emu.KDEContainer)
    monkeypatch.setattr(emu, "KDEContainer", bkde.KDEContainer)
    infos, samples, scores = _run_reference_flow(emu, emu.KDEContainer)
    assert np.abs(samples.values - ref_samples.values).max() <= 1e-9
    assert list(scores.columns) == list(ref_scores.columns)
    # exact oracle for the rows the reference scored: its own container at atol = rtol = 0 is only exact near the
    # peak (SURVEY finding 5), so compare against the fp64 brute force on the container's own whitened data
    from oracle import brute64
    s = _settings(emu, bkde.KDEContainer)
    for name, key, cols in (("dc", "deep_c", "cols_dc"), ("wr", "wide_r", "cols_wr"), ("wcr", "wide_cr", "cols_wcr")):
        bulk = ref_scores[name].values >= ref_scores[name].values.max() - 6.0
        assert np.abs(scores[name].values - ref_scores[name].values)[bulk].max() <= 2e-2
        cont = infos[0][key]["container"]                   # rng state after the flow: re-fit consumes the next draws
        cont.standardize_data()
        cont.construct_kde(0.2)
        tab = infos[0]["sample"][COLUMNS[cols]]
        lp, jac = cont.score_samples(tab)
        ref = brute64.score(np.asarray(cont._data), np.asarray(cont.weights), 0.2, np.asarray(cont.pca_transform(tab)))
        near = ref >= ref.max() - 100
        assert np.abs(lp - ref)[near].max() <= 1e-4
        cont.drop_kde()
    # the reference's run_scores forks one process per chunk; device handles are dropped before pickling, so the
    # in-process calc_scores loop above is the supported pattern

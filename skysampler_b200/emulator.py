"""Callers of the KDE path, mirrored from ``skysampler/emulator.py:525-748`` on the B200 backend.

Same names, arguments, return types and DataFrame columns as the reference:

* ``make_naive_infodicts`` / ``make_classifier_infodicts``  (emulator.py:626-657 / 525-559)
* ``calc_scores`` / ``calc_scores2``                        (emulator.py:660-695 / 562-606)
* ``run_scores`` / ``run_scores2``                          (emulator.py:698-712 / 609-623)
* ``accept_concentric`` — the accept test of ``read_concentric`` (emulator.py:737-746) on in-memory
  tables (the FITS reading around it is I/O and out of scope).

Differences, all below the surface:

* ``run_scores*`` do not fork one process per chunk (device handles are not fork-safe and the
  reference's 150 workers each re-fit PCA and rebuild a tree, emulator.py:668-689): every KDE is
  fitted ONCE per call and all chunks are scored in one fused pass per KDE group
  (``skb_ratio_accept``), which also yields the accept decision in the same launch when a uniform
  stream is supplied.  In the reference every worker unpickles the same container state, so all
  chunks see the same whitening; fitting once reproduces exactly that.
* ``calc_scores*`` keep the reference's per-chunk re-fit behaviour for callers that use them
  directly.
"""
import copy
import ctypes as C

import numpy as np
import pandas as pd

from . import _cabi
from ._cabi import check, ptr
from .kde import KDEContainer, _as_matrix


def partition(lst, n):
    """Divides a list into N roughly equal chunks (utils.py:142-185, same rounding rule)."""
    division = len(lst) / float(n)
    return [lst[int(round(division * i)): int(round(division * (i + 1)))] for i in range(n)]


# ------------------------------------------------------------------------------------------------
def _draw_proposals(deep_smc, wide_r, nsamples, bandwidth, rmin, rmax):
    """The shared head of make_*_infodicts (emulator.py:530-544 / 629-643)."""
    deep_smc_emu = deep_smc["container"]
    deep_smc_emu.standardize_data()
    deep_smc_emu.construct_kde(bandwidth)
    wide_r_emu = wide_r["container"]
    wide_r_emu.standardize_data()
    wide_r_emu.construct_kde(bandwidth)
    samples_smc = deep_smc_emu.random_draw(nsamples)
    samples_r = wide_r_emu.random_draw(nsamples, rmin=rmin, rmax=rmax)
    samples = pd.merge(samples_smc, samples_r, left_index=True, right_index=True)
    deep_smc_emu.drop_kde()
    wide_r_emu.drop_kde()
    return samples


def make_classifier_infodicts(wide_cr_clust, wide_r_ref, wide_cr_rands,
                              deep_c, deep_smc, columns,
                              nsamples=1e5, nchunks=1, bandwidth=0.1,
                              rmin=None, rmax=None, rcol="LOGR"):
    """emulator.py:525-559."""
    samples = _draw_proposals(deep_smc, wide_r_ref, nsamples, bandwidth, rmin, rmax)
    sample_inds = partition(list(samples.index), nchunks)
    infodicts = []
    for i in np.arange(nchunks):
        info = {
            "columns": columns,
            "bandwidth": bandwidth,
            "wide_cr_clust": wide_cr_clust,
            "wide_cr_rands": wide_cr_rands,
            "deep_c": deep_c,
            "wide_r_ref": wide_r_ref,
            "sample": samples.loc[sample_inds[i]],
            "rmin": rmin,
            "rmax": rmax,
        }
        infodicts.append(info)
    return infodicts, samples


def make_naive_infodicts(wide_cr, wide_r, deep_c, deep_smc, columns,
                         nsamples=1e5, nchunks=1, bandwidth=0.1,
                         rmin=None, rmax=None, rcol="LOGR"):
    """emulator.py:626-657."""
    samples = _draw_proposals(deep_smc, wide_r, nsamples, bandwidth, rmin, rmax)
    sample_inds = partition(list(samples.index), nchunks)
    infodicts = []
    for i in np.arange(nchunks):
        info = {
            "columns": columns,
            "bandwidth": bandwidth,
            "wide_cr": wide_cr,
            "deep_c": deep_c,
            "wide_r": wide_r,
            "sample": samples.loc[sample_inds[i]],
            "rmin": rmin,
            "rmax": rmax,
        }
        infodicts.append(info)
    return infodicts, samples


# ------------------------------------------------------------------------------------------------
_NAIVE = (("dc", "deep_c", "cols_dc"), ("wr", "wide_r", "cols_wr"), ("wcr", "wide_cr", "cols_wcr"))
_CLASSIFIER = (("dc", "deep_c", "cols_dc"), ("wr", "wide_r_ref", "cols_wr"),
               ("wcr_clust", "wide_cr_clust", "cols_wcr"), ("wcr_rands", "wide_cr_rands", "cols_wcr"))


def _calc(info, layout):
    scores = pd.DataFrame()
    try:
        columns = info["columns"]
        bandwidth = info["bandwidth"]
        sample = info["sample"]
        scores = pd.DataFrame()
        for name, key, colkey in layout:
            emu = info[key]["container"]
            emu.standardize_data()
            emu.construct_kde(bandwidth=bandwidth)
            _score, _jacobian = emu.score_samples(sample[columns[colkey]])
            scores[name] = _score
            scores[name + "_jac"] = _jacobian
            emu.drop_kde()
    except KeyboardInterrupt:
        pass
    return scores


def calc_scores(info):
    """emulator.py:660-695: columns dc, dc_jac, wr, wr_jac, wcr, wcr_jac."""
    return _calc(info, _NAIVE)


def calc_scores2(info):
    """emulator.py:562-606: columns dc, dc_jac, wr, wr_jac, wcr_clust, wcr_clust_jac, wcr_rands, wcr_rands_jac."""
    return _calc(info, _CLASSIFIER)


def score_group(containers, colsets, table, sign=None, log_m=0.0, u=None, band=2e-4):
    """One fused pass: every KDE in ``containers`` scores its column subset ``colsets[k]`` of the same
    proposal ``table`` (DataFrame); optionally forms the density ratio and the accept decision.

    Returns dict(logp=[K arrays], log_ratio, accept, in_band).  ``sign[k]`` is +1 for the reference
    density and -1 for proposal densities (emulator.py:743-746).
    """
    lib = _cabi.lib()
    K = len(containers)
    names = list(table.columns)
    X = _as_matrix(table, allow_torch=False)
    M, ld = X.shape
    cols = np.full((K, _cabi.SKB_MAXD), -1, dtype=np.int32)
    for k, cs in enumerate(colsets):
        idx = [names.index(c) for c in cs]
        if len(idx) != containers[k].ndim:
            raise ValueError("KDE %d expects %d features, got %d columns" % (k, containers[k].ndim, len(idx)))
        cols[k, :len(idx)] = idx
    handles = (C.c_void_p * K)(*[c.kde._handle() for c in containers])
    sg = np.ascontiguousarray(sign if sign is not None else np.ones(K), dtype=np.int32)
    ljac = np.ascontiguousarray([np.log(np.abs(c._jacobian_det)) for c in containers], dtype=np.float64)
    logp = np.empty((K, M), dtype=np.float64)
    log_ratio = np.empty(M, dtype=np.float64)
    accept = in_band = None
    uu = None
    if u is not None:
        uu = np.ascontiguousarray(u, dtype=np.float64)
        if uu.shape != (M,):
            raise ValueError("u must have one entry per proposal row")
        accept = np.empty(M, dtype=np.uint8)
        in_band = np.empty(M, dtype=np.uint8)
    check(lib.skb_ratio_accept(handles, K, ptr(sg), ptr(ljac), float(log_m), ptr(X), _cabi.dtype_code(X), M, ld,
                               ptr(cols), ptr(uu), float(band), ptr(logp), ptr(log_ratio), ptr(accept),
                               ptr(in_band), None))
    return {"logp": [logp[k] for k in range(K)], "log_ratio": log_ratio,
            "accept": None if accept is None else accept.astype(bool),
            "in_band": None if in_band is None else in_band.astype(bool)}


def _run(infodicts, layout):
    if len(infodicts) == 0:
        return pd.DataFrame()
    info0 = infodicts[0]
    columns, bandwidth = info0["columns"], info0["bandwidth"]
    table = pd.concat([info["sample"] for info in infodicts])
    conts, colsets = [], []
    rng_states = []
    for name, key, colkey in layout:
        emu = info0[key]["container"]
        # the reference re-fits inside forked workers, which leaves the CALLER's containers (and their RandomState)
        # untouched (emulator.py:698-712); fitting in-process must not advance the caller's stream either
        rng_states.append((emu, emu.rng.get_state()))
        emu.standardize_data()
        emu.construct_kde(bandwidth=bandwidth)
        conts.append(emu)
        colsets.append(columns[colkey])
    try:
        res = score_group(conts, colsets, table)
    finally:
        jacs = [c._jacobian_det for c in conts]
        for c in conts:
            c.drop_kde()
        for emu, state in rng_states:
            emu.rng.set_state(state)
    # per-chunk frames concatenated, like pd.concat(result) of the pool's outputs (index restarts per chunk)
    frames = []
    a = 0
    for info in infodicts:
        n = len(info["sample"])
        fr = pd.DataFrame()
        for k, (name, key, colkey) in enumerate(layout):
            fr[name] = res["logp"][k][a:a + n]
            fr[name + "_jac"] = jacs[k]
        frames.append(fr)
        a += n
    return pd.concat(frames)


def run_scores(infodicts):
    """emulator.py:698-712 without the process pool: one fit per KDE, one fused pass over all chunks."""
    return _run(infodicts, _NAIVE)


def run_scores2(infodicts):
    """emulator.py:609-623, same treatment."""
    return _run(infodicts, _CLASSIFIER)


def accept_concentric(scores, m_factor=20, seed=6, uniform=None):
    """The accept test of read_concentric (emulator.py:737-746) on an in-memory scores table.

    Host arithmetic on a table that already holds the log-densities (this is O(M) bookkeeping; the
    fused device version is ``score_group(..., u=...)``)."""
    dc_score = np.exp(scores["dc"]) * np.abs(scores["dc_jac"])
    wr_score = np.exp(scores["wr"]) * np.abs(scores["wr_jac"])
    wcr_score = np.exp(scores["wcr"]) * np.abs(scores["wcr_jac"])
    if uniform is None:
        rng = np.random.RandomState(seed)
        uniform = rng.uniform(0, 1, len(scores))
    p_proposal = m_factor * dc_score * wr_score
    p_ref = wcr_score
    inds = uniform < p_ref / p_proposal
    return np.asarray(inds)


# ------------------------------------------------------------------------------------------------
# Bandwidth cross-validation (emulator.py:397-490) — the third caller of construct_kde/score_samples
# ------------------------------------------------------------------------------------------------
class KFoldCV(object):
    """k-fold bandwidth CV on the B200 backend; same constructor / methods as emulator.py:397-458.

    The reference fans each test fold out to ``nprocess`` pool workers but hands EVERY worker the first
    partition (``iparts[0]`` at emulator.py:430-431), so its score is the weighted mean over the first
    1/nprocess of the fold.  ``reference_bug_compat=True`` reproduces that; the default scores the whole
    fold (the evident intent).  Either way one fold = one ``score_samples`` call here.
    """

    def __init__(self, cont, nfold=5, nprocess=100, seed=None, reference_bug_compat=False):
        self.cont = cont
        self.cont.standardize_data()
        self.nfold = nfold
        self.nprocess = nprocess
        self.seed = seed
        self.rng = np.random.RandomState(seed)
        self.labels = self.rng.randint(0, self.nfold, len(self.cont.data))
        self.reference_bug_compat = reference_bug_compat

    def _loop_bandwidths(self, bandwidths):
        """emulator.py:422-426.  The folds do not depend on the bandwidth (no random numbers are drawn in the loops), so a
        sweep splits and re-whitens each fold ONCE and only re-fits the device KDE per bandwidth (a few ms: the fit is a
        device kd ordering) — the reference re-slices the DataFrames and rebuilds its tree for every (bandwidth, fold)."""
        folds = [self._split(ifold) for ifold in np.arange(self.nfold)]
        scores = []
        for bw in bandwidths:
            scores.append(self._loop_cv(bw, folds=folds))
        return scores

    def _loop_cv(self, bandwidth, folds=None):
        mscores = []
        for ifold in np.arange(self.nfold):
            train, test = folds[ifold] if folds is not None else self._split(ifold)
            iarr = np.arange(len(test.data))
            iparts = partition(iarr, self.nprocess)
            if self.reference_bug_compat:
                sel = np.concatenate([iparts[0]] * len(iparts)) if len(iparts[0]) else iparts[0]
            else:
                sel = iarr
            info = {
                "train": train,
                "test_data": test.data.iloc[sel],
                "test_weights": pd.Series(np.asarray(test.weights))[sel] if not isinstance(test.weights, pd.Series)
                else test.weights.iloc[sel],
                "bandwidth": bandwidth,
            }
            result = run_cv_scores([info])
            # (the reference takes log(jac) of a determinant whose sign is arbitrary, emulator.py:437; the
            #  density needs |J| — as read_concentric uses it, :737-739)
            mscore = np.sum(result["weights"] * (result["scores"] + np.log(np.abs(result["jac"])))) / np.sum(result["weights"])
            mscores.append(mscore)
        return mscores

    def _split(self, label):
        """splits data into train and test k-fold (emulator.py:442-450)"""
        train = self._shrink(self.labels != label)
        test = self._shrink(self.labels == label)
        return train, test

    def _shrink(self, index):
        """emulator.py:452-458: shares the fitted whitening, re-whitens the subset."""
        _cont = copy.copy(self.cont)
        _cont.kde = None
        _cont.data = _cont.data.iloc[index].copy()
        if isinstance(_cont.weights, pd.Series):
            _cont.weights = _cont.weights.iloc[index].copy()
        else:
            _cont.weights = np.asarray(_cont.weights)[index].copy()
        _cont._data = _cont.pca_transform(_cont.data)
        return _cont


def _score_cv_samples(info):
    """emulator.py:478-490."""
    train = info["train"]
    train.construct_kde(bandwidth=info["bandwidth"])
    scores, jac = train.score_samples(info["test_data"])
    train.kde.close()
    train.kde = None
    result = pd.DataFrame()
    result["scores"] = scores
    result["jac"] = np.ones(len(result)) * jac
    result["weights"] = np.asarray(info["test_weights"])
    return result


def run_cv_scores(infodicts):
    """emulator.py:461-475, in-process."""
    return pd.concat([_score_cv_samples(info) for info in infodicts])


# ------------------------------------------------------------------------------------------------
# On-disk side of the accept step (emulator.py:717-748)
# ------------------------------------------------------------------------------------------------
def _fits():
    try:
        import fitsio as fio                      # the reference's reader / writer when it is installed
        return fio
    except Exception:
        from . import fitsio_lite as fio          # same table layout, see the module header
        return fio


def read_concentric(score_path_expr, m_factor=20, seed=6):
    """emulator.py:717-748: glob the ``*_scores.fits`` files, load the matching ``*_samples.fits``,
    apply the accept test, return the accepted proposal rows (same DataFrame the reference returns)."""
    import glob
    fio = _fits()
    fname_scores = np.sort(glob.glob(score_path_expr))
    fname_samples = [f.replace("scores", "samples") for f in fname_scores]
    samples = pd.concat([pd.DataFrame.from_records(fio.read(f)) for f in fname_samples])
    scores = pd.concat([pd.DataFrame.from_records(fio.read(f)) for f in fname_scores])
    inds = accept_concentric(scores, m_factor=m_factor, seed=seed)
    return samples[inds]


def write_tables(outname, samples, scores):
    """The two per-shell files of the drivers: ``<outname>_samples.fits`` and ``<outname>_scores.fits``
    (bin/epsilon_concentric_sampler_v05.py:162-164, 179-181), written from ``DataFrame.to_records()``."""
    fio = _fits()
    fio.write(outname + "_samples.fits", samples.to_records(), clobber=True)
    fio.write(outname + "_scores.fits", scores.to_records(), clobber=True)

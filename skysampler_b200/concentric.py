"""Concentric-shell line-of-sight sampler loop on the B200 backend.

Mirrors the inner loop of ``bin/epsilon_concentric_sampler_v05.py:127-181`` (and
``bin/delta_concentric_sampler_v01.py:79-136``): for every radial shell, draw proposals from the deep
(size/magnitude/colour) KDE and the wide radial KDE inside the shell's LOGR window, then score them
against the shell's KDEs.  The reference runs the shells one after the other, each through a 150-process
pool; here all shells are fitted first and scored as ONE batched launch per feature dimension
(``skb_score_batch``: blockIdx.y = (shell, KDE) job) — BASELINE.json config 3.
"""
import ctypes as C

import numpy as np
import pandas as pd
import torch

from . import _cabi
from ._cabi import check


def draw_shell_proposals(deep_smc, wide_r, nsamples, bandwidth, rmin, rmax, philox=True):
    """Proposal table of one shell: deep_smc draws merged with windowed radial draws
    (make_classifier_infodicts head, emulator.py:530-544)."""
    for c in (deep_smc, wide_r):
        if philox:
            c._rng_mode = "philox"
        c.standardize_data()
        c.construct_kde(bandwidth)
    samples_smc = deep_smc.random_draw(nsamples)
    samples_r = wide_r.random_draw(nsamples, rmin=rmin, rmax=rmax)
    samples = pd.merge(samples_smc, samples_r, left_index=True, right_index=True)
    deep_smc.drop_kde()
    wide_r.drop_kde()
    return samples


def score_shells_batched(jobs, device=None):
    """jobs: list of (container with a fitted device KDE, proposal DataFrame, column names).
    All jobs are scored in as few launches as SKB_MAX_JOBS allows; returns a list of float64 arrays."""
    lib = _cabi.lib()
    dev = torch.device("cuda", jobs[0][0].kde.device_ if device is None else device)
    outs = []
    st = _cabi.stream_ptr(torch.cuda.current_stream(dev))
    keep = []
    for a in range(0, len(jobs), _cabi.SKB_MAX_JOBS):
        chunk = jobs[a:a + _cabi.SKB_MAX_JOBS]
        arr = (_cabi.SkbJob * len(chunk))()
        for j, (cont, table, cols) in enumerate(chunk):
            names = list(table.columns)
            q = torch.from_numpy(np.ascontiguousarray(table.values, dtype=np.float64)).to(dev)
            ci = np.ascontiguousarray([names.index(c) for c in cols], dtype=np.int32)
            if len(ci) != cont.ndim:
                raise ValueError("job %d: KDE expects %d features, got %d columns" % (a + j, cont.ndim, len(ci)))
            out = torch.empty(len(table), dtype=torch.float64, device=dev)
            keep.append((q, ci, out))
            arr[j].kde = cont.kde._handle()
            arr[j].q = q.data_ptr()
            arr[j].q_dtype = _cabi.SKB_F64
            arr[j].q_is_whitened = 0
            arr[j].M = q.shape[0]
            arr[j].ldq = q.shape[1]
            arr[j].cols = ci.ctypes.data
            arr[j].out_logp = out.data_ptr()
            outs.append(out)
        check(lib.skb_score_batch(arr, len(chunk), st))
    torch.cuda.current_stream(dev).synchronize()
    return [o.cpu().numpy() for o in outs]


def run_concentric(shells, columns, nsamples=1e5, bandwidth=0.1, philox=True):
    """shells: list of dicts with containers under the reference's keys
    (wide_cr_clust, wide_r_ref, wide_cr_rands, deep_c, deep_smc — each {"container": KDEContainer})
    plus "rmin"/"rmax".  Returns [(samples DataFrame, scores DataFrame)] per shell, scores with the
    columns of calc_scores2 (emulator.py:562-606)."""
    layout = (("dc", "deep_c", "cols_dc"), ("wr", "wide_r_ref", "cols_wr"),
              ("wcr_clust", "wide_cr_clust", "cols_wcr"), ("wcr_rands", "wide_cr_rands", "cols_wcr"))
    tables = []
    for sh in shells:
        tables.append(draw_shell_proposals(sh["deep_smc"]["container"], sh["wide_r_ref"]["container"],
                                           int(nsamples), bandwidth, sh.get("rmin"), sh.get("rmax"), philox=philox))
    jobs, index = [], []
    for s, (sh, tab) in enumerate(zip(shells, tables)):
        for name, key, colkey in layout:
            cont = sh[key]["container"]
            cont.standardize_data()
            cont.construct_kde(bandwidth)
            jobs.append((cont, tab, columns[colkey]))
            index.append((s, name, cont))
    lps = score_shells_batched(jobs)
    results = [pd.DataFrame() for _ in shells]
    for (s, name, cont), lp in zip(index, lps):
        results[s][name] = lp
        results[s][name + "_jac"] = cont._jacobian_det
    for _, _, cont in index:
        cont.drop_kde()
    return list(zip(tables, results))

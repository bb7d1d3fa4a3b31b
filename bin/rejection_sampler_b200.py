#!/usr/bin/env python
"""Single-shot rejection sampler on synthetic catalogues — the flow of bin/delta_rejection_sampler_v01.py:83-111
on the B200 backend: feature tables -> KDE containers -> proposals -> scores -> files -> accept.

    construct_deep_container x2, construct_wide_container x2          (:95-98)
    make_naive_infodicts(nsamples, nchunks, bandwidth, rmax)           (:100-102)
    <out>_samples.fits, run_scores, <out>_scores.fits                  (:105-111)
    read_concentric(<out>*_scores.fits)                                (emulator.py:717-748)

The reference reads DES catalogues (a FITS deep field and a pickled indexer output); this script builds
synthetic stand-ins with the same columns so the whole chain can run anywhere.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def synthetic_inputs(n_deep, n_wide, seed):
    rs = np.random.RandomState(seed)
    base = rs.normal(21.0, 1.0, n_deep)
    deep = np.zeros(n_deep, dtype=[("MAG_G", "f8"), ("MAG_R", "f8"), ("MAG_I", "f8"), ("MAG_Z", "f8"), ("SIZE", "f8")])
    deep["MAG_I"] = base
    deep["MAG_R"] = base + rs.normal(0.4, 0.25, n_deep)
    deep["MAG_G"] = deep["MAG_R"] + rs.normal(0.9, 0.35, n_deep)
    deep["MAG_Z"] = base - rs.normal(0.3, 0.2, n_deep)
    deep["SIZE"] = 10 ** rs.normal(-0.3, 0.25, n_deep)

    class Info(object):
        pass
    info = Info()
    nb = 8
    info.redges = np.logspace(-2.5, 1.2, nb + 1)
    info.rcens = np.sqrt(info.redges[1:] * info.redges[:-1])
    info.rareas = np.pi * (info.redges[1:] ** 2 - info.redges[:-1] ** 2)
    info.survey = None; info.numprof = None
    info.target = Info(); info.target.nrow = 100
    parts = []
    for b in range(nb):
        m = n_wide // nb
        mi = rs.normal(20.8, 1.0, m)
        red = rs.rand(m) < (0.6 - 0.05 * b)                      # a red-sequence excess towards the centre
        mr = mi + np.where(red, rs.normal(0.55, 0.08, m), rs.normal(0.35, 0.3, m))
        mg = mr + np.where(red, rs.normal(1.25, 0.1, m), rs.normal(0.8, 0.4, m))
        mz = mi - rs.normal(0.3, 0.2, m)
        parts.append(pd.DataFrame({"MOF_CM_MAG_CORRECTED_G": mg, "MOF_CM_MAG_CORRECTED_R": mr, "MOF_CM_MAG_CORRECTED_I": mi,
                                   "MOF_CM_MAG_CORRECTED_Z": mz, "DIST": rs.uniform(info.redges[b], info.redges[b + 1], m),
                                   "WEIGHT": np.full(m, info.rareas[b] / info.rareas[0] * 0 + rs.uniform(0.5, 2.0))}))
    info.samples = parts
    return deep, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ndeep", type=int, default=200000)
    ap.add_argument("--nwide", type=int, default=400000)
    ap.add_argument("--nsamples", type=int, default=1000000)
    ap.add_argument("--nchunks", type=int, default=32)
    ap.add_argument("--bandwidth", type=float, default=0.05)
    ap.add_argument("--out", default="/tmp/skysampler_b200_delta")
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--rng", default="philox", choices=["philox", "host"],
                    help="philox: device counter RNG, window redraws inside the kernel; host: consume numpy "
                         "RandomState streams exactly as the reference does")
    args = ap.parse_args()
    from skysampler_b200 import emulator, features, _cabi
    deep, mdl = synthetic_inputs(args.ndeep, args.nwide, args.seed)
    catalog_rmax, draw_rmax = 16.0, 5.0
    deep_c_settings = {"columns": [("COLOR_G_R", ("MAG_G", "MAG_R", "-")), ("COLOR_R_I", ("MAG_R", "MAG_I", "-")),
                                   ("COLOR_I_Z", ("MAG_I", "MAG_Z", "-"))], "logs": [False] * 3, "limits": [(-1, 3)] * 3}
    deep_smc_settings = {"columns": [("MAG_I", "MAG_I"), ("LOGSIZE", "SIZE"), ("COLOR_G_R", ("MAG_G", "MAG_R", "-")),
                                     ("COLOR_R_I", ("MAG_R", "MAG_I", "-")), ("COLOR_I_Z", ("MAG_I", "MAG_Z", "-"))],
                         "logs": [False, True, False, False, False], "limits": [(17, 25), (1e-3, 10), (-1, 3), (-1, 3), (-1, 3)]}
    wide_cr_settings = {"columns": [("COLOR_G_R", ("MOF_CM_MAG_CORRECTED_G", "MOF_CM_MAG_CORRECTED_R", "-")),
                                    ("COLOR_R_I", ("MOF_CM_MAG_CORRECTED_R", "MOF_CM_MAG_CORRECTED_I", "-")),
                                    ("COLOR_I_Z", ("MOF_CM_MAG_CORRECTED_I", "MOF_CM_MAG_CORRECTED_Z", "-")), ("LOGR", "DIST")],
                        "logs": [False, False, False, True], "limits": [(-1, 3), (-1, 3), (-1, 3), (1e-3, catalog_rmax)]}
    wide_r_settings = {"columns": [("LOGR", "DIST")], "logs": [True], "limits": [(1e-3, catalog_rmax)]}
    columns = {"cols_dc": ["COLOR_G_R", "COLOR_R_I", "COLOR_I_Z"], "cols_wr": ["LOGR"],
               "cols_wcr": ["COLOR_G_R", "COLOR_R_I", "COLOR_I_Z", "LOGR"]}
    seeds = np.random.RandomState(args.seed).randint(0, np.iinfo(np.int32).max, 4)
    t0 = time.time()
    deep_c = features.construct_deep_container(deep, deep_c_settings, seed=seeds[0])
    deep_smc = features.construct_deep_container(deep, deep_smc_settings, seed=seeds[1])
    wide_cr = features.construct_wide_container(mdl, wide_cr_settings, seed=seeds[2])
    wide_r = features.construct_wide_container(mdl, wide_r_settings, seed=seeds[3])
    for c in (deep_smc, wide_r):
        c["container"]._rng_mode = args.rng
    t1 = time.time()
    infodicts, samples = emulator.make_naive_infodicts(wide_cr, wide_r, deep_c, deep_smc, columns, nsamples=args.nsamples,
                                                       nchunks=args.nchunks, bandwidth=args.bandwidth, rmin=None,
                                                       rmax=np.log10(draw_rmax))
    t2 = time.time()
    result = emulator.run_scores(infodicts)
    t3 = time.time()
    emulator.write_tables(args.out, samples, result)
    accepted = emulator.read_concentric(args.out + "*_scores.fits", m_factor=20, seed=6)
    t4 = time.time()
    pairs = float(len(samples)) * sum(len(c["container"].data) for c in (deep_c, wide_r, wide_cr))
    print(json.dumps({"proposals": len(samples), "accepted": len(accepted), "kde_rows": {"dc": len(deep_c["container"].data),
                      "wr": len(wide_r["container"].data), "wcr": len(wide_cr["container"].data)}, "pairs": pairs,
                      "seconds": {"containers": t1 - t0, "proposals": t2 - t1, "scores": t3 - t2, "files_and_accept": t4 - t3},
                      "pairs_per_s_scores_wall": pairs / (t3 - t2), "launches": int(_cabi.lib().skb_launch_count()),
                      "files": [args.out + "_samples.fits", args.out + "_scores.fits"]}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Concentric-shell sampler on synthetic containers (BASELINE.json config 3): the loop of
bin/epsilon_concentric_sampler_v05.py:127-181 with every shell fitted up front and all shell KDEs
scored as one batched launch.  Prints a JSON summary; writes <out>_rbin{i}_samples.fits / _scores.fits when
--out is given (through fitsio when installed, else skysampler_b200.fitsio_lite: same BINTABLE layout).

    python bin/concentric_sampler_b200.py --shells 10 --ntrain 100000 --nsamples 1000000
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_shell(s, n, seed0):
    from skysampler_b200 import KDEContainer, synth
    cols_smc = ["MAG", "C0", "C1", "C2"]
    cols_cr = ["C0", "C1", "C2", "LOGR"]
    mk = lambda arr, cols, w, seed: {"container": KDEContainer(pd.DataFrame(arr, columns=cols), w, seed=seed)}
    smc = synth.mixture(4, n, seed0 + 10 * s)
    wcr = synth.mixture(4, n, seed0 + 10 * s + 1)
    lo, hi = -1.0 + 0.15 * s, -0.85 + 0.15 * s                       # this shell's LOGR window
    wcr[:, 3] = 0.5 * (lo + hi) + 0.2 * (wcr[:, 3] - wcr[:, 3].mean()) / wcr[:, 3].std()
    r = 0.5 * (lo + hi) + 0.2 * synth.mixture(1, n, seed0 + 10 * s + 2)
    return {
        "deep_smc": mk(smc, cols_smc, None, seed0 + 10 * s + 3),
        "deep_c": mk(synth.mixture(4, n, seed0 + 10 * s + 4)[:, 1:], cols_cr[:3], None, seed0 + 10 * s + 5),
        "wide_cr_clust": mk(wcr, cols_cr, pd.Series(synth.wide_weights(n, seed0 + 10 * s + 6)), seed0 + 10 * s + 7),
        "wide_cr_rands": mk(wcr[::-1].copy(), cols_cr, pd.Series(synth.wide_weights(n, seed0 + 10 * s + 8)), seed0 + 10 * s + 9),
        "wide_r_ref": mk(r, ["LOGR"], pd.Series(synth.wide_weights(n, seed0 + 10 * s + 6)), seed0 + 10 * s + 7),
        "rmin": lo, "rmax": hi,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shells", type=int, default=10)
    ap.add_argument("--ntrain", type=int, default=100000)
    ap.add_argument("--nsamples", type=int, default=1000000)
    ap.add_argument("--bandwidth", type=float, default=0.1)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from skysampler_b200 import concentric, _cabi
    columns = {"cols_dc": ["C0", "C1", "C2"], "cols_wr": ["LOGR"], "cols_wcr": ["C0", "C1", "C2", "LOGR"]}
    shells = [make_shell(s, args.ntrain, 300) for s in range(args.shells)]
    l0 = _cabi.lib().skb_launch_count()
    t0 = time.time()
    res = concentric.run_concentric(shells, columns, nsamples=args.nsamples, bandwidth=args.bandwidth, philox=True)
    dt = time.time() - t0
    pairs = args.shells * 4 * float(args.ntrain) * args.nsamples
    summary = {"shells": args.shells, "ntrain_per_kde": args.ntrain, "proposals_per_shell": args.nsamples,
               "kdes_scored_per_shell": 4, "pairs": pairs, "wall_s": dt, "pairs_per_s_wall": pairs / dt,
               "kernel_launches": int(_cabi.lib().skb_launch_count() - l0),
               "note": "wall time includes host PCA fits, uploads, draws and pandas merges"}
    if args.out:
        from skysampler_b200 import emulator
        for i, (samples, scores) in enumerate(res):          # the reference's per-shell files (v05 driver :162-181)
            emulator.write_tables("%s_rbin%d" % (args.out, i), samples, scores)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()

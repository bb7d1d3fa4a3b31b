"""The five BASELINE.json configurations at (or near) full size on one GPU: throughput + parity samples.

    python tools/run_configs.py [--c4-queries 100000000] > profiles/rXX_configs.json

C1  4-D, 1e4 train, 1e5 queries, h=0.1, weighted        full parity vs the fp64 oracle
C2  8-D, 2 KDEs x 1e6 train, 1e7 candidates, ratio+accept   (bench.py is the timed version; here one pass over the pool)
C3  10 shell KDEs as a batch, 1e5 train / 1e6 queries each, 4-D
C4  16-D, 2e6 train, 1e8 queries (device-generated in chunks), h in {0.1}; parity on a 1e4 host subset... of chunk 0
C5  sweep N in {1e4, 1e5, 1e6, 4e6} x 2^20 queries, 8-D
"""
import argparse
import ctypes as C
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from skysampler_b200 import GaussianKDE, _cabi, synth
from oracle import brute64, cbrute

lib = _cabi.lib()
_cabi.set_option("stats", 1)


def whitened(d, n, seed):
    x = synth.mixture(d, n, seed)
    m1, pm, comps, std2 = synth.whiten_params(x)
    return (x - m1) @ comps.T / std2


def timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def parity(kde, z, w, h, q, n=2000, seed=0):
    sub = np.random.RandomState(seed).choice(len(q), min(n, len(q)), replace=False)
    got = kde.score_samples(q[sub])
    ref = cbrute.score(z, w, h, q[sub])
    near = ref >= ref.max() - 100
    return float(np.abs(got - ref)[near].max()), int(near.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c4-queries", type=int, default=100000000)
    args = ap.parse_args()
    out = {}
    st = (C.c_uint64 * 4)()

    # ---- C1 ----
    z = whitened(4, 10000, 101); w = synth.wide_weights(10000, 1); q = synth.queries(z, 0.1, 100000, 5)
    kde = GaussianKDE(0.1).fit(z, sample_weight=w)
    qd = torch.from_numpy(q).cuda(); o = kde.score_samples(qd)
    ms = timed(lambda: kde.score_samples(qd, out=o), 20)
    ref = brute64.score(z, w, 0.1, q, block=4000)
    near = ref >= ref.max() - 100
    out["C1"] = dict(d=4, n_train=10000, n_query=100000, ms=ms, pairs_per_s=1e9 / ms * 1e3,
                     max_abs_dlogp_within_100_nats=float(np.abs(o.cpu().numpy() - ref)[near].max()), rows_checked=int(near.sum()))
    kde.close()

    # ---- C3 ----
    from skysampler_b200 import _cabi as cabi
    kdes, qs, outs, zs = [], [], [], []
    for s in range(10):
        zz = whitened(4, 100000, 300 + s); zs.append(zz)
        kdes.append(GaussianKDE(0.1).fit(zz, sample_weight=synth.wide_weights(100000, 300 + s)))
        qs.append(torch.from_numpy(synth.queries(zz, 0.1, 1000000, 400 + s)).cuda())
        outs.append(torch.empty(1000000, dtype=torch.float64, device="cuda"))
    jobs = (cabi.SkbJob * 10)()
    for j in range(10):
        jobs[j].kde = kdes[j]._handle(); jobs[j].q = qs[j].data_ptr(); jobs[j].q_dtype = 0; jobs[j].q_is_whitened = 1
        jobs[j].M = 1000000; jobs[j].ldq = 4; jobs[j].cols = None; jobs[j].out_logp = outs[j].data_ptr()
    cabi.check(lib.skb_score_batch(jobs, 10, None))
    ms = timed(lambda: cabi.check(lib.skb_score_batch(jobs, 10, None)), 3)
    sub = np.random.RandomState(1).choice(1000000, 500, replace=False)
    ref = brute64.score(zs[3], synth.wide_weights(100000, 303), 0.1, qs[3].cpu().numpy()[sub], block=250)
    near = ref >= ref.max() - 100
    out["C3"] = dict(shells=10, d=4, n_train_per_shell=100000, n_query_per_shell=1000000, ms=ms, pairs_per_s=1e12 / ms * 1e3,
                     one_batched_launch=True, max_abs_dlogp_within_100_nats=float(np.abs(outs[3].cpu().numpy()[sub] - ref)[near].max()))
    for k in kdes:
        k.close()
    del qs, outs

    # ---- C5 sweep (8-D) ----
    sweep = []
    for n in (10000, 100000, 1000000, 4000000):
        z = whitened(8, n, 500 + int(np.log2(n)))
        kde = GaussianKDE(0.1).fit(z)
        qd = torch.from_numpy(synth.queries(z, 0.1, 1 << 20, 7)).cuda(); o = kde.score_samples(qd)
        lib.skb_debug_counters(0, st, 1)
        ms = timed(lambda: kde.score_samples(qd, out=o), 2)
        lib.skb_debug_counters(0, st, 1)
        err, nchk = parity(kde, z, None, 0.1, qd.cpu().numpy(), n=300 if n >= 1000000 else 1000)
        sweep.append(dict(n_train=n, n_query=1 << 20, ms=ms, fit_ms_device=kde.fit_ms_, pairs_per_s=n * float(1 << 20) / ms * 1e3,
                          tiles_culled_frac=st[2] / max(st[3], 1), max_abs_dlogp_within_100_nats=err, rows_checked=nchk))
        kde.close()
    out["C5_sweep_d8_1gpu"] = sweep

    # ---- C4 ----
    n = 2000000
    z = whitened(16, n, 401)
    kde = GaussianKDE(0.1).fit(z)
    chunk = 1 << 22
    nchunks = max(1, args.c4_queries // chunk)
    o = torch.empty(chunk, dtype=torch.float64, device="cuda")
    total_ms = 0.0
    t0 = time.time()
    first = None
    for c in range(nchunks):
        # queries: Philox draws from the KDE itself (whitened space), 10 % inflated x2
        qd, _, _ = kde.sample_philox(chunk, seed=402, offset=c * chunk, device_out=True)
        qd[: chunk // 10] *= 2.0
        if first is None:
            first = qd[:200000:200].cpu().numpy()
        total_ms += timed(lambda: kde.score_samples(qd, out=o), 1)
    err, nchk = parity(kde, z, None, 0.1, first, n=600)
    out["C4"] = dict(d=16, n_train=n, n_query=nchunks * chunk, h=0.1, kernel_s=total_ms / 1e3, wall_s=time.time() - t0,
                     pairs_per_s=n * float(nchunks * chunk) / total_ms * 1e3, path="direct FP32 per-query kernel",
                     max_abs_dlogp_within_100_nats=err, rows_checked=nchk)
    kde.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

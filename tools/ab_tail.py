"""Cost of bulk vs inflated-tail queries in the d = 8 score launch: python tools/ab_tail.py"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, _cabi, synth
import ctypes as C
lib = _cabi.lib()
d, n, m = 8, 1000000, 1 << 20
xx = synth.mixture_big(d, n, 201)
m1, pm, comps, std2 = synth.whiten_params(xx[::3])
x = np.ascontiguousarray((xx - m1) @ comps.T / std2)
kde = GaussianKDE(0.1).fit(x, sample_weight=synth.wide_weights(n, 3))
for tf in (0.0, 0.1, 0.3, 1.0):
    q = torch.from_numpy(synth.queries(x, 0.1, m, 5, tail_frac=tf)).cuda()
    out = kde.score_samples(q)
    _cabi.set_option("stats", 1)
    st = (C.c_uint64 * 4)(); lib.skb_debug_counters(0, st, 1)
    kde.score_samples(q, out=out)
    lib.skb_debug_counters(0, st, 1)
    _cabi.set_option("stats", 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        kde.score_samples(q, out=out)
    e1.record(); torch.cuda.synchronize()
    print("tail_frac %.1f: %.2f ms, exec_frac %.4f, finite %.3f, mean logp %.1f" % (
        tf, e0.elapsed_time(e1) / 3, st[1] * 32.0 / (float(n) * m), float(torch.isfinite(out).float().mean()), float(out.mean())), flush=True)
kde.close()

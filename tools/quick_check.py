"""Dev-time GPU sanity: parity vs the fp64 oracle on small cases + a first throughput number."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import torch
from skysampler_b200 import GaussianKDE, _cabi, synth
from oracle import brute64

lib = _cabi.lib()
print("devices", lib.skb_device_count())
for d in (1, 2, 3, 4, 5, 8, 11, 16):
    for weighted in (False, True):
        n, m, h = 5000, 3000, 0.1
        x = synth.mixture(d, n, 10 + d)
        mean1, pm, comps, std2 = synth.whiten_params(x)
        z = (x - mean1) @ comps.T / std2
        w = synth.wide_weights(n, 3) if weighted else None
        q = synth.queries(z, h, m, 99, tail_frac=0.2, tail_scale=2.5)
        kde = GaussianKDE(h).fit(z, sample_weight=w)
        t = time.time(); got = kde.score_samples(q); dt = time.time() - t
        ref = brute64.score(z, w, h, q)
        err = np.abs(got - ref)
        peak = ref.max()
        near = ref > peak - 100
        print("d=%2d w=%d  max|dlogp| all=%.2e  within100=%.2e (n=%d)  min ref=%.1f  nan=%d  %.3fs" % (
            d, weighted, err.max(), err[near].max(), near.sum(), ref.min(), np.isnan(got).sum(), dt))
        kde.close()

# throughput: d=8, N=1e6, M=2^20 resident on device
st = (C.c_uint64 * 4)(); lib.skb_debug_counters(0, st, 1)
for d, n, m in ((8, 1000000, 1 << 20), (4, 1000000, 1 << 20), (16, 1000000, 1 << 19), (1, 1000000, 1 << 20), (4, 10000, 100000)):
    x = np.random.RandomState(1).normal(size=(n, d))
    kde = GaussianKDE(0.1).fit(x)
    q = torch.randn(m, d, dtype=torch.float64, device="cuda")
    out = kde.score_samples(q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        out = kde.score_samples(q, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    st = (C.c_uint64 * 4)(); lib.skb_debug_counters(0, st, 1)
    print("d=%2d N=%d M=%d  %.3f ms  %.3e pairs/s  out[0]=%.4f  reruns %d / %d blocks" % (d, n, m, ms, n * m / ms * 1e3, out[0].item(), st[0], st[1]))
    kde.close()
fma = C.c_double(); ex2 = C.c_double()
print(lib.skb_measure_pipes(0, 200.0, C.byref(fma), C.byref(ex2)), "fma lane/s %.3e  ex2/s %.3e" % (fma.value, ex2.value))

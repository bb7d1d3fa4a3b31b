"""Throughput of the culled path per dimension on bench-like data (whitened mixture, KDE-draw queries):
python tools/scan_new.py dims N [kernel option]   (library from SKB_LIB_PATH)"""
import sys, os, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, _cabi, synth
lib = _cabi.lib()
dims = [int(x) for x in sys.argv[1].split(",")]
n = int(sys.argv[2])
_cabi.set_option("kernel", int(sys.argv[3]) if len(sys.argv) > 3 else 0)
fma, ex2 = C.c_double(), C.c_double()
lib.skb_measure_pipes(0, 50.0, C.byref(fma), C.byref(ex2))
for d in dims:
    m = (1 << 20) if d <= 8 else (1 << 19)
    xx = synth.mixture_big(d, n, 201)
    m1, pm, comps, std2 = synth.whiten_params(xx[::3])
    x = np.ascontiguousarray((xx - m1) @ comps.T / std2)
    q = torch.from_numpy(synth.queries(x, 0.1, m, 5)).cuda()
    kde = GaussianKDE(0.1).fit(x, sample_weight=synth.wide_weights(n, 3))
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
    ms = e0.elapsed_time(e1) / 3
    peak = min(fma.value / (2 * d + 1), ex2.value)
    ex = st[1] * 32.0
    print("%s d=%2d N=%d M=%d %8.2f ms  %.3e pairs/s  exec_frac=%.4f  frac_exec=%.3f  culled=%.3f" % (
        os.environ.get("SKB_LIB_PATH", "default")[-40:], d, n, m, ms, float(n) * m / ms * 1e3, ex / (float(n) * m),
        ex / (ms * 1e-3) / peak, st[2] / max(st[3], 1)), flush=True)
    kde.close()

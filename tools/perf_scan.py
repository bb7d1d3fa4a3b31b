"""Throughput of the score kernel per feature dimension for the library named by SKB_LIB_PATH."""
import sys, os, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, _cabi
lib = _cabi.lib()
_cabi.set_option("stats", 1)
fma, ex2 = C.c_double(), C.c_double()
lib.skb_measure_pipes(0, 100.0, C.byref(fma), C.byref(ex2))
dims = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 3, 4, 6, 8, 12, 16]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
for d in dims:
    m = (1 << 20) if d <= 8 else (1 << 19)
    mode = os.environ.get("SKB_SCAN_DATA", "normal")
    if mode == "normal":
        x = np.random.RandomState(1).normal(size=(n, d))
        q = torch.randn(m, d, dtype=torch.float64, device="cuda")
    else:                                   # whitened mixture + KDE-draw queries (10 % inflated), like the bench
        from skysampler_b200 import synth
        xx = synth.mixture(d, n, 201); m1, pm, comps, std2 = synth.whiten_params(xx)
        x = (xx - m1) @ comps.T / std2
        q = torch.from_numpy(synth.queries(x, 0.1, m, 5)).cuda()
    kde = GaussianKDE(0.1).fit(x)
    out = kde.score_samples(q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        kde.score_samples(q, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    rate = n * m / ms * 1e3
    peak = min(fma.value / (2 * d + 1), ex2.value)
    st = (C.c_uint64 * 4)(); lib.skb_debug_counters(0, st, 1)
    exec_pairs = st[1] * 32.0 / 4          # 4 launches since the last reset
    print("%s d=%2d N=%d M=%d %8.2f ms  %.3e pairs/s  frac_alg=%.3f  exec_frac=%.4f  frac_exec=%.3f  culled=%.3f  fit=%.1f ms" % (
        os.path.basename(_cabi.LIB_PATH), d, n, m, ms, rate, rate / peak, exec_pairs / (float(n) * m),
        exec_pairs / (ms * 1e-3) / peak, st[2] / max(st[3], 1), kde.fit_ms_), flush=True)
    kde.close()

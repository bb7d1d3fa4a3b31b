"""e2e (C ABI, pinned host rows) per step for several staged-chunk sizes: python tools/ab_e2e.py"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from skysampler_b200 import GaussianKDE, _cabi
from skysampler_b200.pipeline import RatioSampler
tr = bench.make_training(1000000)
kdes, lj = [], []
for name in ("cluster", "field"):
    t = tr[name]
    kdes.append(GaussianKDE(bench.H).fit(t["z"], sample_weight=t["w"], whitening=(t["mu"], t["W"], t["Winv"]), is_whitened=True))
    lj.append(float(np.log(abs(np.linalg.det(t["Winv"])))))
s = RatioSampler(kdes, signs=[1, -1], log_abs_jac=lj, log_m=0.0)
B = 1 << 20
q = s.draw(1, B, 402, offset=0)
qh = torch.empty((4, B, bench.D), dtype=torch.float64, pin_memory=True)
uh = torch.empty((4, B), dtype=torch.float64, pin_memory=True)
for b in range(4):
    qh[b].copy_(s.draw(1, B, 402, offset=(b + 1) * B)); uh[b].uniform_()
fl = torch.empty(B, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
for rows in (0, 1 << 19, 1 << 18, 3 << 17):
    _cabi.set_option("host_stage_rows", rows)
    for b in range(4): s.score_accept_host(qh[b], uh[b], fl)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(12):
        s.score_accept_host(qh[i % 4], uh[i % 4], fl); n = int(fl.sum())
    torch.cuda.synchronize()
    print("host_stage_rows %8d: %.2f ms per step (accepted %d)" % (rows, (time.perf_counter() - t0) / 12 * 1e3, n), flush=True)

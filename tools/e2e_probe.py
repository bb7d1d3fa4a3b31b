import sys, time, os
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from skysampler_b200 import GaussianKDE
from skysampler_b200.pipeline import RatioSampler
class A: pass
t0 = time.time(); tr = bench.make_training(1000000); print("make_training %.1fs" % (time.time() - t0))
kdes, lj = [], []
for name in ("cluster", "field"):
    t = tr[name]; t0 = time.time()
    kdes.append(GaussianKDE(0.1, device=0).fit(t["z"], sample_weight=t["w"], whitening=(t["mu"], t["W"], t["Winv"])))
    torch.cuda.synchronize(); print("fit %s N=1e6 d=8: %.3fs" % (name, time.time() - t0)); lj.append(t["log_abs_jac"])
for n, d in ((100000, 4), (100000, 1), (500000, 4)):
    x = np.random.RandomState(0).normal(size=(n, d)); t0 = time.time(); k = GaussianKDE(0.1).fit(x); torch.cuda.synchronize()
    print("fit N=%d d=%d: %.3fs" % (n, d, time.time() - t0)); k.close()
M = 1 << 20
pool, _, _ = kdes[1].sample_philox(M * 3, seed=402, device_out=True)
u = torch.rand(M * 3, dtype=torch.float64, device="cuda")
s = RatioSampler(kdes, [1, -1], lj, log_m=3.0)
qh = torch.empty((3, M, 8), dtype=torch.float64, pin_memory=True); uh = torch.empty((3, M), dtype=torch.float64, pin_memory=True)
for b in range(3): qh[b].copy_(pool[b*M:(b+1)*M]); uh[b].copy_(u[b*M:(b+1)*M])
fl = torch.empty(M, dtype=torch.uint8, pin_memory=True)
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    b = rep % 3
    e = [ev() for _ in range(6)]
    torch.cuda.synchronize()
    e[0].record(); qd = s._get("q_in", (M, 8), torch.float64); ud = s._get("u_in", (M,), torch.float64)
    qd.copy_(qh[b], non_blocking=True); ud.copy_(uh[b], non_blocking=True)
    e[1].record(); acc, _, _, _ = s.score_accept(qd, ud)
    e[2].record(); out, n = s.compact(acc, qd)
    e[3].record(); fl.copy_(acc, non_blocking=True)
    e[4].record(); torch.cuda.synchronize()
    print("rep %d: h2d %.2f ms  score %.2f  compact %.2f  d2h %.2f  (n=%d)" % (rep, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[3].elapsed_time(e[4]), n))
    e[0].record(); acc, _, _, _ = s.score_accept(pool[b*M:(b+1)*M], u[b*M:(b+1)*M]); e[1].record(); torch.cuda.synchronize()
    print("        resident score %.2f ms" % e[0].elapsed_time(e[1]))

"""The BASELINE target shape on ONE B200: 8-D, 1e6 weighted training rows x 1e8 queries (north_star: >= 60 % of the pair
roofline = <= ~76 s).  Queries are Philox draws from the KDE itself on the device, 10 % inflated x2, generated batch by
batch (2^20 rows); parity is checked on a fixed subset against the fp64 brute-force oracle.

    python tools/run_target.py [n_query] > profiles/rXX_target_8d_1e6x1e8.json
"""
import ctypes as C, json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, _cabi, synth
from oracle import cbrute

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
n, d, h, batch = 1000000, 8, 0.1, 1 << 20
x = synth.mixture_big(d, n, 201)
w = synth.wide_weights(n, 202)
m1, pm, comps, std2 = synth.whiten_params(x[::3], w[::3])
z = np.ascontiguousarray((x - m1) @ comps.T / std2)
kde = GaussianKDE(h).fit(z, sample_weight=w)
lib = _cabi.lib()
fma, ex2 = C.c_double(), C.c_double()
lib.skb_measure_pipes(0, 200.0, C.byref(fma), C.byref(ex2))
peak = min(fma.value / (2 * d + 1), ex2.value)
nb = (M + batch - 1) // batch
q = [torch.empty((batch, d), dtype=torch.float64, device="cuda") for _ in range(2)]
out = torch.empty(batch, dtype=torch.float64, device="cuda")
check_q, check_lp = [], []
def gen(b, buf):
    kde_q, _, _ = kde.sample_philox(batch, seed=402, offset=b * batch, device_out=True)     # whitened-space draws
    buf.copy_(kde_q); buf[: batch // 10] *= 2.0
    return buf
gen(0, q[0]); kde.score_samples(q[0], out=out); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
lsum = torch.zeros((), dtype=torch.float64, device="cuda")
t0 = time.time(); e0.record()
for b in range(nb):
    buf = gen(b, q[b & 1])
    kde.score_samples(buf, out=out)
    lsum += out.sum()
    if b % 16 == 0:
        check_q.append(buf[::4096].clone()); check_lp.append(out[::4096].clone())
e1.record(); torch.cuda.synchronize(); wall = time.time() - t0
ms = e0.elapsed_time(e1)
cq = torch.cat(check_q).cpu().numpy(); clp = torch.cat(check_lp).cpu().numpy()
ref = cbrute.score(z, w, h, cq)
near = ref >= ref.max() - 100
res = dict(shape="8-D, %d weighted training rows x %d queries (draws incl. generation on device), 1 x B200" % (n, nb * batch),
           seconds=ms / 1e3, wall_seconds=wall, pairs_per_s=float(n) * nb * batch / ms * 1e3,
           pair_roofline_per_s=peak, frac_of_pair_roofline_algorithmic=float(n) * nb * batch / ms * 1e3 / peak,
           target_seconds_at_60pct=float(n) * nb * batch / (0.6 * peak),
           parity_rows=int(near.sum()), max_abs_dlogp_within_100_nats=float(np.abs(clp - ref)[near].max()),
           mean_logp=float(lsum.item() / (nb * batch)), fit_ms_device=kde.fit_ms_)
print(json.dumps(res, indent=1))

"""Tiny end-to-end exercise of every kernel for compute-sanitizer (one tool per gpurun call)."""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np, pandas as pd, torch
from skysampler_b200 import GaussianKDE, KDEContainer, emulator, synth, _cabi
rs = np.random.RandomState(0)
for d in (1, 3, 4, 8, 16):
    z = rs.normal(size=(700, d)); w = rs.uniform(0, 2, 700); w[:30] = 0
    kde = GaussianKDE(0.2).fit(z, sample_weight=w)
    q = 1.5 * rs.normal(size=(900, d))
    a = kde.score_samples(q); m, s = kde.score_partial(q[:100])
    kde.sample(300, random_state=1); kde.sample_philox(500, seed=3)
    kde.close()
    print("d", d, "ok", float(a[0]))
z = rs.normal(size=(40000, 4))
kde = GaussianKDE(0.1).fit(z)                      # > 128 tiles: split + last-arriver merge path
print("split path", float(kde.score_samples(rs.normal(size=(300, 4)))[0])); kde.close()
df, w = synth.frame(4, 2000, 5)
c = KDEContainer(df, w, seed=1); c.standardize_data(); c.construct_kde(0.1)
prop = c.random_draw(1500, rmin=float(df["LOGR"].quantile(0.2)), rmax=float(df["LOGR"].quantile(0.8)))
out = emulator.score_group([c], [list(df.columns)], prop, sign=[1], log_m=-3.0, u=rs.uniform(size=len(prop)))
fl = torch.from_numpy(out["accept"].astype(np.uint8)).cuda(); rows = torch.from_numpy(prop.values).cuda()
packed = torch.empty_like(rows); cnt = C.c_int64()
_cabi.check(_cabi.lib().skb_compact(_cabi.ptr(fl), len(prop), _cabi.ptr(rows), 0, 4, _cabi.ptr(packed), None, C.byref(cnt), None))
print("accepted", cnt.value, int(out["accept"].sum())); c.drop_kde()
torch.cuda.synchronize(); print("done")

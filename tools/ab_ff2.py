"""2-D far-field cell records on / off / auto: python tools/ab_ff2.py"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, _cabi
m = 1 << 20
for n in (50000, 100000, 200000, 300000, 500000, 1000000):
    x = np.random.RandomState(1).normal(size=(n, 2))
    q = torch.randn(m, 2, dtype=torch.float64, device="cuda")
    kde = GaussianKDE(0.1).fit(x)
    res = []
    for mode in (0, 1, -1):
        _cabi.set_option("ff2_cells", mode)
        out = kde.score_samples(q)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3): kde.score_samples(q, out=out)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 3)
    print("N=%7d  cells off %.2f ms  on %.2f ms  auto %.2f ms" % (n, res[0], res[1], res[2]), flush=True)
    kde.close()

import sys
sys.path.insert(0, ".")
import numpy as np
from skysampler_b200 import GaussianKDE, _cabi
from oracle import brute64
for n in (16391, 16448, 32775, 20000):
    rs = np.random.RandomState(n)
    z = rs.normal(size=(n, 8)); q = 1.2 * rs.normal(size=(777, 8))
    w = rs.uniform(0.5, 2.0, n) if n > 1 else None
    ref = brute64.score(z, w, 0.4, q)
    for rep in range(4):
        kde = GaussianKDE(0.4).fit(z, sample_weight=w)
        out = []
        for name, opts in (("pq", dict(kernel=1)), ("sparse", dict(kernel=2)), ("dense", dict(cull=0))):
            for k, v in opts.items():
                _cabi.set_option(k, v)
            got = kde.score_samples(q)
            _cabi.set_option("cull", 1); _cabi.set_option("kernel", 0)
            err = np.abs(got - ref)
            out.append("%s %.2e (%d bad)" % (name, err.max(), int((err > 1e-4).sum())))
        print("n=%d rep=%d  " % (n, rep) + "   ".join(out), flush=True)
        kde.close()

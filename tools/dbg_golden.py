import sys; sys.path.insert(0, ".")
import numpy as np, pandas as pd
from skysampler_b200 import KDEContainer, GaussianKDE
from oracle import brute64, draw
lg = lambda n: dict(np.load("tests/golden/%s.npz" % n))
g0 = lg("c1_train"); data, weights = g0["data"], g0["weights"]; cols = [str(c) for c in g0["columns"]]
for case in ["c1_score_w_h0p05", "c1_score_u_h0p1"]:
    g = lg(case); h = float(g["h"]); wt = bool(g["weighted"])
    w = weights if wt else None
    z = draw.whiten(data, g["mean1"], g["pca_mean"], g["components"], g["std2"])
    zq = draw.whiten(g["query"], g["mean1"], g["pca_mean"], g["components"], g["std2"])
    ref = brute64.score(z, w, h, zq); near = ref >= ref.max() - 100
    c = KDEContainer(pd.DataFrame(data, columns=cols), pd.Series(weights) if wt else None)
    c.set_whitening(g["mean1"], g["pca_mean"], g["components"], g["std2"])
    c.construct_kde(h)
    lp, _ = c.score_samples(pd.DataFrame(g["query"], columns=cols))
    k2 = GaussianKDE(h).fit(z, sample_weight=w if wt else np.ones(len(z)))
    lp2 = k2.score_samples(zq)
    k3 = GaussianKDE(h).fit(z, sample_weight=w)
    lp3 = k3.score_samples(zq)
    for name, v in (("container raw", lp), ("whitened w", lp2), ("whitened", lp3)):
        e = np.abs(v - ref); i = np.argmax(np.where(near, e, 0))
        print(case, name, "max within100 %.2e at ref=%.2f (i=%d)  bulk(ref>-20) %.2e" % (e[near].max(), ref[i], i, e[ref > -20].max()))
    bad = np.argsort(-np.where(near, np.abs(lp - ref), 0))[:8]
    print("  worst:", [(int(i), round(float(ref[i]), 2), float(lp[i] - ref[i]), float(lp2[i]-ref[i])) for i in bad])
    print("  data diff", np.abs(np.asarray(c._data) - z).max())

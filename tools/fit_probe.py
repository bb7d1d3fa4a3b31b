"""Fit cost: device time (skb_fit_ms) and wall time of skb_create for a few (d, N), host and device inputs."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, synth

for d, n in ((1, 1000000), (4, 100000), (4, 1000000), (8, 1000000), (8, 4000000), (16, 2000000)):
    x = synth.mixture_big(d, n, 7)
    w = synth.wide_weights(n, 8)
    xd = torch.from_numpy(x).cuda()
    for label, src in (("host", x), ("device", xd)):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            kde = GaussianKDE(0.1).fit(src, sample_weight=w)
            wall = (time.perf_counter() - t0) * 1e3
            fit = kde.fit_ms_
            kde.close()
        print("fit d=%2d N=%8d input=%-6s  device %.2f ms  wall %.1f ms" % (d, n, label, fit, wall), flush=True)

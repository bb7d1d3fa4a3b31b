"""A/B of score-kernel variants on bench-like data: python tools/ab_pq.py dims N tag   (library from SKB_LIB_PATH).
Prints ms per launch and compares the log-densities with the first variant's (saved under gpurun_out/ab_ref_d*.npy):
the variants only move data differently, so the bits must agree."""
import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, _cabi, synth
dims = [int(x) for x in sys.argv[1].split(",")]
n = int(sys.argv[2])
tag = sys.argv[3] if len(sys.argv) > 3 else "default"
os.makedirs("gpurun_out", exist_ok=True)
for d in dims:
    m = (1 << 20) if d <= 8 else (1 << 18)
    xx = synth.mixture_big(d, n, 201)
    m1, pm, comps, std2 = synth.whiten_params(xx[::3])
    x = np.ascontiguousarray((xx - m1) @ comps.T / std2)
    q = torch.from_numpy(synth.queries(x, 0.1, m, 5)).cuda()
    kde = GaussianKDE(0.1).fit(x, sample_weight=synth.wide_weights(n, 3))
    out = kde.score_samples(q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(4):
        kde.score_samples(q, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    got = out.cpu().numpy()
    ref_path = "gpurun_out/ab_ref_d%d_n%d.npy" % (d, n)
    if os.path.exists(ref_path):
        ref = np.load(ref_path)
        same = bool(np.array_equal(got, ref, equal_nan=True))
        diff = float(np.nanmax(np.abs(got - ref)))
    else:
        np.save(ref_path, got)
        same, diff = True, 0.0
    print("%-10s d=%2d N=%d M=%d %8.2f ms  %.3e pairs/s  bits_equal=%s maxdiff=%.2e" % (
        tag, d, n, m, ms, float(n) * m / ms * 1e3, same, diff), flush=True)
    kde.close()

"""One culled score launch on bench-like data (whitened mixture, KDE-draw queries), for ncu:
python tools/prof_mix.py D N M [kernel option] [reps]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, _cabi, synth
d, n, m = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
_cabi.set_option("kernel", int(sys.argv[4]) if len(sys.argv) > 4 else 0)
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
xx = synth.mixture_big(d, n, 201)
m1, pm, comps, std2 = synth.whiten_params(xx[::3])
x = np.ascontiguousarray((xx - m1) @ comps.T / std2)
q = torch.from_numpy(synth.queries(x, 0.1, m, 5)).cuda()
kde = GaussianKDE(0.1).fit(x, sample_weight=synth.wide_weights(n, 3))
out = torch.empty(m, dtype=torch.float64, device="cuda")
for _ in range(reps):
    kde.score_samples(q, out=out)
torch.cuda.synchronize()
print("ok", out[:2].tolist())

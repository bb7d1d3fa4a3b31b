"""One score launch of a given shape, for ncu: python tools/prof_case.py D N M [reps]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE
d, n, m = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
x = np.random.RandomState(1).normal(size=(n, d))
kde = GaussianKDE(0.1).fit(x)
q = torch.randn(m, d, dtype=torch.float64, device="cuda")
out = torch.empty(m, dtype=torch.float64, device="cuda")
for _ in range(reps):
    kde.score_samples(q, out=out)
torch.cuda.synchronize()
print("ok", out[:2].tolist())

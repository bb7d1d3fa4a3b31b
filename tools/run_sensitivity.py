"""How the headline rate depends on the synthetic recipe and on the bandwidth.

bench.py scores candidates drawn from the field KDE against a cluster and a field KDE that are two samples of the SAME
8-D mixture (field covariance x1.2); SURVEY §8d's recipe uses two INDEPENDENT mixtures (field x1.5).  This tool runs the
fused ratio launch (2 KDEs x 1e6 rows, 2^20 candidates) on both recipes and on h in {0.05, 0.1, 0.2, 0.3}, and prints
pairs/s, the executed-pair fraction and the executed-work roofline fraction for each.

    python tools/run_sensitivity.py > profiles/rXX_bench_sensitivity.json
"""
import ctypes as C, json, sys
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, _cabi, synth
from skysampler_b200.pipeline import RatioSampler

D, N, M = 8, 1000000, 1 << 20
lib = _cabi.lib()
fma, ex2 = C.c_double(), C.c_double()
lib.skb_measure_pipes(0, 100.0, C.byref(fma), C.byref(ex2))
peak = min(fma.value / (2 * D + 1), ex2.value)


def training(recipe):
    out = []
    specs = ((201, 201, 1.0), (201, 202, 1.2)) if recipe == "bench" else ((201, None, 1.0), (202, None, 1.5))
    for mix_seed, sample_seed, cov in specs:
        x = synth.mixture(D, N, mix_seed, cov_scale=cov, sample_seed=sample_seed)
        w = synth.wide_weights(N, mix_seed + 7919 + (sample_seed or 0))
        mean1, pmean, comps, std2 = synth.whiten_params(x, w)
        mu = mean1 + pmean
        out.append(dict(z=(x - mu) @ (comps.T / std2[None, :]), w=w, mu=mu, W=comps.T / std2[None, :], Winv=std2[:, None] * comps))
    return out


res = []
for recipe in ("bench", "survey"):
    tr = training(recipe)
    for h in (0.05, 0.1, 0.2, 0.3):
        if recipe == "survey" and h not in (0.1,):
            continue
        kdes = [GaussianKDE(h).fit(t["z"], sample_weight=t["w"], whitening=(t["mu"], t["W"], t["Winv"])) for t in tr]
        s = RatioSampler(kdes, signs=[1, -1], log_abs_jac=[0.0, 0.0], log_m=0.0)
        q = s.draw(1, M, seed=11)
        mu_f = torch.from_numpy(tr[1]["mu"]).cuda()
        q[:M // 10].mul_(2.0).sub_(mu_f)
        u = torch.rand(M, dtype=torch.float64, device="cuda")
        st = (C.c_uint64 * 4)()
        with _cabi.option("stats", 1):
            lib.skb_debug_counters(0, st, 1)
            s.score_accept(q, u)
            lib.skb_debug_counters(0, st, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            s.score_accept(q, u)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        ex = st[1] * 32.0
        res.append(dict(recipe=recipe, h=h, ms_per_launch=ms, pairs_per_s=2.0 * N * M / ms * 1e3,
                        executed_pair_fraction=ex / (2.0 * N * M), frac_executed=ex / (ms * 1e-3) / peak,
                        tiles_culled_frac=st[2] / max(st[3], 1)))
        for k in kdes:
            k.close()
print(json.dumps({"what": "fused 2-KDE ratio launch, d = 8, 1e6 training rows per KDE, 2^20 candidates (10 % inflated x2)",
                  "pair_roofline_per_s": peak, "runs": res}, indent=1))

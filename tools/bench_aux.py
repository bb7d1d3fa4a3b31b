"""HBM-bound helpers of the path: achieved GB/s of the draw and compaction kernels vs the measured copy peak."""
import sys, json, os, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, _cabi, synth
lib = _cabi.lib()
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]; src = "MEASURED_PEAKS.json"
except Exception:
    peak = 6650.0; src = "fallback"
d, n, m = 8, 1000000, 1 << 24
x = synth.mixture(d, n, 201); w = synth.wide_weights(n, 1)
m1, pm, comps, std2 = synth.whiten_params(x, w)
mu = m1 + pm; W = comps.T / std2[None, :]; Winv = std2[:, None] * comps
kde = GaussianKDE(0.1).fit((x - mu) @ W, sample_weight=w, whitening=(mu, W, Winv))
dev = torch.device("cuda", 0)
out = torch.empty((m, d), dtype=torch.float64, device=dev)
fl = torch.empty(m, dtype=torch.uint8, device=dev)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
st = _cabi.stream_ptr(torch.cuda.current_stream())
def draw(rcol=-1, lo=0.0, hi=0.0):
    _cabi.check(lib.skb_draw(kde._handle(), m, None, None, 7, 0, None, rcol, lo, hi, 1000, None, _cabi.ptr(out), d, None,
                             _cabi.ptr(fl), None, st))
res = {}
ms = timed(draw)
byts = m * (8 * d + 8 * d + 1)            # centre gather + row write + flag
res["draw_philox"] = dict(ms=ms, draws_per_s=m / ms * 1e3, algorithmic_GBs=byts / ms / 1e6, frac_of_hbm=byts / ms / 1e6 / peak)
lo, hi = np.percentile(x[:, d - 1], [20, 70])
ms = timed(lambda: draw(d - 1, float(lo), float(hi)))
res["draw_philox_window"] = dict(ms=ms, draws_per_s=m / ms * 1e3, note="in-kernel redraw until inside [p20, p70] of the last column")
packed = torch.empty_like(out)
for frac in (0.02, 0.5):
    flags = (torch.rand(m, device=dev) < frac).to(torch.uint8)
    cnt = C.c_int64()
    def comp():
        _cabi.check(lib.skb_compact(_cabi.ptr(flags), m, _cabi.ptr(out), 0, d, _cabi.ptr(packed), None, None, st))
    ms = timed(comp)
    k = int(flags.sum().item())
    byts = 2 * m + 2 * k * 8 * d
    res["compact_f%.2f" % frac] = dict(ms=ms, rows_per_s=m / ms * 1e3, algorithmic_GBs=byts / ms / 1e6, frac_of_hbm=byts / ms / 1e6 / peak)
res["hbm_peak_gbs"] = peak; res["peak_source"] = src; res["rows"] = m; res["d"] = d
print(json.dumps(res, indent=1))
kde.close()

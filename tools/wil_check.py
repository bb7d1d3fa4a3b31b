"""Quick A/B of the work-item-per-lane kernel against the round-1 culled kernels and the fp64 oracle."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import GaussianKDE, _cabi, synth
from oracle import cbrute

def case(d, n, m, h=0.1, weighted=True, seed=1):
    x = synth.mixture_big(d, n, seed)
    m1, pm, comps, std2 = synth.whiten_params(x[::3])
    z = np.ascontiguousarray((x - m1) @ comps.T / std2)
    w = synth.wide_weights(n, seed + 1) if weighted else None
    q = synth.queries(z, h, m, seed + 2)
    kde = GaussianKDE(h).fit(z, sample_weight=w)
    qd = torch.from_numpy(q).cuda()
    res = {}
    for kern in (1, 2):
        _cabi.set_option("kernel", kern)
        out = kde.score_samples(qd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            kde.score_samples(qd, out=out)
        e1.record(); torch.cuda.synchronize()
        res[kern] = (out.cpu().numpy(), e0.elapsed_time(e1) / 3)
    _cabi.set_option("kernel", 2)
    sub = np.random.RandomState(0).choice(m, min(m, 1500), replace=False)
    ref = cbrute.score(z, w, h, q[sub])
    near = ref >= ref.max() - 100
    e_new = np.abs(res[2][0][sub] - ref)[near].max()
    e_old = np.abs(res[1][0][sub] - ref)[near].max()
    # batching invariance of the new kernel
    a = kde.score_samples(qd[1000:3000].contiguous()).cpu().numpy()
    same = np.array_equal(a, res[2][0][1000:3000])
    print("d=%2d N=%8d M=%8d  old %8.2f ms  new %8.2f ms  speedup %.2fx  err old %.2e new %.2e  new-vs-old max %.2e  batch-invariant %s"
          % (d, n, m, res[1][1], res[2][1], res[1][1] / res[2][1], e_old, e_new,
             np.abs(res[2][0] - res[1][0])[np.isfinite(res[1][0])].max(), same), flush=True)
    _cabi.set_option("kernel", 0)
    kde.close()

if __name__ == "__main__":
    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    case(8, 20000, 5000)
    case(4, 20000, 5000)
    case(2, 5000, 3000, weighted=False)
    case(8, 100000, 50000)
    if not small:
        for d in (1, 2, 3, 4, 6, 8, 12, 16):
            case(d, 1000000, (1 << 20) if d <= 8 else (1 << 19))

"""Offline: fraction of (query warp, tile) pairs that survive the cull for alternative tilings / bounds."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from skysampler_b200 import synth
torch.set_num_threads(16)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
TN = int(sys.argv[3]) if len(sys.argv) > 3 else 64
h = 0.1
c2 = np.log2(np.e) / (2 * h * h)
xx = synth.mixture(d, N, 201); m1, pm, comps, std2 = synth.whiten_params(xx)
z = ((xx - m1) @ comps.T / std2).astype(np.float32)
M = 1 << 20
q = synth.queries(z.astype(np.float64), h, M, 5).astype(np.float32)
# morton order
lo, hi = z.min(0), z.max(0)
bits = min(63 // d, 20)
g = np.clip(((q - lo) / (hi - lo) * (1 << bits)), 0, (1 << bits) - 1).astype(np.uint64)
key = np.zeros(M, np.uint64)
for b in range(bits - 1, -1, -1):
    for k in range(d):
        key = (key << np.uint64(1)) | ((g[:, k] >> np.uint64(b)) & np.uint64(1))
order = np.argsort(key, kind="stable")
qs = q[order]
rs = np.random.RandomState(0)
NW = 150
warps = rs.choice(M // 32, NW, replace=False)
Q = np.stack([qs[w * 32:(w + 1) * 32] for w in warps])       # NW,32,d
Qt = torch.from_numpy(Q.reshape(-1, d))
zt = torch.from_numpy(z)
# exact nn distance^2
t0 = time.time()
nn = torch.full((Qt.shape[0],), float("inf"))
for s in range(0, N, 100000):
    dd = torch.cdist(Qt, zt[s:s + 100000]) ** 2
    nn = torch.minimum(nn, dd.min(1).values)
print("nn done %.1fs  median nn dist %.3f" % (time.time() - t0, float(nn.sqrt().median())), flush=True)
thr = (torch.floor(nn * c2) - 90 + 128) / c2                   # cull when lb > thr

def split_tree(idx, choose):
    out = []
    stack = [idx]
    while stack:
        ix = stack.pop()
        n = len(ix)
        if n <= TN:
            out.append(ix); continue
        nun = (n + TN - 1) // TN
        nl = ((nun + 1) // 2) * TN
        proj = choose(z[ix])
        part = np.argpartition(proj, nl)
        stack.append(ix[part[nl:]]); stack.append(ix[part[:nl]])
    return out

def widest(p): return p[:, np.argmax(p.max(0) - p.min(0))]
def maxvar(p): return p[:, np.argmax(p.var(0))]
def pcadir(p):
    pc = p - p.mean(0)
    if len(p) > 20000: sub = pc[:: len(p) // 20000]
    else: sub = pc
    w, v = np.linalg.eigh(sub.T @ sub)
    return pc @ v[:, -1]
def twomeans(p):
    pc = p - p.mean(0)
    sub = pc if len(p) <= 20000 else pc[:: len(p) // 20000]
    w, v = np.linalg.eigh(sub.T @ sub)
    dirn = v[:, -1]
    for _ in range(3):
        s = sub @ dirn > 0
        if s.all() or (~s).all(): break
        dirn = sub[s].mean(0) - sub[~s].mean(0); dirn /= np.linalg.norm(dirn) + 1e-30
    return pc @ dirn

def evaluate(name, tiles):
    T = len(tiles)
    blo = np.stack([z[t].min(0) for t in tiles]); bhi = np.stack([z[t].max(0) for t in tiles])
    cen = np.stack([z[t].mean(0) for t in tiles])
    rad = np.array([np.sqrt(((z[t] - c) ** 2).sum(1).max()) for t, c in zip(tiles, cen)])
    blo, bhi, cen, rad = map(torch.from_numpy, (blo, bhi, cen, rad.astype(np.float32)))
    need_box = 0; need_ball = 0; need_both = 0; true_need = 0
    for w in range(NW):
        qq = Qt[w * 32:(w + 1) * 32]                           # 32,d
        gap = torch.clamp(torch.maximum(blo[None] - qq[:, None], qq[:, None] - bhi[None]), min=0)
        lb_box = (gap ** 2).sum(-1)                              # 32,T
        lb_ball = torch.clamp(torch.cdist(qq, cen) - rad[None], min=0) ** 2
        th = thr[w * 32:(w + 1) * 32, None]
        need_box += int(((lb_box <= th).any(0)).sum())
        need_ball += int(((lb_ball <= th).any(0)).sum())
        need_both += int(((torch.maximum(lb_box, lb_ball) <= th).any(0)).sum())
    print("%-10s tiles %d  mean radius %.3f  box-diag %.3f | needed frac: box %.4f  ball %.4f  both %.4f" % (
        name, T, float(rad.mean()), float(((bhi - blo) ** 2).sum(1).sqrt().mean()), need_box / (NW * T), need_ball / (NW * T), need_both / (NW * T)), flush=True)

for name, fn in (("kd-widest", widest), ("kd-maxvar", maxvar), ("pca-tree", pcadir), ("2means", twomeans)):
    t0 = time.time()
    tiles = split_tree(np.arange(N), fn)
    print("build %.1fs" % (time.time() - t0))
    evaluate(name, tiles)

# ---- query grouping by home tile (kd leaf order) instead of Morton ----
print("---- query order experiments (kd-maxvar / kd-widest tiles, box bound) ----")
for name, fn in (("kd-widest", widest), ("kd-maxvar", maxvar)):
    tiles = split_tree(np.arange(N), fn)
    T = len(tiles)
    blo = torch.from_numpy(np.stack([z[t].min(0) for t in tiles])); bhi = torch.from_numpy(np.stack([z[t].max(0) for t in tiles]))
    # home tile of a 2^17 subsample of queries would change density; do all 2^20 in chunks
    qt_all = torch.from_numpy(q)
    home = torch.empty(M, dtype=torch.int64)
    for s in range(0, M, 4096):
        qq = qt_all[s:s + 4096]
        gap = torch.clamp(torch.maximum(blo[None] - qq[:, None], qq[:, None] - bhi[None]), min=0)
        home[s:s + 4096] = (gap ** 2).sum(-1).argmin(1)
    key2 = home.numpy().astype(np.uint64) * np.uint64(1 << 40) + (key >> np.uint64(24))
    order2 = np.argsort(key2, kind="stable")
    qs2 = q[order2]
    Q2 = torch.from_numpy(np.stack([qs2[w * 32:(w + 1) * 32] for w in warps]).reshape(-1, d))
    nn2 = torch.full((Q2.shape[0],), float("inf"))
    for s in range(0, N, 100000):
        nn2 = torch.minimum(nn2, (torch.cdist(Q2, zt[s:s + 100000]) ** 2).min(1).values)
    thr2 = (torch.floor(nn2 * c2) - 90 + 128) / c2
    need = 0; pairneed = 0
    for w in range(NW):
        qq = Q2[w * 32:(w + 1) * 32]
        gap = torch.clamp(torch.maximum(blo[None] - qq[:, None], qq[:, None] - bhi[None]), min=0)
        lb = (gap ** 2).sum(-1)
        ok = lb <= thr2[w * 32:(w + 1) * 32, None]
        need += int(ok.any(0).sum()); pairneed += int(ok.sum())
    print("%-10s home-tile order: needed frac %.4f   (per-query needed frac %.4f)" % (name, need / (NW * T), pairneed / (NW * T * 32)), flush=True)

"""Where does the end-to-end step spend its time?  Per-batch device timings of RatioSampler.run_host's
pipeline (H2D on a side stream, score + accept, compaction, D2H) next to a resident-input step."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from skysampler_b200 import GaussianKDE
from skysampler_b200.pipeline import RatioSampler

tr = bench.make_training(1000000)
kdes, lj = [], []
for name in ("cluster", "field"):
    t = tr[name]
    kdes.append(GaussianKDE(0.1, device=0).fit(t["z"], sample_weight=t["w"], whitening=(t["mu"], t["W"], t["Winv"])))
    lj.append(t["log_abs_jac"])
M, NB = 1 << 20, 6
pool, _, _ = kdes[1].sample_philox(M * NB, seed=402, device_out=True)
u = torch.rand(M * NB, dtype=torch.float64, device="cuda")
s = RatioSampler(kdes, [1, -1], lj, log_m=3.0)
qh = torch.empty((NB, M, 8), dtype=torch.float64, pin_memory=True); uh = torch.empty((NB, M), dtype=torch.float64, pin_memory=True)
for b in range(NB):
    qh[b].copy_(pool[b * M:(b + 1) * M]); uh[b].copy_(u[b * M:(b + 1) * M])
torch.cuda.synchronize()


def ev():
    return torch.cuda.Event(enable_timing=True)


for b in range(NB):                                   # resident steps
    e0, e1, e2 = ev(), ev(), ev()
    e0.record(); acc, _, _, _ = s.score_accept(pool[b * M:(b + 1) * M], u[b * M:(b + 1) * M]); e1.record()
    rows, n = s.compact(acc, pool[b * M:(b + 1) * M]); e2.record(); torch.cuda.synchronize()
    print("resident batch %d: score %.2f ms  compact %.2f ms  n=%d" % (b, e0.elapsed_time(e1), e1.elapsed_time(e2), n))

batches = [(qh[b], uh[b]) for b in range(NB)]
s.run_host(batches)                                    # warm-up (slots, pinned outputs)
main = torch.cuda.current_stream(); side = s._side
slots = [(s._get("q_pipe%d" % k, (M, 8), torch.float64), s._get("u_pipe%d" % k, (M,), torch.float64), s._get("acc_pipe%d" % k, (M,), torch.uint8)) for k in range(2)]
ready = [torch.cuda.Event(), torch.cuda.Event()]; free = [torch.cuda.Event(), torch.cuda.Event()]
marks = []
fh = torch.empty((NB, M), dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
t0 = ev(); t1 = ev(); t0.record()


def upload(i):
    qd, ud, _ = slots[i % 2]
    with torch.cuda.stream(side):
        if i >= 2:
            side.wait_event(free[i % 2])
        a, b = ev(), ev(); a.record(side)
        qd.copy_(batches[i][0], non_blocking=True); ud.copy_(batches[i][1], non_blocking=True)
        b.record(side); ready[i % 2].record(side)
    return a, b


up = [None] * NB
up[0] = upload(0)
for i in range(NB):
    if i + 1 < NB:
        up[i + 1] = upload(i + 1)
    qd, ud, keep = slots[i % 2]
    main.wait_event(ready[i % 2])
    a, b, c = ev(), ev(), ev()
    a.record(); acc, _, _, _ = s.score_accept(qd, ud, stream=main); b.record()
    keep.copy_(acc); s.compact(keep, qd, stream=main, sync_count=False); free[i % 2].record(main)
    fh[i].copy_(keep, non_blocking=True); c.record()
    marks.append((a, b, c))
t1.record(); torch.cuda.synchronize()
for i, (a, b, c) in enumerate(marks):
    print("pipelined batch %d: h2d %.2f ms  score %.2f ms  rest %.2f ms" % (i, up[i][0].elapsed_time(up[i][1]), a.elapsed_time(b), b.elapsed_time(c)))
print("pipelined total %.2f ms for %d batches = %.2f ms/batch" % (t0.elapsed_time(t1), NB, t0.elapsed_time(t1) / NB))

# ---- the library call itself, as bench.py times it (10 batches), with and without nvidia-smi polling ----
import subprocess
NB2 = 10
qh2 = torch.empty((NB2, M, 8), dtype=torch.float64, pin_memory=True); uh2 = torch.empty((NB2, M), dtype=torch.float64, pin_memory=True)
for b in range(NB2):
    qh2[b].copy_(pool[(b % NB) * M:((b % NB) + 1) * M]); uh2[b].copy_(u[(b % NB) * M:((b % NB) + 1) * M])
b10 = [(qh2[b], uh2[b]) for b in range(NB2)]
s.run_host(b10); torch.cuda.synchronize()
for poll in (False, True, False):
    proc = None
    if poll:
        proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "200"],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(0.3)
    a, b = ev(), ev(); torch.cuda.synchronize(); w0 = time.time(); a.record()
    s.run_host(b10)
    b.record(); torch.cuda.synchronize(); w1 = time.time()
    print("run_host x10 (nvidia-smi polling %s): %.2f ms/batch by events, %.2f ms/batch wall" % (poll, a.elapsed_time(b) / NB2, (w1 - w0) * 1e3 / NB2))
    if proc:
        proc.terminate()

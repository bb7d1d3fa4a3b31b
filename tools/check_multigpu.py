"""2+ GPU validation of skysampler_b200.dist on real NCCL (run under torchrun):
query-sharded scoring is bit-identical to one GPU; training-sharded partial-LSE combine matches the
full-set score and the fp64 oracle.  Prints one summary line per check on rank 0."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from skysampler_b200 import GaussianKDE, synth
from skysampler_b200 import dist as skd
from oracle import brute64

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.pop("NCCL_DEBUG", None)
dist.init_process_group("nccl", device_id=dev)
d, n, m, h = 8, 300000, 200000, 0.1
if rank == 0:
    x = synth.mixture(d, n, 501); w = synth.wide_weights(n, 502)
    m1, pm, comps, std2 = synth.whiten_params(x, w)
    z = (x - m1) @ comps.T / std2
    q = synth.queries(z, h, m, 503)
else:
    z = w = q = None
z = skd.broadcast_array(z, 0, dev); w = skd.broadcast_array(w, 0, dev); q = skd.broadcast_array(q, 0, dev)
qd = torch.from_numpy(q).to(dev)

single = GaussianKDE(h, device=local).fit(z, sample_weight=w)
lp1 = single.score_samples(qd)
single.close()

qs = skd.ShardedKDE(h, mode="query", device=local).fit(z, sample_weight=w)
lpq = qs.score_samples(qd)
qs.close()
ts = skd.ShardedKDE(h, mode="train", device=local).fit(z, sample_weight=w)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
torch.cuda.synchronize(); dist.barrier(); e0.record()
lpt = ts.score_samples(qd)
e1.record(); torch.cuda.synchronize()
ms = skd.timed_max_ms(e0.elapsed_time(e1), device=dev)
ts.close()
same = torch.stack([lpq, lpt]); ref = same.clone()
dist.broadcast(ref, src=0)
ident = bool(torch.equal(same, ref))
if rank == 0:
    sub = np.random.RandomState(0).choice(m, 256, replace=False)
    b = brute64.score(z, w, h, q[sub], block=64)
    near = b >= b.max() - 100
    print("world=%d query-sharded == single GPU bit-identical: %s" % (world, bool(torch.equal(lpq, lp1))))
    print("world=%d train-sharded vs single: max|d|=%.2e (within 100 nats %.2e); vs brute64 on 256 rows: %.2e; "
          "identical on all ranks: %s; %.1f ms for %d x %d pairs" % (
              world, float((lpt - lp1).abs().max()),
              float((lpt - lp1).abs()[lp1 >= lp1.max() - 100].max()),
              float(np.abs(lpt.cpu().numpy()[sub] - b)[near].max()), ident, ms, n, m))
    assert torch.equal(lpq, lp1) and ident
dist.barrier()
dist.destroy_process_group()

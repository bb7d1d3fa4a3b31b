"""BASELINE config 5: the scaling sweep N_train 1e4..4e6 x N_query 1e5..1e9 in 8-D, under torchrun on 1..8 GPUs.

Query-sharded for every N (each rank scores its block of the queries against a replica of the KDE; queries are Philox
draws from the KDE generated on the device, batch by batch); training-sharded for the top N (each rank holds N / G rows,
partial log-sum-exp states combined with two all-reduces per batch).  Prints one JSON document on rank 0.

    torchrun --nproc-per-node 8 tools/run_sweep.py [max_queries] > profiles/rXX_sweep_8gpu.json
"""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from skysampler_b200 import GaussianKDE, synth
from skysampler_b200 import dist as skd

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
MAXQ = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 9
D, H, BATCH = 8, 0.1, 1 << 20
rows = []
for N in (10000, 100000, 1000000, 4000000):
    z = w = None
    if rank == 0:
        x = synth.mixture_big(D, N, 500 + int(np.log2(N)))
        w = synth.wide_weights(N, 501)
        m1, pm, comps, std2 = synth.whiten_params(x[::max(1, N // 200000)])
        z = np.ascontiguousarray((x - m1) @ comps.T / std2)
    z = skd.broadcast_array(z, 0, dev); w = skd.broadcast_array(w, 0, dev)
    kde = GaussianKDE(H, device=local).fit(z, sample_weight=w)
    out = torch.empty(BATCH, dtype=torch.float64, device=dev)
    for M in (10 ** 5, 10 ** 6, 10 ** 7, 10 ** 8, 10 ** 9):
        if M > MAXQ:
            continue
        a, b = skd.shard_bounds(M, world, rank)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        lsum = torch.zeros((), dtype=torch.float64, device=dev)
        for s in range(a, b, BATCH):
            m = min(BATCH, b - s)
            q, _, _ = kde.sample_philox(m, seed=77, offset=s, device_out=True)
            kde.score_samples(q, out=out[:m])
            lsum += out[:m].sum()
        e1.record(); torch.cuda.synchronize()
        ms = skd.timed_max_ms(e0.elapsed_time(e1), device=dev)
        tot = skd.allreduce_sum_scalar(float(lsum.item()), device=dev)
        if rank == 0:
            rows.append(dict(mode="query-sharded", n_train=N, n_query=M, n_gpus=world, seconds=ms / 1e3,
                             pairs_per_s=float(N) * M / ms * 1e3, mean_logp=tot / M))
    kde.close()
    if N == 4000000 and world > 1:
        a, b = skd.shard_bounds(N, world, rank)
        part = GaussianKDE(H, device=local).fit(z[a:b], sample_weight=w[a:b])
        full = GaussianKDE(H, device=local).fit(z, sample_weight=w)          # only to generate the same queries on every rank
        sum_w = float(w.sum())
        for M in (10 ** 6, 10 ** 7, 10 ** 8):
            if M > MAXQ:
                continue
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            lsum = torch.zeros((), dtype=torch.float64, device=dev)
            for s in range(0, M, BATCH):
                m = min(BATCH, M - s)
                q, _, _ = full.sample_philox(m, seed=78, offset=s, device_out=True)
                pm_, ps_ = part.score_partial(q)
                pm_, ps_ = skd.combine_partial_lse(pm_, ps_)
                lsum += skd.finish_logp(pm_, ps_, D, H, sum_w).sum()
            e1.record(); torch.cuda.synchronize()
            ms = skd.timed_max_ms(e0.elapsed_time(e1), device=dev)
            if rank == 0:
                rows.append(dict(mode="train-sharded", n_train=N, n_query=M, n_gpus=world, seconds=ms / 1e3,
                                 pairs_per_s=float(N) * M / ms * 1e3, mean_logp=float(lsum.item()) / M))
        part.close(); full.close()
if rank == 0:
    print(json.dumps({"what": "config 5 sweep, 8-D weighted KDE, h = 0.1, queries = Philox draws from the KDE (generated inside the timed region)",
                      "n_gpus": world, "runs": rows}, indent=1))
if world > 1:
    dist.barrier(); dist.destroy_process_group()

"""The N>1 host logic on CPU: world_size-2 gloo process groups, partials from the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import brute64
from skysampler_b200 import dist as skd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _data():
    rs = np.random.RandomState(5)
    z = rs.normal(size=(2001, 4)); w = rs.uniform(0.2, 2.0, 2001); q = 1.5 * rs.normal(size=(333, 4))
    w[:40] = 0.0
    return z, w, q


def _train_sharded(rank, world):
    z, w, q = _data()
    a, b = skd.shard_bounds(len(z), world, rank)
    m, s = brute64.partial_lse(z[a:b], w[a:b], 0.15, q)
    m = torch.from_numpy(np.where(np.isfinite(m), m, 0.0)); s = torch.from_numpy(s)
    m, s = skd.combine_partial_lse(m, s)
    return skd.finish_logp(m, s, 4, 0.15, w.sum()).numpy()


def test_training_sharded_combine_equals_full_score():
    z, w, q = _data()
    full = brute64.score(z, w, 0.15, q)
    for world in (2, 3):
        outs = _run(_train_sharded, world)
        for o in outs:
            np.testing.assert_allclose(o, full, rtol=0, atol=1e-11)
        assert all(np.array_equal(outs[0], o) for o in outs[1:])     # identical on every rank


def _weightless_shard(rank, world):
    z, w, q = _data()
    w = w.copy(); w[:1100] = 0.0                                    # rank 0's whole shard carries no weight
    a, b = skd.shard_bounds(len(z), world, rank)
    m, s = brute64.partial_lse(z[a:b], w[a:b], 0.15, q)
    m = torch.from_numpy(np.where(np.isfinite(m), m, 123.0)); s = torch.from_numpy(s)
    m, s = skd.combine_partial_lse(m, s)
    return skd.finish_logp(m, s, 4, 0.15, w.sum()).numpy()


def test_combine_ignores_weightless_shards():
    z, w, q = _data()
    w = w.copy(); w[:1100] = 0.0
    full = brute64.score(z, w, 0.15, q)
    for o in _run(_weightless_shard, 2):
        np.testing.assert_allclose(o, full, rtol=0, atol=1e-11)


def _query_sharded(rank, world):
    z, w, q = _data()
    a, b = skd.shard_bounds(len(q), world, rank)
    lp = torch.from_numpy(brute64.score(z, w, 0.15, q[a:b]))
    rows = torch.from_numpy(q[a:b][lp.numpy() > -3.0])               # variable-length "accepted rows"
    return skd.gather_concat(lp).numpy(), skd.gather_concat(rows).numpy(), skd.timed_max_ms(1.0 + rank)


def test_query_sharded_gather_is_bit_identical_to_one_rank():
    z, w, q = _data()
    full = brute64.score(z, w, 0.15, q)
    for world in (2, 3):
        for lp, rows, tmax in _run(_query_sharded, world):
            assert np.array_equal(lp, full)
            assert np.array_equal(rows, q[full > -3.0])
            assert tmax == float(world)


def _bcast(rank, world):
    arr = np.arange(24, dtype=np.float64).reshape(6, 4) if rank == 0 else None
    out = skd.broadcast_array(arr, src=0)
    tot = skd.allreduce_sum_scalar(rank + 1.0)
    return out, tot


def test_broadcast_and_scalar_reduce():
    for out, tot in _run(_bcast, 2):
        assert np.array_equal(out, np.arange(24, dtype=np.float64).reshape(6, 4)) and tot == 3.0


def test_shard_bounds_tile_the_range():
    for n in (0, 1, 7, 1000, 10 ** 7 + 3):
        for ws in (1, 2, 3, 8):
            b = [skd.shard_bounds(n, ws, r) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [y - x for x, y in b]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_paths_are_no_ops():
    m = torch.zeros(3, dtype=torch.float64); s = torch.ones(3, dtype=torch.float64)
    assert skd.combine_partial_lse(m, s)[0] is m
    assert skd.gather_concat(m) is m and skd.world() == (1, 0)

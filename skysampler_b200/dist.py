"""Multi-GPU plumbing of the KDE path: one process per GPU, ``torch.distributed`` (NCCL over NVLink on
the box, gloo in the CPU tests).

The reference's only parallelism is data-parallel over query chunks with the training sets copied
into every worker (``run_scores``: one OS process per chunk, emulator.py:698-712).  The equivalent
here:

* **query sharding** (default): rank g scores rows ``shard_bounds(M, G, g)`` of the proposal table
  against its own replica of the KDE; there is NO communication during scoring and results are
  bit-identical to one GPU.  Accepted rows / log-densities are gathered at the end.
* **training sharding** (training set too large for one GPU, or the top of the scaling sweep): rank g
  holds training rows ``shard_bounds(N, G, g)``, scores ALL queries of a block against them and emits
  the per-query partial log-sum-exp state ``(m_g, s_g)``; the combine is
  ``m = allreduce_max(m_g)``, ``s = allreduce_sum(s_g * exp(m_g - m))``.

Every function takes torch tensors on whatever device the process group works on; none of them
computes densities (that is ``GaussianKDE`` / libskb), so the same code is exercised by the gloo
tests with oracle-made partials.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous block [a, b) of rank ``rank``; blocks differ by at most one row and tile [0, n)."""
    a = (n * rank) // world_size
    b = (n * (rank + 1)) // world_size
    return int(a), int(b)


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def broadcast_array(arr, src=0, device=None, group=None):
    """Replicate a numpy array from ``src`` to every rank (the one-off training-set broadcast)."""
    ws, rank = world(group)
    if ws == 1:
        return arr
    device = device or torch.device("cpu")
    if rank == src:
        a = np.ascontiguousarray(arr)
        meta = torch.tensor([a.ndim] + list(a.shape) + [0] * (4 - a.ndim) + [_DT.index(str(a.dtype))],
                            dtype=torch.int64, device=device)
    else:
        meta = torch.zeros(6, dtype=torch.int64, device=device)
    dist.broadcast(meta, src=src, group=group)
    meta = meta.cpu().tolist()
    shape = tuple(meta[1:1 + meta[0]])
    dtype = np.dtype(_DT[meta[5]])
    if rank == src:
        t = torch.from_numpy(a).to(device)
    else:
        t = torch.empty(shape, dtype=_torch_dtype(dtype), device=device)
    dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy()


_DT = ["float64", "float32", "int64", "int32", "uint8", "bool"]


def _torch_dtype(dt):
    return {"float64": torch.float64, "float32": torch.float32, "int64": torch.int64, "int32": torch.int32,
            "uint8": torch.uint8, "bool": torch.bool}[str(dt)]


def combine_partial_lse(m, s, group=None):
    """Merge per-rank partial log-sum-exp states in place.

    ``m``, ``s``: float64 tensors [M] with ``sum_j w_j K_j (this rank's rows) = s * exp(m)``.
    Returns ``(m_all, s_all)`` identical on every rank.  Two collectives of 8*M bytes each.
    """
    ws, _ = world(group)
    if ws == 1:
        return m, s
    # ranks whose shard carries no weight report s == 0 (and an arbitrary m): exclude them from the max
    m_eff = torch.where(s > 0, m, torch.full_like(m, -float("inf")))
    m_all = m_eff.clone()
    dist.all_reduce(m_all, op=dist.ReduceOp.MAX, group=group)
    scale = torch.exp(torch.where(torch.isfinite(m_eff), m_eff - m_all, torch.full_like(m, -float("inf"))))
    s_all = s * scale
    dist.all_reduce(s_all, op=dist.ReduceOp.SUM, group=group)
    return m_all, s_all


def finish_logp(m, s, ndim, bandwidth, sum_weight_total):
    """log p from a combined partial state: log_knorm + m + log s - log(sum of ALL weights)
    (sklearn _binary_tree.pxi.tp:448-478, _kde.py:273-287)."""
    log_knorm = -0.5 * ndim * float(np.log(2.0 * np.pi)) - ndim * float(np.log(bandwidth))
    return log_knorm + m + torch.log(s) - float(np.log(sum_weight_total))


def allreduce_sum_scalar(x, device=None, group=None):
    ws, _ = world(group)
    if ws == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def gather_concat(t, group=None):
    """Variable-length all-gather along dim 0, concatenated in rank order (accepted rows, scores).

    Counts are exchanged first (8 bytes per rank), then one padded all_gather moves the rows."""
    ws, _ = world(group)
    if ws == 1:
        return t
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts) if counts else 0
    pad_shape = (nmax,) + tuple(t.shape[1:])
    buf = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
    buf[:t.shape[0]] = t
    outs = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(outs, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(outs, counts)], dim=0)


def timed_max_ms(ms, device=None, group=None):
    """Max over ranks of a device-measured duration (never wall clock)."""
    ws, _ = world(group)
    if ws == 1:
        return float(ms)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device or torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class ShardedKDE(object):
    """A Gaussian KDE spread over the process group.

    mode="query":  every rank fits the FULL training set; ``score_samples`` scores this rank's query
                   block and (optionally) gathers.  mode="train": every rank fits its row block only;
                   ``score_samples`` scores all queries and combines partial LSE states.
    """

    def __init__(self, bandwidth, mode="query", group=None, device=None):
        from .kde import GaussianKDE
        self._cls = GaussianKDE
        self.bandwidth = float(bandwidth)
        self.mode = mode
        self.group = group
        self.device = device
        self.kde = None

    def fit(self, z, sample_weight=None):
        ws, rank = world(self.group)
        z = np.asarray(z)
        self.ndim = z.shape[1]
        w = None if sample_weight is None else np.asarray(sample_weight, dtype=np.float64)
        self.sum_weight_total = float(len(z)) if w is None else float(w.sum())
        if self.mode == "train":
            a, b = shard_bounds(len(z), ws, rank)
            z = z[a:b]
            w = None if w is None else w[a:b]
        self.kde = self._cls(self.bandwidth, device=self.device).fit(z, sample_weight=w)
        return self

    def score_samples(self, q, gather=True):
        """q: torch CUDA tensor or numpy [M, d] (whitened).  Returns float64 log p (torch tensor)."""
        ws, rank = world(self.group)
        dev = torch.device("cuda", self.kde.device_)
        if not torch.is_tensor(q):
            q = torch.from_numpy(np.ascontiguousarray(q)).to(dev)
        if self.mode == "train":
            m, s = self.kde.score_partial(q)
            m, s = combine_partial_lse(m, s, group=self.group)
            return finish_logp(m, s, self.ndim, self.bandwidth, self.sum_weight_total)
        a, b = shard_bounds(q.shape[0], ws, rank)
        lp = self.kde.score_samples(q[a:b].contiguous())
        return gather_concat(lp, group=self.group) if gather else lp

    def close(self):
        if self.kde is not None:
            self.kde.close()
            self.kde = None

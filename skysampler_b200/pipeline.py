"""In-process batched driver for the rejection step — the next-row replacement of the
reference's ``run_scores`` process pool + offline ``read_concentric`` accept (emulator.py:698-748).

One ``RatioSampler.step`` = score a batch of proposal rows against every KDE of the group, form
``log p_ref - log p_proposal - log m`` and the accept decision in the scoring kernel's epilogue
(``skb_ratio_accept``), then pack the accepted rows in input order (``skb_compact``).  PyTorch
supplies device buffers, pinned host buffers and the stream; the arithmetic is all in libskb.
"""
import ctypes as C

import os

import numpy as np
import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr


class RatioSampler(object):
    def __init__(self, kdes, signs, log_abs_jac=None, log_m=0.0, band=2e-4, cols=None):
        """kdes: fitted GaussianKDE objects on one device; signs[k] = +1 (reference density) / -1 (proposal).
        cols[k]: column indices of the proposal table feeding KDE k (None -> first d columns)."""
        self.kdes = list(kdes)
        self.K = len(self.kdes)
        self.device = torch.device("cuda", self.kdes[0].device_)
        self.handles = (C.c_void_p * self.K)(*[k._handle() for k in self.kdes])
        self.sign = np.ascontiguousarray(signs, dtype=np.int32)
        self.log_abs_jac = np.ascontiguousarray(np.zeros(self.K) if log_abs_jac is None else log_abs_jac, dtype=np.float64)
        self.log_m = float(log_m)
        self.band = float(band)
        self.cols = None
        if cols is not None:
            self.cols = np.full((self.K, _cabi.SKB_MAXD), -1, dtype=np.int32)
            for k, cs in enumerate(cols):
                self.cols[k, :len(cs)] = cs
        self.pairs_per_row = int(sum(k.n_train_ for k in self.kdes))
        self._buf = {}

    def _get(self, name, shape, dtype):
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._buf[name] = t
        return t

    def score_accept(self, q, u, want_logp=False, want_ratio=False, stream=None):
        """Device tensors in, device tensors out; asynchronous on ``stream`` (default: current).  The returned tensors are
        buffers owned by this object: they are valid until the next call that produces the same output."""
        M, ld = q.shape
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        accept = self._get("accept", (M,), torch.uint8)
        in_band = self._get("in_band", (M,), torch.uint8)
        logp = self._get("logp", (self.K, M), torch.float64) if want_logp else None
        ratio = self._get("ratio", (M,), torch.float64) if want_ratio else None
        check(_cabi.lib().skb_ratio_accept(self.handles, self.K, ptr(self.sign), ptr(self.log_abs_jac), self.log_m,
                                           ptr(q), _cabi.dtype_code(q), M, ld, ptr(self.cols), ptr(u), self.band,
                                           ptr(logp), ptr(ratio), ptr(accept), ptr(in_band), stream_ptr(st)))
        return accept, in_band, logp, ratio

    def compact(self, accept, rows, stream=None, sync_count=True, out=None, count_out=None):
        """Accepted rows in input order (``samples[inds]``, emulator.py:746-747).

        ``count_out``: a 1-element int64 CUDA tensor that receives the number of accepted rows without any host
        synchronisation (the call stays asynchronous); otherwise ``sync_count=True`` returns it as a Python int."""
        M, ld = rows.shape
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        if out is None:
            out = self._get("packed", (M, ld), rows.dtype)
        cnt = C.c_int64(0)
        if count_out is not None:
            n_ptr = C.cast(C.c_void_p(count_out.data_ptr()), C.POINTER(C.c_int64))
        else:
            n_ptr = C.byref(cnt) if sync_count else None
        check(_cabi.lib().skb_compact(ptr(accept), M, ptr(rows), _cabi.dtype_code(rows), ld, ptr(out), None,
                                      n_ptr, stream_ptr(st)))
        return out, int(cnt.value)

    def draw(self, k, n, seed, offset=0, out=None, stream=None):
        """``n`` Philox KDE draws from member ``k`` of the group straight into a device table (raw space when that KDE
        carries an inverse whitening map) — the proposal step of the pipeline (emulator.py:530-544), asynchronous."""
        kde = self.kdes[k]
        if out is None:
            out = self._get("draw%d" % k, (n, kde.ndim_), torch.float64)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(_cabi.lib().skb_draw(kde._handle(), n, None, None, int(seed), int(offset), None, -1, 0.0, 0.0, 1, None,
                                   ptr(out), kde.ndim_, None, None, None, stream_ptr(st)))
        return out

    def score_accept_host(self, q_host, u_host, accept_host, logp_host=None, stream=None):
        """The C-ABI call with HOST buffers (numpy arrays or pinned torch tensors): rows + uniforms in, accept flags
        (and optionally the log-densities) out; the library stages the copies itself.  Synchronous."""
        M, ld = q_host.shape
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(_cabi.lib().skb_ratio_accept(self.handles, self.K, ptr(self.sign), ptr(self.log_abs_jac), self.log_m,
                                           ptr(q_host), _cabi.dtype_code(q_host), M, ld, ptr(self.cols), ptr(u_host), self.band,
                                           ptr(logp_host), None, ptr(accept_host), None, stream_ptr(st)))
        return accept_host

    def step(self, q, u):
        """Device-resident batch: returns (accepted rows [n_acc, ld] view, n_acc)."""
        accept, _, _, _ = self.score_accept(q, u)
        out, n = self.compact(accept, q)
        return out[:n], n

    def step_host(self, q_host, u_host, flags_host=None):
        """End-to-end batch from (pinned) host memory: H2D of rows + uniforms, fused score/accept,
        compaction, D2H of the accept flags and the accepted rows' count."""
        M, ld = q_host.shape
        qd = self._get("q_in", (M, ld), q_host.dtype)
        ud = self._get("u_in", (M,), torch.float64)
        qd.copy_(q_host, non_blocking=True)
        ud.copy_(u_host, non_blocking=True)
        accept, _, _, _ = self.score_accept(qd, ud)
        out, n = self.compact(accept, qd)
        if flags_host is None:
            flags_host = torch.empty(M, dtype=torch.uint8, pin_memory=True)
        flags_host.copy_(accept, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return flags_host, n, out[:n]

    def run_host(self, batches):
        """Pipelined end-to-end pass over a sequence of (rows, uniforms) batches in PINNED host memory:
        the H2D copy of batch i+1 runs on a side stream while batch i is scored, accept flags and accepted
        counts stream back asynchronously, one synchronisation at the end.  Returns (list of pinned uint8
        flag tensors, list of accepted counts); the flag tensors are views of a pinned buffer that the next
        call reuses."""
        batches = list(batches)
        if not batches:
            return [], []
        main = torch.cuda.current_stream(self.device)
        side = getattr(self, "_side", None)
        if side is None:
            side = self._side = torch.cuda.Stream(device=self.device)
        M = max(b[0].shape[0] for b in batches)          # slots sized by the largest batch; shorter ones use a prefix
        ld = batches[0][0].shape[1]
        slots = []
        for s in range(2):
            slots.append((self._get("q_pipe%d" % s, (M, ld), batches[0][0].dtype), self._get("u_pipe%d" % s, (M,), torch.float64),
                          self._get("acc_pipe%d" % s, (M,), torch.uint8)))
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        # pinned result buffers are kept between calls (cudaHostAlloc costs milliseconds and blocks the host)
        pinned = getattr(self, "_pinned_out", None)
        if pinned is None or pinned[0].shape[0] < len(batches) or pinned[0].shape[1] != M:
            pinned = self._pinned_out = (torch.empty((len(batches), M), dtype=torch.uint8, pin_memory=True),
                                         torch.zeros(len(batches), dtype=torch.int64, pin_memory=True))
        flags_host = [pinned[0][i] for i in range(len(batches))]
        counts_host = pinned[1][:len(batches)]

        def upload(i):
            qd, ud, _ = slots[i % 2]
            with torch.cuda.stream(side):
                if i >= 2:
                    side.wait_event(free[i % 2])              # the kernels of batch i-2 are done with this slot
                m_i = batches[i][0].shape[0]
                qd[:m_i].copy_(batches[i][0], non_blocking=True)
                ud[:m_i].copy_(batches[i][1], non_blocking=True)
                ready[i % 2].record(side)

        trace = [] if os.environ.get("SKB_TRACE_RUN_HOST") else None      # debug: per-batch device timeline on stderr
        if trace is not None:
            import sys, time
            t_h0 = time.perf_counter()
        upload(0)
        if trace is not None:
            sys.stderr.write("run_host: first upload enqueued after %.2f ms on the host\n" % ((time.perf_counter() - t_h0) * 1e3))
        for i in range(len(batches)):
            if i + 1 < len(batches):
                upload(i + 1)
            qd, ud, keep = slots[i % 2]
            if trace is not None:
                trace.append([torch.cuda.Event(enable_timing=True) for _ in range(3)])
                trace[-1][0].record(main)
            main.wait_event(ready[i % 2])
            if trace is not None:
                trace[-1][1].record(main)
            m_i = batches[i][0].shape[0]
            qd, ud, keep = qd[:m_i], ud[:m_i], keep[:m_i]
            accept, _, _, _ = self.score_accept(qd, ud, stream=main)
            if trace is not None:
                trace[-1][2].record(main)
                if i == 0:
                    sys.stderr.write("run_host: first score enqueued after %.2f ms on the host\n" % ((time.perf_counter() - t_h0) * 1e3))
            keep.copy_(accept)
            self.compact(keep, qd, stream=main, sync_count=False)
            free[i % 2].record(main)
            flags_host[i][:m_i].copy_(keep, non_blocking=True)
            counts_host[i:i + 1].copy_(keep.sum(dtype=torch.int64).reshape(1), non_blocking=True)
        main.synchronize()
        if trace is not None:
            for i, (a, b, c) in enumerate(trace):
                sys.stderr.write("run_host batch %d: wait for rows %.2f ms, score %.2f ms, since first batch %.2f ms\n" % (
                    i, a.elapsed_time(b), b.elapsed_time(c), trace[0][0].elapsed_time(c)))
        return flags_host, [int(c) for c in counts_host]

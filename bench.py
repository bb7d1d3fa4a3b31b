#!/usr/bin/env python
"""Headline benchmark: KDE kernel-pairs/s (and accepted LOS samples/s) on BASELINE.json config 2.

Workload ("c2_widefield_ratio"): two weighted Gaussian KDEs in an 8-D feature space — cluster and
field, 10^6 synthetic training galaxies each, bandwidth 0.1, each with its own whitening — score
candidate rows drawn from the field KDE (10 % inflated x2 for tail coverage); a candidate is accepted
when u < p_cluster / (m p_field).  One STEP = one batch of 2^20 candidate rows through the hot path:
fused whitening + scoring of both KDEs + ratio + accept in one launch (skb_ratio_accept), then
order-preserving compaction of the accepted rows (skb_compact).  The candidate pool holds 10 batches
(1.05e7 rows, ~10^7 as BASELINE names) and steps rotate through it, so consecutive steps read different
rows (pool 671 MB + training tiles 72 MB > the 126 MB L2; the path is FMA-bound anyway).

  value   pairs/s with the pool resident in HBM (CUDA events, max over ranks)
  e2e     pairs/s through RatioSampler.run_host: pinned host rows + uniforms -> H2D (overlapped) -> kernels -> D2H flags
  roofline  the score kernel against the FP32-FMA/SFU pipe peaks measured live on the same GPU
            (pairs/s <= min(FMA lane-ops/s / (2d+1), EX2/s); HBM is not the bound of this kernel)
  cpu_baseline / --impl reference   the reference's own engine for this path — sklearn KernelDensity
            KD-tree at the shipped atol=rtol=1e-6 behind the restated KDEContainer (oracle/container.py),
            one process per host core, on a bounded sample of the same workload.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torchrun (one rank per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D = 8
H = 0.1
N_TRAIN = 1000000
BATCH = 1 << 20
N_POOL_BATCHES = 10
METRIC = "kde_kernel_pairs_per_s"
UNIT = "pairs/s"


def workload_config(args, n_gpus):
    return {
        "workload": "c2_widefield_ratio: 2 weighted KDEs (cluster, field) x %d train rows, d=%d, h=%g; "
                    "%d candidate rows per step per GPU from a %d-batch pool; fused score+ratio+accept+compact"
                    % (args.ntrain, D, H, args.batch, N_POOL_BATCHES),
        "n_train_per_kde": args.ntrain, "n_kde": 2, "d": D, "bandwidth": H,
        "batch_rows_per_gpu": args.batch, "pool_rows_per_gpu": args.batch * N_POOL_BATCHES,
        "sharding": "query blocks per rank, training set replicated (NCCL broadcast at fit)",
        "l2_policy": "inputs larger than L2 (10-batch pool rotated between steps)",
        "n_gpus": n_gpus,
    }


# ------------------------------------------------------------------------------------------------
# synthetic problem (shared by both arms)
# ------------------------------------------------------------------------------------------------
def make_training(ntrain):
    from skysampler_b200 import synth
    out = {}
    # cluster and field are two samples of the SAME 8-component mixture, the field one with covariance
    # x1.2: overlapping populations (a cluster line of sight against its field) so that the density
    # ratio is bounded and the accept rate is a few per cent, as in the reference's use (§3.5).
    for name, seed, cov in (("cluster", 201, 1.0), ("field", 202, 1.2)):
        x = synth.mixture(D, ntrain, 201, cov_scale=cov, sample_seed=seed)
        w = synth.wide_weights(ntrain, seed + 7919)
        mean1, pmean, comps, std2 = synth.whiten_params(x, w)
        mu = mean1 + pmean
        W = comps.T / std2[None, :]
        Winv = std2[:, None] * comps
        out[name] = dict(x=x, w=w, mean1=mean1, pmean=pmean, comps=comps, std2=std2, mu=mu, W=W, Winv=Winv,
                         z=(x - mu) @ W, log_abs_jac=float(np.log(abs(np.linalg.det(Winv)))))
    return out


def host_candidates(field, m, seed):
    """Field-KDE draws in raw space (a10 semantics), 10 % inflated x2 about the field mean."""
    rs = np.random.RandomState(seed)
    idx = rs.randint(0, len(field["z"]), size=m)
    zq = field["z"][idx] + H * rs.normal(size=(m, D))
    x = zq @ field["Winv"] + field["mu"]
    nt = m // 10
    x[:nt] = field["mu"] + 2.0 * (x[:nt] - field["mu"])
    return np.ascontiguousarray(x)


# ------------------------------------------------------------------------------------------------
# reference arm: sklearn KD-tree behind the restated KDEContainer, one process per core
# ------------------------------------------------------------------------------------------------
_W = {}


def _ref_init(ntrain, atol, rtol):
    import pandas as pd
    from oracle.container import RefKDEContainer
    tr = make_training(ntrain)
    cols = ["F%d" % i for i in range(D)]
    for name in ("cluster", "field"):
        t = tr[name]
        c = RefKDEContainer(pd.DataFrame(t["x"], columns=cols), pd.Series(t["w"]), seed=1, atol=atol, rtol=rtol)
        c.set_whitening(t["mean1"], t["pmean"], t["comps"], t["std2"])
        c.construct_kde(H)
        _W[name] = c
    _W["cols"] = cols
    return True


def _ref_score(q):
    import pandas as pd
    qdf = pd.DataFrame(q, columns=_W["cols"])
    a, ja = _W["cluster"].score_samples(qdf)
    b, jb = _W["field"].score_samples(qdf)
    return len(q), float(a[0] - b[0])


def _ref_ready(_):
    return os.getpid()


def run_reference(args):
    import multiprocessing as mp
    P = args.cpu_procs or (os.cpu_count() or 1)
    ctx = mp.get_context("fork")
    t0 = time.time()
    pool = ctx.Pool(P, initializer=_ref_init, initargs=(args.ntrain, 1e-6, 1e-6))
    pool.map(_ref_ready, range(P * 4))
    fit_s = time.time() - t0
    tr = make_training(args.ntrain)
    pairs_per_row = 2 * args.ntrain
    # calibrate rows per process per step so that the whole run stays bounded
    probe = host_candidates(tr["field"], 4 * P, 900)
    t0 = time.time()
    pool.map(_ref_score, [probe[i * 4:(i + 1) * 4] for i in range(P)])
    per_row = max((time.time() - t0) / 4.0, 1e-4)
    budget = args.cpu_seconds / float(max(args.steps + args.warmup, 1))
    msub = int(max(2, min(4096, budget / per_row)))
    times = []
    for s in range(args.warmup + args.steps):
        q = host_candidates(tr["field"], msub * P, 1000 + s)
        t0 = time.time()
        pool.map(_ref_score, [q[i * msub:(i + 1) * msub] for i in range(P)])
        dt = time.time() - t0
        if s >= args.warmup:
            times.append(dt)
    pool.close()
    pool.join()
    total = float(sum(times))
    pairs = float(msub * P * pairs_per_row * len(times))
    value = pairs / total
    sample = ("%d steps x %d procs x %d candidate rows scored against both %d-row KD-trees (tree fit %.1f s "
              "per process excluded); sklearn %s KernelDensity atol=rtol=1e-6, weighted, depth-first"
              % (len(times), P, msub, args.ntrain, fit_s, __import__("sklearn").__version__))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(len(times), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": P, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.time(), ln.strip()))

    def wait_ready(self, timeout=8.0):
        """Block until nvidia-smi has delivered its first sample: its start-up attaches to the driver and can stall
        work submission for tens of milliseconds, which must not land inside a timed region."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.05)
        time.sleep(0.1)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except Exception:
                continue
            for nme, val in zip(names, f[3:7]):
                if val == "Active":
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from skysampler_b200 import GaussianKDE, _cabi
    from skysampler_b200 import dist as skd
    from skysampler_b200.pipeline import RatioSampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL prints its version banner there at any NCCL_DEBUG level
        if "SKB_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["SKB_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()

    # ---- fit: rank 0 builds the training sets, one NCCL broadcast replicates them ----
    tr = make_training(args.ntrain) if rank == 0 else None
    kdes, ljac, fieldp = [], [], None
    for name in ("cluster", "field"):
        keys = ("z", "w", "mu", "W", "Winv")
        arrs = {k: skd.broadcast_array(tr[name][k] if rank == 0 else None, src=0, device=dev) for k in keys}
        lj = float(np.log(abs(np.linalg.det(arrs["Winv"]))))
        kde = GaussianKDE(H, device=local).fit(arrs["z"], sample_weight=arrs["w"],
                                                whitening=(arrs["mu"], arrs["W"], arrs["Winv"]), is_whitened=True)
        kdes.append(kde); ljac.append(lj)
        if name == "field":
            fieldp = arrs
    del tr

    # ---- candidate pool on the device: Philox draws from the field KDE, 10 % inflated ----
    pool_rows = args.batch * N_POOL_BATCHES
    pool, _, _ = kdes[1].sample_philox(pool_rows, seed=402 + rank, device_out=True)
    mu_f = torch.from_numpy(fieldp["mu"]).to(dev)
    for b in range(N_POOL_BATCHES):
        sl = pool[b * args.batch: b * args.batch + args.batch // 10]
        sl.copy_(mu_f + 2.0 * (sl - mu_f))
    u_pool = torch.from_numpy(np.random.RandomState(6 + rank).uniform(0, 1, pool_rows)).to(dev)

    # ---- proposal bound m: 99.9th percentile of the density ratio on a 1e5 pilot ----
    sampler = RatioSampler(kdes, signs=[1, -1], log_abs_jac=ljac, log_m=0.0)
    # (pilot rows are taken from the un-inflated part of the first batch: in the inflated tail both
    #  densities hang on single neighbours and the ratio is noise spanning hundreds of nats)
    p0 = args.batch // 10
    npilot = min(100000, args.batch - p0)
    _, _, _, ratio = sampler.score_accept(pool[p0:p0 + npilot], u_pool[p0:p0 + npilot], want_ratio=True)
    log_m = float(torch.quantile(ratio[torch.isfinite(ratio)], 0.999).item())
    sampler.log_m = log_m
    torch.cuda.synchronize()

    def batch(i):
        b = i % N_POOL_BATCHES
        return pool[b * args.batch:(b + 1) * args.batch], u_pool[b * args.batch:(b + 1) * args.batch]

    def one_step(i, ev=None):
        q, u = batch(i)
        if ev is not None:
            ev[0].record()
        accept, _, _, _ = sampler.score_accept(q, u)
        if ev is not None:
            ev[1].record()
        rows, n = sampler.compact(accept, q)
        if world > 1:                                   # accepted rows gathered over NVLink (NCCL all_gather)
            rows = skd.gather_concat(rows[:n].contiguous())
            return rows.shape[0], n
        return n, n

    # ---- warm-up, then exactly K timed steps ----
    for i in range(args.warmup):
        one_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    clocks.wait_ready()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.skb_launch_count()
    lib.skb_debug_counters(local, (C.c_uint64 * 4)(), 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall0 = time.time()
    e0.record()
    accepted_local = 0
    for i in range(args.steps):
        _, n = one_step(args.warmup + i, kev[i])
        accepted_local += n
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    launches = lib.skb_launch_count() - launches0
    ms_total = skd.timed_max_ms(e0.elapsed_time(e1), device=dev)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    kernel_ms_ranks = [kernel_ms]
    if world > 1:                                        # rank skew diagnostic: each rank's mean score-launch time
        kt = torch.tensor([kernel_ms], dtype=torch.float64, device=dev)
        kts = [torch.zeros_like(kt) for _ in range(world)]
        dist.all_gather(kts, kt)
        kernel_ms_ranks = [round(float(k.item()), 2) for k in kts]
    clk = clocks.stop(t_wall0, t_wall1)
    accepted = skd.allreduce_sum_scalar(accepted_local, device=dev)

    pairs_per_step_rank = float(args.batch) * sampler.pairs_per_row
    value = pairs_per_step_rank * world * args.steps / (ms_total * 1e-3)
    st = (C.c_uint64 * 4)()
    lib.skb_debug_counters(local, st, 1)

    # ---- end-to-end: pinned host rows + uniforms -> H2D -> kernels -> D2H flags ----
    nb_host = min(N_POOL_BATCHES, max(2, args.steps))
    q_host = torch.empty((nb_host, args.batch, D), dtype=torch.float64, pin_memory=True)
    u_host = torch.empty((nb_host, args.batch), dtype=torch.float64, pin_memory=True)
    for b in range(nb_host):
        q_host[b].copy_(pool[b * args.batch:(b + 1) * args.batch])
        u_host[b].copy_(u_pool[b * args.batch:(b + 1) * args.batch])
    flags_host = torch.empty(args.batch, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()
    for b in range(max(args.warmup, nb_host)):            # warm-up: every pinned batch crosses the bus once
        sampler.step_host(q_host[b % nb_host], u_host[b % nb_host], flags_host)
    if world > 1:
        dist.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks2 = ClockSampler(local)
    clocks2.start()
    clocks2.wait_ready()
    e2e_batches = [(q_host[i % nb_host], u_host[i % nb_host]) for i in range(args.steps)]
    sampler.run_host(e2e_batches)                         # warm-up of the pipelined path (side stream, slots, pinned result buffers)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # Three timed passes of exactly K steps each; the MEDIAN is reported and all three are listed.  A single pass
    # was seen to carry a one-off 35-400 ms host-side stall on some boxes (kernel times inside the pass unchanged,
    # see SKB_TRACE_RUN_HOST), which at K = 10 moves the per-step figure by up to 80 %.
    t_e0 = time.time()
    e2e_runs = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        f0.record()
        e2e_flags, e2e_counts = sampler.run_host(e2e_batches)  # H2D of batch i+1 overlaps the scoring of batch i
        f1.record()
        torch.cuda.synchronize()
        e2e_runs.append(skd.timed_max_ms(f0.elapsed_time(f1), device=dev))
    t_e1 = time.time()
    clk2 = clocks2.stop(t_e0, t_e1)
    e2e_ms = float(np.median(e2e_runs))
    # host-path diagnostic: one pinned batch host -> device on an idle GPU (explains e2e on boxes with a slow host path)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    qd_probe = torch.empty((args.batch, D), dtype=torch.float64, device=dev)
    qd_probe.copy_(q_host[0], non_blocking=True)
    torch.cuda.synchronize()
    h0.record()
    for b in range(3):
        qd_probe.copy_(q_host[b % nb_host], non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * args.batch * D * 8 / (h0.elapsed_time(h1) * 1e-3) / 1e9
    del qd_probe
    e2e_value = pairs_per_step_rank * world * args.steps / (e2e_ms * 1e-3)
    h2d = args.batch * D * 8 + args.batch * 8
    d2h = args.batch + 8

    # ---- the same kernel with culling off (every pair evaluated): the FMA-pipe efficiency of the arithmetic ----
    dense_ms = None
    if rank == 0:
        os.environ["SKB_NO_CULL"] = "1"
        sampler.score_accept(*batch(0))
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for i in range(2):
            sampler.score_accept(*batch(1 + i))
        d1.record()
        torch.cuda.synchronize()
        dense_ms = d0.elapsed_time(d1) / 2.0
        os.environ.pop("SKB_NO_CULL", None)
        lib.skb_debug_counters(local, (C.c_uint64 * 4)(), 1)

    # ---- roofline of the dominant kernel (skb_score_pq_kernel<8>), peaks measured live on this GPU ----
    roof = None
    if rank == 0:
        fma, ex2 = C.c_double(), C.c_double()
        _cabi.check(lib.skb_measure_pipes(local, 300.0, C.byref(fma), C.byref(ex2)))
        peak_pairs = min(fma.value / (2 * D + 1), ex2.value)
        ach = pairs_per_step_rank / (kernel_ms * 1e-3)
        traffic = None
        tf = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tf):
            try:
                tj = json.load(open(tf))
                if tj.get("batch") == args.batch and tj.get("ntrain") == args.ntrain:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        flop_pair = 3 * D + 1
        # executed pairs: (query, point) pairs per lane counted at warp level x 32 lanes
        exec_pairs_per_launch = st[1] * 32.0 / max(args.steps, 1)
        exec_frac = exec_pairs_per_launch / pairs_per_step_rank
        ach_exec = exec_pairs_per_launch / (kernel_ms * 1e-3)
        dense_rate = pairs_per_step_rank / (dense_ms * 1e-3)
        roof = {
            # bound: the FP32 FMA pipe (with the SFU as second limit) — not HBM, not tensor cores
            "bound": "fp32", "kernel": "skb::skb_score_pq_kernel<8>",
            "achieved": ach * flop_pair / 1e12, "peak": peak_pairs * flop_pair / 1e12, "unit": "TFLOP/s",
            "frac": ach / peak_pairs, "traffic": traffic,
            "definition": "achieved = ALGORITHMIC work: pairs/launch x (3d+1) FP32 FLOP / CUDA-event launch time, "
                          "every (query, training) pair counted; peak = min(FFMA lane-ops/s / (2d+1), EX2/s) x (3d+1), "
                          "the pair roofline of SURVEY.md §8d in FLOP/s.  frac > 1 is expected: tiles whose box test "
                          "proves that every pair term flushes to zero in fp32 are skipped (executed_pair_fraction); "
                          "frac_executed and frac_dense give the arithmetic's own pipe efficiency",
            "executed_pair_fraction": exec_frac,
            "frac_executed": ach_exec / peak_pairs,
            "frac_dense": dense_rate / peak_pairs, "dense_ms_per_launch": dense_ms,
            "dense_note": "same launch with SKB_NO_CULL=1 (all pairs evaluated), 2 launches after the timed region",
            "achieved_gpairs": ach / 1e9, "peak_gpairs": peak_pairs / 1e9,
            "fp32_fma_peak_tflops_measured": 2.0 * fma.value / 1e12, "ex2_peak_per_s_measured": ex2.value,
            "peak_source": "measured live on this GPU by skb_measure_pipes (FFMA2 and MUFU.EX2 saturation loops, "
                           "300 ms each); MEASURED_PEAKS.json holds only HBM and bf16-tensor figures",
            "kernel_ms_per_launch": kernel_ms, "pairs_per_launch": pairs_per_step_rank,
            "hbm_gbs_peak_for_reference": _hbm_peak(),
            "rerun_blocks_frac": (st[0] / float(st[1])) if st[1] else None,
            "tiles_culled_frac": (st[2] / float(st[3])) if st[3] else None,
        }

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "dtype_note": "fp32 pair terms, fp64 log-sum-exp state",
            "data": "synthetic",
            "config": workload_config(args, world),
            "accepted_samples_per_s": accepted / (ms_total * 1e-3),
            "accept_fraction": accepted / float(args.batch * world * args.steps), "log_m": log_m,
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": e2e_ms / args.steps, "api": "skysampler_b200.pipeline.RatioSampler.run_host (pinned host batches in, accept flags + counts out, H2D overlapped)",
                    "h2d_gbs_idle_gpu": round(h2d_gbs, 2),
                    "ms_per_step_each_pass": [round(r / args.steps, 2) for r in e2e_runs], "value_is": "median of the 3 timed passes",
                    "clocks": clk2},
            "kernel_ms_each_step": [round(a.elapsed_time(b), 2) for a, b in kev], "kernel_ms_mean_per_rank": kernel_ms_ranks,
            "gpu_launches": int(launches),
            "roofline": roof,
        }
    for k in kdes:
        k.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def _hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def cpu_baseline_subprocess(args):
    """Time the reference arm in a child process (no CUDA there), bounded to ~cpu_seconds."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--cpu-seconds", str(args.cpu_seconds), "--ntrain", str(args.ntrain), "--gpus", "1"]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: " + out.stderr[-300:]}
    except Exception as e:  # pragma: no cover
        return {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: %r" % (e,)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--ntrain", type=int, default=N_TRAIN)
    ap.add_argument("--cpu-seconds", type=float, default=None, help="CPU budget of the reference sample")
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if args.cpu_seconds is None:
            args.cpu_seconds = 90.0
        if rank != 0:
            return 0
        print(json.dumps(run_reference(args)), flush=True)
        return 0
    if args.cpu_seconds is None:
        args.cpu_seconds = 20.0
    cpu = None
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_subprocess(args)          # before CUDA is initialised in this process
    line = run_ours(args)
    if rank == 0:
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())

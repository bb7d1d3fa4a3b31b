#!/usr/bin/env python
"""Headline benchmark: KDE kernel-pairs/s (and accepted LOS samples/s) on BASELINE.json config 2.

Workload ("c2_widefield_ratio"): two weighted Gaussian KDEs in an 8-D feature space — cluster and
field, 10^6 synthetic training galaxies each, bandwidth 0.1, each with its own whitening — score
candidate rows drawn from the field KDE (10 % inflated x2 for tail coverage); a candidate is accepted
when u < p_cluster / (m p_field).  One STEP = one batch of 2^20 candidates through the whole pipeline of SURVEY §8d:
Philox draw of the batch from the field KDE (skb_draw; fresh rows every step, so nothing is L2-resident from the step
before) -> fused whitening + scoring of both KDEs + ratio + accept (skb_ratio_accept) -> order-preserving compaction of
the accepted rows (skb_compact) -> at N > 1 an NCCL all-gather of the accepted rows on a side stream.

  value   pairs/s of that pipeline, training sets resident in HBM (CUDA events, max over ranks)
  e2e     pairs/s through the C ABI with HOST buffers: pinned candidate rows + uniforms in, accept flags out
          (skb_ratio_accept stages the copies itself); e2e.plugin = the same work through KDEContainer.score_samples
  roofline  the score kernel against the FP32-FMA/SFU pipe peaks measured live on the same GPU
            (pairs/s <= min(FMA lane-ops/s / (2d+1), EX2/s); HBM is not the bound of this kernel); frac counts the pairs
            the kernel actually evaluates, frac_algorithmic every pair (the metric), frac_dense the same launch unculled
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
                    "%d candidate rows per step per GPU drawn on the device from the field KDE; draw + fused score+ratio+accept "
                    "+ compact (+ gather)" % (args.ntrain, D, H, args.batch),
        "n_train_per_kde": args.ntrain, "n_kde": 2, "d": D, "bandwidth": H,
        "batch_rows_per_gpu": args.batch,
        "sharding": "query blocks per rank, training set replicated (NCCL broadcast at fit)",
        "l2_policy": "inputs larger than L2: every step scores freshly drawn rows (67 MB) against 2 x 40 MB of training tiles "
                     "and their box tables; nothing of a step is reused by the next",
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
def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's own log (whatever NCCL_DEBUG level the launcher set) goes to stderr so that stdout keeps the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from skysampler_b200 import GaussianKDE, _cabi
    from skysampler_b200 import dist as skd
    from skysampler_b200.pipeline import RatioSampler

    world, rank, local, dev = _dist_setup()
    lib = _cabi.lib()

    # ---- fit: rank 0 builds the training sets, one NCCL broadcast replicates them ----
    tr = make_training(args.ntrain) if rank == 0 else None
    kdes, ljac, fieldp, fit_wall_ms, fit_dev_ms = [], [], None, [], []
    for name in ("cluster", "field"):
        keys = ("z", "w", "mu", "W", "Winv")
        arrs = {k: skd.broadcast_array(tr[name][k] if rank == 0 else None, src=0, device=dev) for k in keys}
        lj = float(np.log(abs(np.linalg.det(arrs["Winv"]))))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        kde = GaussianKDE(H, device=local).fit(arrs["z"], sample_weight=arrs["w"],
                                                whitening=(arrs["mu"], arrs["W"], arrs["Winv"]), is_whitened=True)
        fit_wall_ms.append((time.perf_counter() - t0) * 1e3)
        fit_dev_ms.append(kde.fit_ms_)
        kdes.append(kde); ljac.append(lj)
        if name == "field":
            fieldp = arrs
    mu_f = torch.from_numpy(fieldp["mu"]).to(dev)
    tail = args.batch // 10
    sampler = RatioSampler(kdes, signs=[1, -1], log_abs_jac=ljac, log_m=0.0)
    seed = 402 + rank

    def draw_candidates(step, out):
        """One batch of proposals: Philox draws from the field KDE in raw space (a10 semantics, counter offset = step),
        the first 10 % inflated x2 about the field mean (tail coverage)."""
        q = sampler.draw(1, args.batch, seed, offset=step * args.batch, out=out)
        q[:tail].mul_(2.0).sub_(mu_f)
        return q

    u_pool = torch.from_numpy(np.random.RandomState(6 + rank).uniform(0, 1, args.batch * N_POOL_BATCHES)).to(dev)
    cand = [torch.empty((args.batch, D), dtype=torch.float64, device=dev) for _ in range(2)]
    packed = [torch.empty((args.batch, D), dtype=torch.float64, device=dev) for _ in range(2)]
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(2)]
    acc_total = torch.zeros(1, dtype=torch.int64, device=dev)
    acc_max = torch.zeros(1, dtype=torch.int64, device=dev)

    # ---- proposal bound m: 99.9th percentile of the density ratio on a 1e5 pilot (un-inflated draws: in the inflated
    #      tail both densities hang on single neighbours and the ratio is noise spanning hundreds of nats) ----
    pilot = draw_candidates(10 ** 6, cand[0])[tail:tail + min(100000, args.batch - tail)]
    _, _, _, ratio = sampler.score_accept(pilot, u_pool[:pilot.shape[0]], want_ratio=True)
    log_m = float(torch.quantile(ratio[torch.isfinite(ratio)], 0.999).item())
    sampler.log_m = log_m
    torch.cuda.synchronize()

    # ---- accepted rows gathered over NVLink without touching the host: fixed-capacity slots, NCCL on a side stream ----
    cap = max(args.batch // 8, 1024)
    main = torch.cuda.current_stream(dev)
    if world > 1:
        gside = torch.cuda.Stream(device=dev)
        g_rows = [torch.empty((world, cap, D), dtype=torch.float64, device=dev) for _ in range(2)]
        g_cnt = [torch.empty((world, 1), dtype=torch.int64, device=dev) for _ in range(2)]
        gdone = [torch.cuda.Event(), torch.cuda.Event()]

    def one_step(i, ev=None):
        s = i & 1
        if world > 1 and i >= 2:
            main.wait_event(gdone[s])                    # the gather of step i - 2 has read packed[s] / counts[s]
        q = draw_candidates(i, cand[s])
        u = u_pool[(i % N_POOL_BATCHES) * args.batch:(i % N_POOL_BATCHES + 1) * args.batch]
        if ev is not None:
            ev[0].record()
        accept, _, _, _ = sampler.score_accept(q, u)
        if ev is not None:
            ev[1].record()
        sampler.compact(accept, q, out=packed[s], count_out=counts[s])
        acc_total.add_(counts[s])
        torch.maximum(acc_max, counts[s], out=acc_max)
        if world > 1:
            gside.wait_stream(main)
            with torch.cuda.stream(gside):
                dist.all_gather_into_tensor(g_rows[s].view(world * cap, D), packed[s][:cap])
                dist.all_gather_into_tensor(g_cnt[s].view(world), counts[s])
                gdone[s].record(gside)

    # ---- warm-up, then exactly K timed steps ----
    for i in range(args.warmup):
        one_step(i)
    if world > 1:
        main.wait_stream(gside)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    clocks.wait_ready()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc_total.zero_(); acc_max.zero_()
    launches0 = lib.skb_launch_count()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        one_step(args.warmup + i, kev[i])
    if world > 1:
        main.wait_stream(gside)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    launches = lib.skb_launch_count() - launches0
    ms_total = skd.timed_max_ms(e0.elapsed_time(e1), device=dev)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    kernel_ms_ranks = [kernel_ms]
    if world > 1:                                        # rank skew diagnostic: each rank's mean score time
        kt = torch.tensor([kernel_ms], dtype=torch.float64, device=dev)
        kts = [torch.zeros_like(kt) for _ in range(world)]
        dist.all_gather(kts, kt)
        kernel_ms_ranks = [round(float(k.item()), 2) for k in kts]
    clk = clocks.stop(t_wall0, t_wall1)
    accepted = skd.allreduce_sum_scalar(int(acc_total.item()), device=dev)
    if int(acc_max.item()) > cap:
        raise RuntimeError("a step accepted %d rows, more than the gather slot of %d" % (int(acc_max.item()), cap))
    gathered_last = int(g_cnt[(args.warmup + args.steps - 1) & 1].sum().item()) if world > 1 else None

    pairs_per_step_rank = float(args.batch) * sampler.pairs_per_row
    value = pairs_per_step_rank * world * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the C ABI with HOST buffers: pinned candidate rows + uniforms in, accept flags out; the
    #      library stages the copies itself (skb_ratio_accept, what a ctypes stub in the reference would call) ----
    nb_host = min(N_POOL_BATCHES, max(2, args.steps))
    q_host = torch.empty((nb_host, args.batch, D), dtype=torch.float64, pin_memory=True)
    u_host = torch.empty((nb_host, args.batch), dtype=torch.float64, pin_memory=True)
    for b in range(nb_host):
        q_host[b].copy_(draw_candidates(2 * 10 ** 6 + b, cand[0]))
        u_host[b].copy_(u_pool[b * args.batch:(b + 1) * args.batch])
    flags_host = torch.empty(args.batch, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()
    for b in range(max(args.warmup, nb_host)):            # warm-up: every pinned batch crosses the bus once
        sampler.score_accept_host(q_host[b % nb_host], u_host[b % nb_host], flags_host)
    if world > 1:
        dist.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks2 = ClockSampler(local)
    clocks2.start()
    clocks2.wait_ready()
    t_e0 = time.time()
    e2e_runs, e2e_accepted = [], 0
    for _ in range(3):                                    # three passes of K steps, the median is reported
        if world > 1:
            dist.barrier()
        f0.record()
        for i in range(args.steps):
            sampler.score_accept_host(q_host[i % nb_host], u_host[i % nb_host], flags_host)
            e2e_accepted = int(flags_host.sum())          # the step's result is read on the host
        f1.record()
        torch.cuda.synchronize()
        e2e_runs.append(skd.timed_max_ms(f0.elapsed_time(f1), device=dev))
    t_e1 = time.time()
    clk2 = clocks2.stop(t_e0, t_e1)
    e2e_ms = float(np.median(e2e_runs))
    e2e_value = pairs_per_step_rank * world * args.steps / (e2e_ms * 1e-3)
    h2d = args.batch * D * 8 + args.batch * 8
    d2h = args.batch

    # ---- the same work through the drop-in Python surface: KDEContainer.construct_kde (fit, reported separately) and
    #      KDEContainer.score_samples on a host DataFrame for both densities, accept test on the host as the reference
    #      does it (emulator.py:737-746) ----
    plugin = None
    if rank == 0 and world == 1 and not args.no_plugin_leg:
        plugin = plugin_leg(args, fieldp, q_host, u_host, log_m, pairs_per_step_rank)

    # ---- executed pairs of one score launch (diagnostic counters on, outside the timed region) and the same launch
    #      with culling off (every pair evaluated: the FMA-pipe efficiency of the arithmetic) ----
    roof = None
    if rank == 0:
        st = (C.c_uint64 * 4)()
        q = draw_candidates(args.warmup, cand[0])
        u = u_pool[:args.batch]
        with _cabi.option("stats", 1):
            lib.skb_debug_counters(local, st, 1)
            sampler.score_accept(q, u)
            lib.skb_debug_counters(local, st, 1)
        exec_pairs = st[1] * 32.0
        culled_frac = (st[2] / float(st[3])) if st[3] else None
        with _cabi.option("cull", 0):
            sampler.score_accept(q, u)
            torch.cuda.synchronize()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            for i in range(2):
                sampler.score_accept(draw_candidates(args.warmup + 1 + i, cand[i & 1]), u)
            d1.record()
            torch.cuda.synchronize()
            dense_ms = d0.elapsed_time(d1) / 2.0
        fma, ex2 = C.c_double(), C.c_double()
        _cabi.check(lib.skb_measure_pipes(local, 300.0, C.byref(fma), C.byref(ex2)))
        peak_pairs = min(fma.value / (2 * D + 1), ex2.value)
        flop_pair = 3 * D + 1
        ach_exec = exec_pairs / (kernel_ms * 1e-3)
        ach_alg = pairs_per_step_rank / (kernel_ms * 1e-3)
        traffic = None
        tf = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tf):
            try:
                tj = json.load(open(tf))
                if tj.get("batch") == args.batch and tj.get("ntrain") == args.ntrain:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {
            # bound: the FP32 FMA pipe (with the SFU as second limit) — not HBM, not tensor cores
            "bound": "fp32", "kernel": "skb::skb_score_pq_kernel<8, false> (both KDEs of the group in one launch)",
            "achieved": ach_exec * flop_pair / 1e12, "peak": peak_pairs * flop_pair / 1e12, "unit": "TFLOP/s",
            "frac": ach_exec / peak_pairs, "traffic": traffic,
            "definition": "achieved = pairs the launch actually EVALUATES (counted by the kernel with diagnostics on, same "
                          "batch, outside the timed region) x (3d+1) FP32 FLOP / CUDA-event launch time inside the timed "
                          "region; peak = min(FFMA lane-ops/s / (2d+1), EX2/s) x (3d+1), the pair roofline of SURVEY.md §8d. "
                          "frac_algorithmic counts EVERY (query, training) pair of the launch (the pairs/s metric): it exceeds 1 "
                          "because tiles whose box test proves that every term flushes to zero are skipped; frac_dense is the "
                          "same launch with culling off",
            "executed_pair_fraction": exec_pairs / pairs_per_step_rank,
            "frac_algorithmic": ach_alg / peak_pairs,
            "frac_dense": (pairs_per_step_rank / (dense_ms * 1e-3)) / peak_pairs, "dense_ms_per_launch": dense_ms,
            "achieved_gpairs_executed": ach_exec / 1e9, "achieved_gpairs_algorithmic": ach_alg / 1e9, "peak_gpairs": peak_pairs / 1e9,
            "fp32_fma_peak_tflops_measured": 2.0 * fma.value / 1e12, "ex2_peak_per_s_measured": ex2.value,
            "peak_source": "measured live on this GPU by skb_measure_pipes (FFMA2 and MUFU.EX2 saturation loops, 300 ms each); "
                           "MEASURED_PEAKS.json holds only HBM and bf16-tensor figures",
            "kernel_ms_per_launch": kernel_ms, "pairs_per_launch": pairs_per_step_rank,
            "hbm_gbs_peak_for_reference": _hbm_peak(), "tiles_culled_frac": culled_frac,
        }

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "dtype_note": "fp32 pair terms, fp64 / fixed-point log-sum-exp state",
            "data": "synthetic",
            "config": workload_config(args, world),
            "step": "Philox draw of the batch from the field KDE (+10 % tail inflation) -> fused whitening + scoring of both KDEs + "
                    "ratio + accept -> order-preserving compaction -> (N > 1) NCCL all-gather of the accepted rows on a side stream",
            "accepted_samples_per_s": accepted / (ms_total * 1e-3),
            "accept_fraction": accepted / float(args.batch * world * args.steps), "log_m": log_m,
            "pairs_per_candidate": sampler.pairs_per_row, "gathered_rows_last_step": gathered_last,
            "fit_ms": {"wall_per_kde": [round(x, 1) for x in fit_wall_ms], "device_per_kde": [round(x, 2) for x in fit_dev_ms],
                       "what": "skb_create on 1e6 x 8 whitened host rows: upload, fp64 whitening, device kd ordering, tile packing"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "skb_ratio_accept through the C ABI with HOST buffers (skysampler_b200.pipeline.RatioSampler."
                           "score_accept_host: pinned candidate rows + uniforms in, accept flags out, copies staged by the library)",
                    "ms_per_step_each_pass": [round(r / args.steps, 2) for r in e2e_runs], "value_is": "median of the 3 timed passes",
                    "accepted_last_step": e2e_accepted, "plugin": plugin, "clocks": clk2},
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


def plugin_leg(args, fieldp, q_host, u_host, log_m, pairs_per_step):
    """KDEContainer (the reference's class surface) on host DataFrames: fit once, then per step two score_samples calls
    and the reference's accept arithmetic on the host."""
    import pandas as pd
    import torch
    from skysampler_b200 import KDEContainer
    tr = make_training(args.ntrain)
    cols = ["F%d" % i for i in range(D)]
    conts, fit_s = {}, {}
    for name in ("cluster", "field"):
        t = tr[name]
        c = KDEContainer(pd.DataFrame(t["x"], columns=cols), pd.Series(t["w"]), seed=1)
        c.set_whitening(t["mean1"], t["pmean"], t["comps"], t["std2"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c.construct_kde(H)
        fit_s[name] = time.perf_counter() - t0
        conts[name] = c
    del tr
    steps = min(args.steps, 5)
    frames = [pd.DataFrame(q_host[b % q_host.shape[0]].numpy(), columns=cols) for b in range(2)]
    m_factor = float(np.exp(log_m))

    def step(b):
        lc, jc = conts["cluster"].score_samples(frames[b % 2])
        lf, jf = conts["field"].score_samples(frames[b % 2])
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            inds = u_host[b % u_host.shape[0]].numpy() < np.exp(lc) * abs(jc) / (m_factor * np.exp(lf) * abs(jf))
        return int(inds.sum())

    step(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    acc = 0
    for b in range(steps):
        acc = step(b)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    for c in conts.values():
        c.drop_kde()
    return {"api": "skysampler_b200.KDEContainer.score_samples(DataFrame) x 2 densities + host accept (emulator.py:737-746 arithmetic), "
                   "pageable host rows", "ms_per_step": ms, "value": pairs_per_step / (ms * 1e-3), "unit": UNIT, "steps": steps,
            "accepted_last_step": acc,
            "construct_kde_s": {k: round(v, 3) for k, v in fit_s.items()}, "timing": "host wall clock around synchronous calls"}


def run_train_sharded(args):
    """SURVEY §8e item 3 / BASELINE config 5 top end: ONE weighted 8-D KDE whose training rows are sharded over the ranks;
    every rank scores the whole query batch against its shard (skb_score_partial), the per-query partial log-sum-exp
    states are combined with all_reduce(MAX) + all_reduce(SUM) over NVLink, every rank ends with the full log-densities."""
    import torch
    import torch.distributed as dist
    from skysampler_b200 import GaussianKDE, _cabi, synth
    from skysampler_b200 import dist as skd
    world, rank, local, dev = _dist_setup()
    lib = _cabi.lib()
    N = args.ntrain
    z = w = qh = None
    if rank == 0:
        x = synth.mixture_big(D, N, 500)
        w = synth.wide_weights(N, 501)
        m1, pm, comps, std2 = synth.whiten_params(x[::5], w[::5])
        z = np.ascontiguousarray((x - m1) @ comps.T / std2)
        del x
        qh = synth.queries(z, H, args.batch * 2, 502)
    z = skd.broadcast_array(z, 0, dev); w = skd.broadcast_array(w, 0, dev); qh = skd.broadcast_array(qh, 0, dev)
    a, b = skd.shard_bounds(N, world, rank)
    sum_w = float(w.sum())
    kde = GaussianKDE(H, device=local).fit(z[a:b], sample_weight=w[a:b])
    fit_ms = kde.fit_ms_
    del z
    q = torch.from_numpy(qh).to(dev)
    batches = [q[:args.batch], q[args.batch:]]

    def one_step(i):
        m, s = kde.score_partial(batches[i & 1])
        m, s = skd.combine_partial_lse(m, s)
        return skd.finish_logp(m, s, D, H, sum_w)

    for i in range(args.warmup):
        lp = one_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    clocks.wait_ready()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.skb_launch_count()
    t0 = time.time()
    e0.record()
    for i in range(args.steps):
        lp = one_step(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    ms_total = skd.timed_max_ms(e0.elapsed_time(e1), device=dev)
    clk = clocks.stop(t0, t1)
    # every rank must hold the same bits after the combine
    ref = lp.clone()
    if world > 1:
        dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, lp))
    same_all = skd.allreduce_sum_scalar(1.0 if same else 0.0, device=dev) == world
    value = float(args.batch) * N * args.steps / (ms_total * 1e-3)
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "c5_train_sharded: ONE weighted 8-D KDE, %d training rows sharded over %d ranks (%d per rank), "
                                       "%d whitened query rows per step scored by every rank, partial log-sum-exp combined with "
                                       "all_reduce(MAX) + all_reduce(SUM)" % (N, world, b - a, args.batch),
                           "n_train_total": N, "d": D, "bandwidth": H, "batch_rows": args.batch, "n_gpus": world,
                           "l2_policy": "two query batches alternate; training tiles per rank %d MB" % ((b - a) * 40 // 1000000)},
                "fit_ms_device_rank0": fit_ms, "identical_on_all_ranks": bool(same_all),
                "logp_finite_frac": float(torch.isfinite(lp).double().mean().item()),
                "collective_bytes_per_step": 2 * 8 * args.batch, "clocks": clk,
                "gpu_launches": int(lib.skb_launch_count() - launches0),
                "e2e": None, "roofline": None}
    kde.close()
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
    ap.add_argument("--no-plugin-leg", action="store_true")
    ap.add_argument("--shard", default="query", choices=["query", "train"],
                    help="query: rows sharded, training sets replicated (default, the headline); train: ONE KDE's training rows "
                         "sharded over the ranks, partial log-sum-exp states combined with two all-reduces (SURVEY §8e item 3)")
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
    # stdout carries exactly ONE JSON line: anything a library prints there at C level (NCCL's version banner, its
    # NCCL_DEBUG log when no NCCL_DEBUG_FILE is honoured) is sent to stderr instead; NCCL_DEBUG itself is left alone
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    cpu = None
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_subprocess(args)          # before CUDA is initialised in this process
    line = run_train_sharded(args) if args.shard == "train" else run_ours(args)
    if rank == 0:
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=real_stdout, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())

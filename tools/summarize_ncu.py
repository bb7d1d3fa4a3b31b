"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

usage: python tools/summarize_ncu.py <round tag> <launches.csv> <full .ncu-rep> <name>
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
os.makedirs(PROF, exist_ok=True)
tag, launches, rep, name = sys.argv[1:5]
cmdline = sys.argv[5] if len(sys.argv) > 5 else ""

if launches != "-":
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(PROF, "%s_launches_%s.csv" % (tag, name)), "w") as out:
        out.write("# ncu --metrics gpu__time_duration.sum --clock-control none  %s\n" % cmdline)
        out.write("# per-launch device time, cold-cache + serialised: compare SHARES, not absolutes\nid,kernel,grid,block,duration_ms\n")
        agg = collections.OrderedDict(); tot = 0
        for r in rows[1:]:
            if r[ix['Metric Name']] != 'gpu__time_duration.sum':
                continue
            v = float(r[ix['Metric Value']].replace(',', '')); u = r[ix['Metric Unit']]
            ms = v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}[u]
            k = r[ix['Kernel Name']]
            out.write("%s,\"%s\",\"%s\",\"%s\",%.6f\n" % (r[ix['ID']], k[:90], r[ix['Grid Size']], r[ix['Block Size']], ms))
            a = agg.setdefault(k[:90], [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
        out.write("# ---- shares ----\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write("# %-92s n=%4d %12.3f ms %6.2f%%\n" % (k, a[0], a[1], 100 * a[1] / tot))

if rep != "-":
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keep = ["gpu__time_duration.sum", "launch__", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_",
            "sm__pipe_", "smsp__issue_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
            "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "sm__throughput", "lts__t_bytes.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "gpu__dram_throughput", "smsp__average_warp_latency",
            "sm__sass_thread_inst_executed_op_f", "smsp__cycles_active.avg"]
    with open(os.path.join(PROF, "%s_%s_ncu_full.txt" % (tag, name)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on  %s\n" % cmdline)
        f.write("# kernel: %s\n" % vals[hdr.index("Kernel Name")])
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k) for k in keep):
                f.write("%s\t%s\t%s\n" % (h, u, v))
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    sc = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    rd = float(d["dram__bytes_read.sum"]) * sc[u["dram__bytes_read.sum"]]
    wr = float(d["dram__bytes_write.sum"]) * sc[u["dram__bytes_write.sum"]]
    print("dram read %.1f MB write %.1f MB  duration %s %s" % (rd / 1e6, wr / 1e6, d["gpu__time_duration.sum"], u["gpu__time_duration.sum"]))
    if name == "score_d8_bench":
        json.dump({"batch": 1 << 20, "ntrain": 1000000, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                   "source": "profiles/%s_%s_ncu_full.txt (ncu --set full, one launch of skb_score_pq_kernel<8>)" % (tag, name),
                   "algorithmic_bytes_per_launch": (1 << 20) * (8 * 8 + 8 + 1 + 1) + 2 * 15625 * 10 * 64 * 4},
                  open(os.path.join(PROF, "roofline_traffic.json"), "w"), indent=1)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]; data = rows[2:]; ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    tot = sum(f(r, "# Samples") for r in data)
    cats = ["stall_math", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_not_selected", "stall_selected",
            "stall_dispatch", "stall_branch_resolving", "stall_mio", "stall_barrier", "stall_no_inst"]
    with open(os.path.join(PROF, "%s_%s_stalls.txt" % (tag, name)), "w") as o:
        o.write("# warp-state samples by reason, whole kernel (ncu source page)\n")
        for c in cats:
            o.write("%-24s %6.2f%%\n" % (c, 100 * sum(f(r, c) for r in data) / tot))
        o.write("# top instructions by samples\n")
        for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:25]:
            o.write("%6.2f%%  exec=%12d  %s\n" % (100 * f(r, "# Samples") / tot, f(r, "Instructions Executed"), r[ix["Source"]][:100]))
    print(open(os.path.join(PROF, "%s_%s_stalls.txt" % (tag, name))).read()[:1500])

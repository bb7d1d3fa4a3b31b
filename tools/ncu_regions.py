"""Executed-instruction / stall-sample share per contiguous SASS region of a kernel in an .ncu-rep.

usage: python tools/ncu_regions.py report.ncu-rep [min share %]"""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) / 100 if len(sys.argv) > 2 else 0.004
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(src))
print(rows[0][:2])
hdr = rows[1]; isrc = hdr.index("Source"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
data = [(r[isrc], int(r[ie] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) > ie]
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print("total instr %.3e samples %d" % (tot, ts))
seg = []; cur = None
for i, (s, e, sm) in enumerate(data):
    if cur is None or not (0.7 * cur[2] <= e <= 1.4 * cur[2]):
        if cur: seg.append(cur)
        cur = [i, i, e, e, sm, s]
    else:
        cur[1] = i; cur[3] += e; cur[4] += sm
if cur: seg.append(cur)
for s in seg:
    if s[3] > thr * tot or s[4] > thr * ts:
        print("instr %5d-%5d  n=%4d  exec/instr~%11d  instr=%5.1f%%  samples=%5.1f%%   first: %s" % (
            s[0], s[1], s[1] - s[0] + 1, s[2], 100 * s[3] / tot, 100 * s[4] / ts, s[5][:50]))
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = raw[0]; v = raw[2]
for k in ("gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"):
    if k in h: print(k, v[h.index(k)], raw[1][h.index(k)])

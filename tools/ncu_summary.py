#!/usr/bin/env python
"""Summary of one kernel launch out of an `ncu --set full --import-source on` report:
headline metrics, stall reasons, and the source lines that execute the most instructions.
usage: python tools/ncu_summary.py <report.ncu-rep> [units_per_launch] [top_n]
(units_per_launch: e.g. seed hits of the launch, to print thread-instructions per unit)"""
import csv
import collections
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, un, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(un, vals)))
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    if w in d:
        print(f"{w:86s}{d[w][0]:18s}{d[w][1]}")
print("warp stall reasons (cycles per issued instruction, per warp):")
for k, v in d.items():
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
        name = k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
        if float(v[1]) >= 0.05:
            print(f"  {name:40s} {float(v[1]):.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
per = []
for bi, i in enumerate(hi):
    f = rows[i - 2][1].split("/")[-1]
    h = rows[i]
    ci, ct, cw = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("L1 Wavefronts Shared")
    end = hi[bi + 1] - 2 if bi + 1 < len(hi) else len(rows)
    for r in rows[i + 1:end]:
        if len(r) <= ct or r[0] == "":
            continue
        try:
            ln, ie, te, ws = int(r[0]), float(r[ci]), float(r[ct]), float(r[cw] or 0)
        except ValueError:
            continue
        if ie > 0:
            per.append((ie, te, ws, f, ln, r[1].strip()[:100]))
tot = sum(p[0] for p in per) or 1.0
tt = sum(p[1] for p in per)
print(f"source lines by executed warp instructions (total {tot:.4g} warp, {tt:.4g} thread"
      + (f", {tt / units:.1f} thread-instructions per unit" if units else "") + "):")
per.sort(reverse=True)
for p in per[:top_n]:
    u = f"{p[1] / units:6.1f}/unit " if units else ""
    print(f"{p[0] / tot * 100:5.1f}% {u}shared wavefronts {p[2] / 1e9:5.2f}G  {p[3]}:{p[4]}  {p[5]}")
byfile = collections.Counter()
for p in per:
    byfile[p[3]] += p[1]
if units:
    print("thread-instructions per unit by file:", {k: round(v / units, 1) for k, v in byfile.items()})

#!/usr/bin/env python
"""Per-launch DRAM traffic of a kernel out of an `ncu --set full` report, for bench.py's `roofline.traffic`.
usage: python tools/ncu_traffic.py <report.ncu-rep> <kernel-name-substring> <key> [note]
Merges {key: {...}} into profiles/ncu_traffic.json (the longest matching launch of the report)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pat, key = sys.argv[1:4]
note = sys.argv[4] if len(sys.argv) > 4 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
head, units, body = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(head)}


def val(row, name):
    v, u = float(row[col[name]].replace(",", "")), units[col[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}
    return v * scale.get(u, 1)


cand = [r for r in body if pat in r[col["Kernel Name"]]]
assert cand, "no launch of %s in %s" % (pat, rep)
r = max(cand, key=lambda x: val(x, "gpu__time_duration.sum"))
entry = {"kernel": r[col["Kernel Name"]], "dram_bytes_read": val(r, "dram__bytes_read.sum"),
         "dram_bytes_written": val(r, "dram__bytes_write.sum"),
         "duration_ms_under_ncu": val(r, "gpu__time_duration.sum"), "grid": r[col["launch__grid_size"]],
         "report": os.path.basename(rep), "note": note}
entry["dram_bytes_per_launch"] = entry["dram_bytes_read"] + entry["dram_bytes_written"]
# executed instructions: warp-level count x average active threads = thread-instructions (the "issue" roofline)
try:
    wi = val(r, "smsp__inst_executed.sum")
    tpi = val(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
    entry.update({"warp_inst_executed": wi, "threads_per_inst": tpi, "thread_inst_executed": wi * tpi,
                  "thread_inst_per_s_under_ncu": wi * tpi / (entry["duration_ms_under_ncu"] * 1e-3),
                  "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "alu_pipe_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                  "l1_data_pipe_pct": val(r, "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed")})
except Exception as e:  # noqa: BLE001
    entry["inst_note"] = str(e)
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
cur = json.load(open(path)) if os.path.exists(path) else {}
cur[key] = entry
json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(entry))

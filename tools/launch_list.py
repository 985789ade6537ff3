#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_list.py <launches.csv> [comment ...]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
for r in rows:
    if r is hdr or r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ik])
    tot[name] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    cnt[name] += 1
all_ms = sum(tot.values())
for c in sys.argv[2:]:
    print("# " + c)
print(f"# {sum(cnt.values())} launches, {all_ms:.1f} ms in kernels")
for k, v in tot.most_common():
    print(f"{v / all_ms * 100:6.2f} %  {v:10.2f} ms  {cnt[k]:5d} launches  {k}")

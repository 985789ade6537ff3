#!/usr/bin/env python
"""Instruction mix of a SASS address range (e.g. the step loop of nwp_kernel) out of `cuobjdump -sass`.
usage: python tools/sass_mix.py <lib.so> <mangled kernel name> <start hex> <end hex> [cells]
Prints the opcode histogram, the split over the integer pipes as measured by tools/int_peak.cu on sm_100
(ALU pipe: compare / select / min-max / logic / shifts / VIADD... see the table below; FMA pipe: IMAD*, moves
implemented as IMAD.MOV, IADD3 when issued there) and the listing itself."""
import collections
import re
import subprocess
import sys

lib, fun, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
cells = int(sys.argv[5]) if len(sys.argv) > 5 else 0
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
ins = []
for line in txt.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        a = int(m.group(1), 16)
        if lo <= a < hi:
            ins.append((a, m.group(2).strip()))
ops = collections.Counter()
for a, t in ins:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    ops[t.split()[0].split(".")[0]] += 1
ALU = {"LOP3", "ISETP", "SEL", "VIMNMX3", "VIMNMX", "SHF", "PRMT", "LEA", "VIADDMNMX", "ISETP", "PLOP3", "FSEL", "IMNMX", "VABSDIFF", "POPC", "FLO", "BREV"}
FMA = {"IMAD", "VIADD", "IADD3", "IADD", "MOV"}
n = len(ins)
alu = sum(v for k, v in ops.items() if k in ALU)
fma = sum(v for k, v in ops.items() if k in FMA)
print(f"# {fun}  [{lo:#x}, {hi:#x}): {n} instructions" + (f" = {n / cells:.2f} per cell ({cells} cells)" if cells else ""))
print(f"# ALU-pipe class {alu}" + (f" ({alu / cells:.2f}/cell)" if cells else "") + f", FMA-pipe class {fma}" + (f" ({fma / cells:.2f}/cell)" if cells else "")
      + f", other {n - alu - fma} (shared loads, shuffles, branches, barriers)")
for k, v in ops.most_common():
    print(f"#   {k:12s} {v:5d}" + (f"  {v / cells:.2f}/cell" if cells else ""))
for a, t in ins:
    print(f"/*{a:04x}*/  {t}")

#!/usr/bin/env python
"""Small, short-running driver for ncu: a down-scaled cfg2 workload through imsame_gpu_align,
`--reps` times (first = warm-up).  usage: python tools/profile_step.py --scale 0.02 --reps 2"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from imsame_b200 import api, hostlib as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.02)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--L", type=int, default=250)
ap.add_argument("--nw-mode", type=int, default=0)
a = ap.parse_args()
nd, nq, g = int(10_000_000 * a.scale), int(1_000_000 * a.scale), max(2, int(1000 * a.scale))
pool = H.SynthPool(2001, g, 1_000_000)
db = pool.db_reads(0, nd, a.L)
q = pool.query_reads(0, nq, a.L, 0.03)
pool.close()
ds = np.arange(nd + 1, dtype=np.uint64) * a.L
qs = np.arange(nq + 1, dtype=np.uint64) * a.L
ctx = api.Imsame(0)
ctx.set_nw_mode(a.nw_mode)
for _ in range(a.reps):
    out, st = ctx.align((db, ds), (q, qs), api.make_params(n_threads=4))
print(json.dumps({k: st[k] for k in ("n_hits", "n_evalue_pass", "n_pairs", "n_pairs_dp", "n_cells", "ms_k1", "ms_k2",
                                      "ms_k2b", "ms_k3", "ms_total", "total_launches")}
                 | {"accepted": int(out["accepted"].sum()),
                    "gcups": st["n_cells"] / max(st["ms_k3"], 1e-9) / 1e6}))
ctx.close()

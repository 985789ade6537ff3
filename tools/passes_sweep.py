#!/usr/bin/env python
"""cfg2 at full size, device-resident: one scan pass against "early words first" with different numbers of early
bands (IMSAME_EARLY_BANDS), the run's phase times and seed-hit counts.  usage: python tools/passes_sweep.py [--scale S]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from imsame_b200 import api, hostlib as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--bands", default="0,2,3,4,6,8,12")
a = ap.parse_args()
L = 250
nd, nq, g = int(10_000_000 * a.scale), int(1_000_000 * a.scale), max(2, int(1000 * a.scale))
pool = H.SynthPool(2001, g, 1_000_000)
db = pool.db_reads(0, nd, L)
q = pool.query_reads(0, nq, L, 0.03)
pool.close()
ds = np.arange(nd + 1, dtype=np.uint64) * L
qs = np.arange(nq + 1, dtype=np.uint64) * L
ctx = api.Imsame(0)
p = api.make_params(n_threads=4)
ctx.set_query((q, qs), p)
ctx.set_db((db, ds))
ref = None
for nb in [int(v) for v in a.bands.split(",")]:
    if nb == 0:
        ctx.set_passes(1)
    else:
        ctx.set_passes(2)
        os.environ["IMSAME_EARLY_BANDS"] = str(nb)
    for _ in range(2):
        st = ctx.run(p)
    rec = ctx.fetch()
    key = (rec["accepted"].copy(), rec["db_seq"].copy(), rec["qpos_end"].copy(), rec["length"].copy(), rec["identities"].copy())
    same = True if ref is None else all(np.array_equal(x, y) for x, y in zip(ref, key))
    ref = ref or key
    print(json.dumps({"early_bands": nb, "passes": st["scan_passes"], "ms_total": round(st["ms_total"], 1),
                      "ms_k1": round(st["ms_k1"], 1), "ms_k2": round(st["ms_k2"], 1), "ms_k2b": round(st["ms_k2b"], 1),
                      "ms_k3": round(st["ms_k3"], 1), "n_hits": st["n_hits"], "n_pairs": st["n_pairs"],
                      "n_pairs_dp": st["n_pairs_dp"], "accepted": int(rec["accepted"].sum()), "same_records": same}), flush=True)
ctx.close()

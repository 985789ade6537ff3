#!/usr/bin/env python
"""diagnostic: distribution of the accepting k-mer position and DP counts (scaled cfg2)"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from imsame_b200 import api, hostlib as H
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
L = 250
nd, nq, g = int(10_000_000 * scale), int(1_000_000 * scale), max(2, int(1000 * scale))
pool = H.SynthPool(2001, g, 1_000_000)
db = pool.db_reads(0, nd, L); q = pool.query_reads(0, nq, L, 0.03); pool.close()
ds = np.arange(nd + 1, dtype=np.uint64) * L; qs = np.arange(nq + 1, dtype=np.uint64) * L
ctx = api.Imsame(0)
out, st = ctx.align((db, ds), (q, qs), api.make_params(n_threads=4))
acc = out["accepted"] == 1
erel = (out["qpos_end"][acc] - qs[:-1][acc] + 1).astype(np.int64)
print(json.dumps({k: st[k] for k in ("n_hits", "n_evalue_pass", "n_pairs", "n_pairs_dp", "ms_k2", "ms_k3")}))
print("accepted", int(acc.sum()), "of", nq)
h, _ = np.histogram(erel, bins=[0, 12, 16, 24, 32, 48, 64, 96, 128, 192, 260])
print("e_rel hist [0,12,16,24,32,48,64,96,128,192,260):", h.tolist())
ident = out["identities"][acc] / np.maximum(out["length"][acc], 1)
print("identity hist (0.5-0.6,0.6-0.7,0.7-0.8,0.8-0.9,0.9-1.0]:", np.histogram(ident, bins=[0.5, 0.6, 0.7, 0.8, 0.9, 1.01])[0].tolist())

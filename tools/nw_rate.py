#!/usr/bin/env python
"""Cell rate of the NW kernels by read length, through imsame_gpu_nw_batch (explicit pairs; the kernel time is the
CUDA-event time the library records around its launches).  Random pairs, default gap scores; packed-word kernel
(where pw_eligible admits the pair) and generic kernel side by side, results compared.
usage: python tools/nw_rate.py [--pairs N] [--lens 150,250,300,...] [--out file.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from imsame_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=150000)
ap.add_argument("--lens", default="150,250,257,300,308,321,400,600")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--out", default="")
args = ap.parse_args()

ctx = api.Imsame(0)
rng = np.random.default_rng(1)
B = np.frombuffer(b"ACGT", dtype=np.uint8)
rows = []
for L in [int(v) for v in args.lens.split(",")]:
    n = max(1000, int(args.pairs * (250.0 / L) ** 2))
    X = B[rng.integers(0, 4, size=(n, L))]
    Y = B[rng.integers(0, 4, size=(n, L))]
    # one pair in 50 is a noisy copy (the accepted pairs of a run; wide reads: the second run)
    for i in range(0, n, 50):
        Y[i] = X[i]
        hit = rng.random(L) < 0.03
        Y[i, hit] = B[rng.integers(0, 4, size=int(hit.sum()))]
    xs = [np.ascontiguousarray(X[i]) for i in range(n)]
    ys = [np.ascontiguousarray(Y[i]) for i in range(n)]
    cells = float(n) * (L - 1) * (L - 1)
    row = {"read_len": L, "pairs": n}
    res = {}
    for mode, name in ((0, "packed"), (1, "generic")):
        ctx.set_nw_mode(mode)
        best = None
        for _ in range(args.reps):
            got, ms = ctx.nw_batch(xs, ys)
            best = ms if best is None or ms < best else best
        res[name] = np.asarray(got)
        row[name + "_ms"] = round(best, 3)
        row[name + "_gcups"] = round(cells / best / 1e6, 1)
    ctx.set_nw_mode(0)
    row["identical"] = bool(np.array_equal(res["packed"], res["generic"]))
    rows.append(row)
    print(json.dumps(row), flush=True)
ctx.close()
if args.out:
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)

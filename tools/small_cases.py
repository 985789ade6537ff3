#!/usr/bin/env python
"""Small end-to-end cases (each checked against the oracle; also what one would run under a memory checker -- compute-sanitizer is closed on the GPU pool): fixed-length, ragged reads with
word breaks, a run-time seed length, device-resident samples with the device reverse complement, the explicit-pair
NW batch and the winners' traceback.  Every case is also checked against the oracle.
usage: python tools/small_cases.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import helpers as hp  # noqa: E402
import synth_cases as sc  # noqa: E402
from imsame_b200 import api  # noqa: E402


def recs(out):
    return {int(r): (int(o["db_seq"]), int(o["qpos_end"]), int(o["db_pos"]), int(o["length"]), int(o["identities"]))
            for r, o in enumerate(out) if o["accepted"]}


def oracle(db, ds, q, qs, nt, brk=None, **kw):
    best, _ = hp.oracle_align(hp.OracleSeqs(seq=db, start=ds, brk=brk), hp.OracleSeqs(seq=q, start=qs), hp.default_params(n_threads=nt, **kw))
    return hp.best_to_records(best, len(qs) - 1)


ctx = api.Imsame(0)
db, ds, q, qs = sc.fixed_case(11, 2, 20000, 150, 1500, 200, 0.04)
out, st = ctx.align((db, ds), (q, qs), api.make_params(n_threads=4))
assert recs(out) == oracle(db, ds, q, qs, 4) and len(recs(out)) > 50
ops_off, ops, cell = ctx.traceback((db, ds), (q, qs), out)
assert int(ops_off[-1]) == len(ops) > 50
rdb, rds, rq, rqs = sc.ragged_case(12, 2, 20000, 1200, 150, 0.05, lo=20, hi=320)
rng = np.random.default_rng(3)
brk = np.unique(rng.integers(1, len(rdb), size=200)).astype(np.uint64)
brk = np.array([b for b in brk if b not in set(rds.tolist())], dtype=np.uint64)
out, _ = ctx.align((rdb, rds), (rq, rqs), api.make_params(n_threads=3), db_breaks=brk)
assert recs(out) == oracle(rdb, rds, rq, rqs, 3, brk=brk) and len(recs(out)) > 20
ops_off, ops, cell = ctx.traceback((rdb, rds), (rq, rqs), out, api.make_params(n_threads=3), db_breaks=brk)
ctx.set_kmer(9)
out, _ = ctx.align((db, ds), (q, qs), api.make_params(n_threads=2))
assert recs(out) == oracle(db, ds, q, qs, 2, k=9)
ctx.set_kmer(12)
S_db, S_q = ctx.sample((rdb, rds), brk), ctx.sample((rq, rqs))
S_r = ctx.sample_revcomp(S_db)
out, _ = ctx.align_samples(S_db, S_q, api.make_params(n_threads=3))
assert recs(out) == oracle(rdb, rds, rq, rqs, 3, brk=brk)
out_r, _ = ctx.align_samples(S_r, S_q, api.make_params(n_threads=3))
for s_ in (S_db, S_q, S_r):
    s_.free()
xs, ys = sc.random_pairs(4, 120, max_len=280, long_every=40)
got, _ = ctx.nw_batch(xs, ys)
ctx.close()
print("small_cases: all cases equal the oracle")

#!/usr/bin/env python
"""Time-bounded seeded fuzz of the whole GPU path (imsame_gpu_align through the C ABI) against the oracle's
scan-order form on small edge-case inputs: reads of 1..450 bases (shorter than a seed, one word long, the
packed-word classes, the wide classes and the generic kernel), word breaks in the database, more threads than
reads, random thresholds (e-value up to 0.5, coverage 0.01..1, identity 0.3..1) and gap scores (incl. 0), one or
two scan passes.  The same generator family pins the oracle to the compiled reference on the CPU
(tests/test_oracle_fuzz.py).
usage: python tools/gpu_fuzz.py [--seconds 12] [--seed 1] [--dry]      (--dry: oracle only, no GPU)"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import helpers as hp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=12.0)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--dry", action="store_true")
a = ap.parse_args()
B = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_case(rng):
    lo, hi = [(1, 30), (11, 14), (12, 120), (100, 260), (240, 322), (200, 450)][int(rng.integers(0, 6))]
    # longer reads get longer genomes: the number of overlapping (read, database read) pairs -- NW calls of the
    # oracle -- stays in the thousands
    G, Lg = int(rng.integers(1, 4)), int(rng.integers(300, 5000)) * (1 if hi < 130 else 10)
    genomes = [B[rng.integers(0, 4, size=Lg)] for _ in range(G)]
    div = float(rng.choice([0.0, 0.02, 0.05, 0.1, 0.2]))

    def read(dv, absent_p):
        L = int(rng.integers(lo, hi + 1))
        if rng.random() < absent_p:
            return B[rng.integers(0, 4, size=L)].copy()
        g = genomes[int(rng.integers(0, G))]
        L = min(L, len(g))
        at = int(rng.integers(0, len(g) - L + 1))
        r = g[at:at + L].copy()
        m = rng.random(L) < dv
        r[m] = B[rng.integers(0, 4, size=int(m.sum()))]
        if rng.random() < 0.3 and L > 30:
            k = int(rng.integers(5, L - 5))
            r = np.delete(r, slice(k, k + int(rng.integers(1, 4))))
        if rng.random() < 0.3 and len(r) > 30:
            k = int(rng.integers(5, len(r) - 5))
            r = np.insert(r, k, B[rng.integers(0, 4, size=int(rng.integers(1, 4)))])
        return r

    nd, nq = int(rng.integers(3, 600)), int(rng.integers(1, 120))
    dbr = [read(0.005, 0.05) for _ in range(nd)]
    qr = [read(div, 0.3) for _ in range(nq)]
    db, q = np.ascontiguousarray(np.concatenate(dbr)), np.ascontiguousarray(np.concatenate(qr))
    ds = np.concatenate([[0], np.cumsum([len(r) for r in dbr])]).astype(np.uint64)
    qs = np.concatenate([[0], np.cumsum([len(r) for r in qr])]).astype(np.uint64)
    brk = None
    if rng.random() < 0.5 and len(db) > 50:
        cand = np.unique(rng.integers(1, len(db), size=int(rng.integers(1, 40)))).astype(np.uint64)
        starts = set(int(x) for x in ds)
        brk = np.array([b for b in cand if int(b) not in starts], dtype=np.uint64)
        if len(brk) == 0:
            brk = None
    par = dict(n_threads=int(rng.choice([1, 2, 3, 4, 7, 8, 100])),
               evalue=[None, 1e-20, 1e-10, 1e-5, 1e-2, 0.5][int(rng.integers(0, 6))],
               coverage=float(rng.choice([0.5, 0.2, 0.8, 1.0, 0.01])), identity=float(rng.choice([0.5, 0.3, 0.7, 0.9, 1.0])),
               igap=int(rng.choice([5, 0, 1, 10, 30])), egap=int(rng.choice([2, 0, 1, 5])))
    passes = int(rng.choice([0, 1, 2, 2]))
    return db, ds, q, qs, brk, par, passes, (lo, hi, div, nd, nq)


def recs(out):
    return {int(r): (int(o["db_seq"]), int(o["qpos_end"]), int(o["db_pos"]), int(o["length"]), int(o["identities"]))
            for r, o in enumerate(out) if o["accepted"]}


def oracle_of(idx):
    """worker: the oracle's records of case idx (min-key form: one NW per distinct pair; == the scan-order form,
    tests/test_oracle_golden.py)"""
    hp.oracle_set_threads(1)
    db, ds, q, qs, brk, par, passes, shape = make_case(np.random.default_rng(a.seed * 1000003 + idx))
    best, _ = hp.oracle_align(hp.OracleSeqs(seq=db, start=ds, brk=brk), hp.OracleSeqs(seq=q, start=qs),
                              hp.default_params(n_threads=par["n_threads"], evalue=par["evalue"], coverage=par["coverage"],
                                                identity=par["identity"], igap=par["igap"], egap=par["egap"]), bulk=True)
    return idx, hp.best_to_records(best, len(qs) - 1)


def main():
    from concurrent.futures import ProcessPoolExecutor, as_completed
    hp.oracle()  # built / loaded before the workers fork
    workers = max(1, (os.cpu_count() or 2) - 1)
    t_end = time.time() + a.seconds
    n_cases = n_bad = n_rec = n_err = 0
    with ProcessPoolExecutor(max_workers=workers) as pool:
        futs = [pool.submit(oracle_of, i) for i in range(int(a.seconds * workers * 6) + workers)]  # ~0.2-0.3 s per case
        ctx = None
        if not a.dry:
            from imsame_b200 import api
            ctx = api.Imsame(0)  # the context comes up while the workers compute
        for f in as_completed(futs):
            if time.time() > t_end:
                break
            idx, want = f.result()
            db, ds, q, qs, brk, par, passes, shape = make_case(np.random.default_rng(a.seed * 1000003 + idx))
            desc = f"case {a.seed}/{idx}: reads {shape[0]}..{shape[1]} div {shape[2]} nd {shape[3]} nq {shape[4]} " \
                   f"breaks {0 if brk is None else len(brk)} passes {passes} {par}"
            n_cases += 1
            n_rec += len(want)
            if ctx is None:
                continue
            ctx.set_passes(passes)
            try:
                out, st = ctx.align((db, ds), (q, qs), api.make_params(min_e_value=par["evalue"], min_coverage=par["coverage"],
                                                                       min_identity=par["identity"], igap=par["igap"],
                                                                       egap=par["egap"], n_threads=par["n_threads"]), db_breaks=brk)
            except Exception as e:  # noqa: BLE001
                n_err += 1
                print("GPU ERROR", desc, repr(e), flush=True)
                continue
            got = recs(out)
            if got != want:
                n_bad += 1
                diff = sorted(set(got.items()) ^ set(want.items()))[:4]
                print("MISMATCH", desc, f"gpu {len(got)} oracle {len(want)} first differences {diff}", flush=True)
        for f in futs:
            f.cancel()
        if ctx is not None:
            ctx.close()
    print(f"gpu_fuzz seed {a.seed}: {n_cases} cases, {n_rec} oracle records, {n_bad} mismatches, {n_err} GPU errors"
          + (" (dry run: oracle only)" if a.dry else ""), flush=True)
    return 1 if (n_bad or n_err) else 0


if __name__ == "__main__":
    sys.exit(main())

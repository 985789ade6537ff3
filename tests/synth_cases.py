"""Seeded synthetic inputs shared by CPU and GPU tests."""
import numpy as np

from imsame_b200 import hostlib as H


def fixed_case(seed, n_genomes, genome_len, L, nd, nq, divergence):
    """(db_seq, db_start, q_seq, q_start) with fixed-length reads"""
    pool = H.SynthPool(seed, n_genomes, genome_len)
    db = pool.db_reads(0, nd, L)
    q = pool.query_reads(0, nq, L, divergence)
    pool.close()
    ds = np.arange(nd + 1, dtype=np.uint64) * L
    qs = np.arange(nq + 1, dtype=np.uint64) * L
    return db, ds, q, qs


def ragged_case(seed, n_genomes, genome_len, nd, nq, divergence, lo=30, hi=400):
    """variable read lengths: cut fixed-length reads to random lengths"""
    rng = np.random.default_rng(seed)
    db, ds, q, qs = fixed_case(seed, n_genomes, genome_len, hi, nd, nq, divergence)
    dl = rng.integers(lo, hi + 1, size=nd)
    ql = rng.integers(lo, hi + 1, size=nq)
    dseq = np.concatenate([db[i * hi:i * hi + dl[i]] for i in range(nd)])
    qseq = np.concatenate([q[i * hi:i * hi + ql[i]] for i in range(nq)])
    dstart = np.concatenate([[0], np.cumsum(dl)]).astype(np.uint64)
    qstart = np.concatenate([[0], np.cumsum(ql)]).astype(np.uint64)
    return dseq, dstart, qseq, qstart


def random_pairs(seed, n, max_len=300, long_every=0):
    """NW test pairs: noisy/shifted copies with indels, random pairs, tiny reads"""
    rng = np.random.default_rng(seed)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    xs, ys = [], []
    for it in range(n):
        xlen = int(rng.integers(12, max_len))
        ylen = int(rng.integers(11, max_len))
        if long_every and it % long_every == 0:
            xlen = int(rng.integers(300, 3001))
            ylen = int(rng.integers(300, 3001))
        if it % 13 == 0:
            xlen = int(rng.integers(2, 8))
        if it % 17 == 0:
            ylen = int(rng.integers(2, 8))
        if it % 11 == 0:
            xlen = ylen = 250
        x = B[rng.integers(0, 4, size=xlen)]
        mode = it % 4
        if mode == 0:
            y = B[rng.integers(0, 4, size=ylen)]
        else:
            pe = (0.03, 0.15, 0.30)[mode - 1]
            off = int(rng.integers(0, xlen)) - xlen // 3
            y = np.empty(ylen, dtype=np.uint8)
            src = off
            u = rng.random(ylen)
            rb = B[rng.integers(0, 4, size=ylen)]
            for j in range(ylen):
                if u[j] < pe / 6:
                    src += int(rng.integers(1, 5))
                if u[j] > 1 - pe / 6:
                    y[j] = rb[j]
                    continue
                c = x[src] if 0 <= src < xlen else rb[j]
                if 0.5 < u[j] < 0.5 + pe:
                    c = rb[j]
                y[j] = c
                src += 1
        xs.append(np.ascontiguousarray(x))
        ys.append(np.ascontiguousarray(y))
    return xs, ys


def _revcomp(b):
    return b.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1]


def write_allvsall_filter_samples(d):
    """four samples for the all-vs-all tests about revComp's text filter vs the loader's (tests/test_gpu_zz_allvsall_text_filters.py,
    tests/test_host_cpu.py): s0 forward strand and clean; s1-s3 reverse strand (so that the .r.align files hold records)"""
    from imsame_b200 import hostlib as H
    pool = H.SynthPool(4107, 3, 30000)
    L, n = 150, 400
    sets = [pool.db_reads(i * 1000, n, L) for i in range(4)]
    pool.close()
    reads = [[bytes(s[r * L:(r + 1) * L]) for r in range(n)] for s in sets]
    with open(d / "s0.fasta", "wb") as f:  # forward strand, clean
        for r, b in enumerate(reads[0]):
            f.write(b">a%d\n" % r + b + b"\n")
    with open(d / "s1.fasta", "wb") as f:  # reverse strand, multi-line CRLF: '\r' inside every record
        for r, b in enumerate(reads[1]):
            b = _revcomp(b)
            f.write(b">b%d\r\n" % r + b[:70] + b"\r\n" + b[70:] + b"\r\n")
    with open(d / "s2.fasta", "wb") as f:  # reverse strand, gap characters / 'U' / digits inside some records
        for r, b in enumerate(reads[2]):
            b = _revcomp(b)
            if r % 3 == 0:
                b = b[:60] + b"-" + b[60:]
            elif r % 3 == 1:
                b = b[:90] + b"U" + b[90:]
            elif r % 7 == 2:
                b = b[:30] + b"12 " + b[30:]
            f.write(b">c%d\n" % r + b + b"\n")
    with open(d / "s3.fasta", "wb") as f:  # reverse strand, clean apart from 'N' (a letter: kept by revComp): device path
        for r, b in enumerate(reads[3]):
            b = _revcomp(b)
            if r % 5 == 0:
                b = b[:100] + b"N" + b[100:]
            f.write(b">d%d\n" % r + b + b"\n")

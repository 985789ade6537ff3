"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.
Integer / index work: bit-exact."""
import ctypes as C

import numpy as np
import pytest

import helpers as hp
import synth_cases as sc

pytestmark = pytest.mark.gpu


def oracle_records(db, ds, q, qs, n_threads, breaks=None, **kw):
    odb = hp.OracleSeqs(seq=db, start=ds, brk=breaks)
    oq = hp.OracleSeqs(seq=q, start=qs)
    p = hp.default_params(n_threads=n_threads, **kw)
    best, st = hp.oracle_align(odb, oq, p)
    return hp.best_to_records(best, len(qs) - 1), st


def gpu_records(out):
    return {int(r): (int(o["db_seq"]), int(o["qpos_end"]), int(o["db_pos"]), int(o["length"]), int(o["identities"]))
            for r, o in enumerate(out) if o["accepted"]}


def test_nw_batch_matches_oracle(gpu):
    lib = hp.oracle()
    xs, ys = sc.random_pairs(5, 1500, max_len=300)
    x2, y2 = sc.random_pairs(6, 60, max_len=300, long_every=3)  # up to 3000 x 3000: multi-pass path
    xs += x2
    ys += y2
    for igap, egap in ((5, 2), (3, 1), (0, 0), (7, 3)):
        got, ms = gpu.nw_batch(xs, ys, igap=igap, egap=egap)
        for i, (x, y) in enumerate(zip(xs, ys)):
            s = C.c_int32(); bx = C.c_uint32(); by = C.c_uint32(); ln = C.c_uint32(); idn = C.c_uint32()
            u8p = C.POINTER(C.c_ubyte)
            lib.orc_nw_forward(x.ctypes.data_as(u8p), len(x), y.ctypes.data_as(u8p), len(y), -igap, -egap,
                               C.byref(s), C.byref(bx), C.byref(by), C.byref(ln), C.byref(idn))
            want = (s.value, bx.value, by.value, ln.value, idn.value)
            assert tuple(int(v) for v in got[i]) == want, (i, len(x), len(y), igap, egap)


def _oracle_nw(lib, x, y, igap, egap):
    s = C.c_int32(); bx = C.c_uint32(); by = C.c_uint32(); ln = C.c_uint32(); idn = C.c_uint32()
    u8p = C.POINTER(C.c_ubyte)
    lib.orc_nw_forward(x.ctypes.data_as(u8p), len(x), y.ctypes.data_as(u8p), len(y), -igap, -egap,
                       C.byref(s), C.byref(bx), C.byref(by), C.byref(ln), C.byref(idn))
    return (s.value, bx.value, by.value, ln.value, idn.value)


def test_nw_batch_packed_kernel_matches_oracle_and_generic(gpu):
    """short reads run through the packed-word kernel (nwp.cuh); it must agree bit for bit with the
    oracle and with the generic kernel, incl. the largest admitted sizes, low-complexity reads (ties)
    and the gap settings at the edge of pw_eligible"""
    lib = hp.oracle()
    rng = np.random.default_rng(77)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    xs, ys = sc.random_pairs(15, 1800, max_len=257)
    for xl, yl in ((256, 257), (250, 250), (256, 256), (2, 257), (256, 2), (2, 2), (255, 33), (33, 255), (200, 17),
                   (512, 200), (400, 257), (511, 150)):
        for rep in range(6):
            x = B[rng.integers(0, 4 if rep < 4 else 2, size=xl)]
            y = x[:yl].copy() if (rep % 2 and yl <= xl) else B[rng.integers(0, 4 if rep < 4 else 2, size=yl)]
            xs.append(np.ascontiguousarray(x)); ys.append(np.ascontiguousarray(y))
    for igap, egap in ((5, 2), (0, 0), (7, 3), (40, 3), (1, 0)):
        gpu.set_nw_mode(0)
        got, _ = gpu.nw_batch(xs, ys, igap=igap, egap=egap)
        gpu.set_nw_mode(1)
        gen, _ = gpu.nw_batch(xs, ys, igap=igap, egap=egap)
        gpu.set_nw_mode(0)
        assert np.array_equal(np.asarray(got), np.asarray(gen)), (igap, egap)
        for i in range(0, len(xs), 1 if (igap, egap) == (5, 2) else 7):
            if len(xs[i]) < 2 or len(ys[i]) < 2:
                continue
            assert tuple(int(v) for v in got[i]) == _oracle_nw(lib, xs[i], ys[i], igap, egap), (i, len(xs[i]), len(ys[i]), igap, egap)


def test_nw_batch_wide_reads_in_packed_words(gpu):
    """query reads of 257..321 bases (NW classes 9 and 10: 18 / 20 columns per lane): the identities no longer fit their
    8 bits, the statistics word is split by the path's geometry and near-complete overlaps (256 and more identities)
    are run a second time with the identities alone -- bit for bit the oracle's and the generic kernel's 5-tuples,
    incl. identical reads, the largest admitted sizes and database reads of up to 512 bases"""
    lib = hp.oracle()
    rng = np.random.default_rng(303)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    xs, ys = [], []
    sizes = [(300, 300), (301, 301), (309, 309), (257, 257), (256, 257), (290, 289), (289, 290), (250, 321), (321, 321),
             (512, 300), (400, 321), (512, 321), (2, 300), (300, 2 + 256), (150, 310), (310, 150)]
    for xl, yl in sizes:
        for rep in range(10):
            x = B[rng.integers(0, 4 if rep < 8 else 2, size=xl)]
            if rep % 5 == 0:  # a copy (identical up to the shorter length): every identity counts
                y = np.resize(x, yl).copy()
            elif rep % 5 == 1:  # a noisy, slightly shifted copy: 256 and more identities, a few gaps
                off = int(rng.integers(0, 12))
                y = np.resize(np.roll(x, -off), yl).copy()
                hit = rng.random(yl) < 0.02
                y[hit] = B[rng.integers(0, 4, size=int(hit.sum()))]
            elif rep % 5 == 2:  # short overlap at one corner
                y = B[rng.integers(0, 4, size=yl)]
                n = min(40, xl, yl)
                y[:n] = x[xl - n:]
            else:
                y = B[rng.integers(0, 4 if rep < 8 else 2, size=yl)]
            xs.append(np.ascontiguousarray(x)); ys.append(np.ascontiguousarray(y))
    x2, y2 = sc.random_pairs(21, 600, max_len=322)
    xs += x2; ys += y2
    # (the batch's longest reads, 512 x 321, give every pair of it a score offset: 451 with the default gap scores;
    # with (12, 4) the offset no longer fits and the two kernels share the batch pair by pair)
    for igap, egap in ((5, 2), (0, 0), (8, 1), (1, 0), (12, 4)):
        gpu.set_nw_mode(0)
        got, _ = gpu.nw_batch(xs, ys, igap=igap, egap=egap)
        gpu.set_nw_mode(1)
        gen, _ = gpu.nw_batch(xs, ys, igap=igap, egap=egap)
        gpu.set_nw_mode(0)
        assert np.array_equal(np.asarray(got), np.asarray(gen)), (igap, egap)
        n_big = 0
        for i in range(0, len(xs), 1 if (igap, egap) == (5, 2) else 5):
            if len(xs[i]) < 2 or len(ys[i]) < 2:
                continue
            want = _oracle_nw(lib, xs[i], ys[i], igap, egap)
            assert tuple(int(v) for v in got[i]) == want, (i, len(xs[i]), len(ys[i]), igap, egap)
            n_big += want[4] >= 256
        assert n_big >= 10  # pairs whose identities overflow the 8-bit field were among them


def test_align_wide_reads_matches_oracle(gpu):
    """the whole path on 300-base reads (2 x 300 sequencing) and on ragged reads of 240..321 bases (the longest of them
    need the run's score offset, nwp_core.cuh: pw_bias): every candidate pair fits the packed-word kernels, no generic
    launch is left, records == oracle == all-generic run"""
    from imsame_b200 import api
    p = api.make_params(n_threads=4)
    for case in ("fixed", "ragged"):
        if case == "fixed":
            db, ds, q, qs = sc.fixed_case(3001, 3, 60000, 300, 5000, 800, 0.03)
        else:
            db, ds, q, qs = sc.ragged_case(3002, 3, 60000, 5000, 800, 0.04, lo=240, hi=321)
        want, _ = oracle_records(db, ds, q, qs, 4)
        out, st = gpu.align((db, ds), (q, qs), p)
        assert st["k3_launches"] > 0 and st["k3_packed_launches"] == st["k3_launches"], case
        assert gpu_records(out) == want and len(want) > 100, case
        assert max(v[4] for v in want.values()) >= 256, case  # records whose identities took the second run
        gpu.set_nw_mode(1)
        try:
            out_g, st_g = gpu.align((db, ds), (q, qs), p)
        finally:
            gpu.set_nw_mode(0)
        assert st_g["k3_packed_launches"] == 0 and gpu_records(out_g) == want, case


def test_read_longer_than_the_scan_limit_is_refused(gpu):
    """documented implementation limit: a read of more than 32 767 bases -> IMSAME_ELIMIT, not a wrong answer"""
    from imsame_b200 import api
    rng = np.random.default_rng(9)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    long_read = B[rng.integers(0, 4, size=40000)]
    short = B[rng.integers(0, 4, size=300)]
    seq = np.concatenate([long_read, short])
    start = np.array([0, 40000, 40300], dtype=np.uint64)
    with pytest.raises(api.ImsameError) as e:
        gpu.align((seq, start), (short, np.array([0, 300], dtype=np.uint64)), api.make_params(n_threads=1))
    assert e.value.code == -7
    out, _ = gpu.align((short, np.array([0, 300], dtype=np.uint64)), (short, np.array([0, 300], dtype=np.uint64)),
                       api.make_params(n_threads=1))
    assert int(out["accepted"].sum()) == 1  # the context is still usable


def test_mixed_run_shares_the_bins_between_both_nw_kernels(gpu):
    """a run in which SOME reads are too long for packed words (here: query reads of 20..400 bases, database reads up
    to 600): pair by pair, the packed-word kernel takes what fits (<= 257 query / 512 database bases) and the generic
    kernel the rest -- records equal the oracle's, and equal the all-generic run"""
    from imsame_b200 import api
    db, ds, q, qs = sc.ragged_case(808, 3, 60000, 6000, 900, 0.04, lo=20, hi=400)
    # stretch a few database reads beyond 512 bases by merging neighbours
    keep = np.ones(len(ds), dtype=bool)
    keep[np.arange(5, len(ds) - 1, 37)] = False
    ds2 = ds[keep]
    assert int(np.diff(ds2).max()) > 512
    p = api.make_params(n_threads=4)
    want, _ = oracle_records(db, ds2, q, qs, 4)
    out, st = gpu.align((db, ds2), (q, qs), p)
    assert 0 < st["k3_packed_launches"] < st["k3_launches"]
    assert gpu_records(out) == want and len(want) > 200
    gpu.set_nw_mode(1)
    try:
        out_g, st_g = gpu.align((db, ds2), (q, qs), p)
    finally:
        gpu.set_nw_mode(0)
    assert st_g["k3_packed_launches"] == 0 and gpu_records(out_g) == want
    # (how many candidates the band-ordered pruning spares varies a little from run to run; the records do not)
    assert abs(st_g["n_pairs_dp"] - st["n_pairs_dp"]) < 0.05 * st["n_pairs_dp"]


def test_pair_table_growth_gives_the_same_records(monkeypatch):
    """the candidate pair table overflows and is grown (the segment is scanned again) until it holds every
    (read, database read) pair: a fresh context starts at 256 slots through the test hook IMSAME_TEST_PAIR_SLOTS"""
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(23, 3, 40000, 150, 6000, 800, 0.04)
    want, _ = oracle_records(db, ds, q, qs, 4)
    monkeypatch.setenv("IMSAME_TEST_PAIR_SLOTS", "256")
    ctx = api.Imsame(0)
    try:
        out, st = ctx.align((db, ds), (q, qs), api.make_params(n_threads=4))
        assert st["k2_launches"] >= 3 and st["n_pairs"] > 2000  # 256 -> ... slots: several scans of the one segment
        assert gpu_records(out) == want and len(want) > 300
        monkeypatch.delenv("IMSAME_TEST_PAIR_SLOTS")
        out2, st2 = ctx.align((db, ds), (q, qs), api.make_params(n_threads=4))
        assert st2["k2_launches"] == 1 and gpu_records(out2) == want
        assert (st2["n_hits"], st2["n_evalue_pass"], st2["n_pairs"]) == (st["n_hits"], st["n_evalue_pass"], st["n_pairs"])
    finally:
        ctx.close()


def _oracle_rc(db, ds, q, qs, n_threads):
    """orc_align_sequential's return code (-5 = "Read size reached for gapped alignment.") and records"""
    odb, oq = hp.OracleSeqs(seq=db, start=ds), hp.OracleSeqs(seq=q, start=qs)
    nq = len(qs) - 1
    best = (hp.OrcBest * nq)()
    st = hp.OrcStats()
    p = hp.default_params(n_threads=n_threads)
    rc = hp.oracle().orc_align_sequential(C.byref(odb.s), C.byref(oq.s), C.byref(p), best, None, C.byref(st))
    return rc, (hp.best_to_records(best, nq) if rc == 0 else None)


def test_read_size_error_exactly_when_the_reference_stops(gpu):
    """src/alignmentFunctions.c:155: the reference stops only when an e-value-passing hit with a read of more
    than 3000 bases is reached before its query read is accepted -- not whenever such a read exists"""
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(71, 3, 50000, 200, 4000, 300, 0.03)
    rng = np.random.default_rng(3)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    contig = B[rng.integers(0, 4, size=5000)]
    # (a) a long database contig no query read is related to, in the middle of the database: runs to completion
    cut = 2000 * 200
    db_a = np.concatenate([db[:cut], contig, db[cut:]])
    ds_a = np.concatenate([ds[:2001], ds[2000:] + 5000]).astype(np.uint64)
    rc, want = _oracle_rc(db_a, ds_a, q, qs, 4)
    assert rc == 0 and len(want) > 100
    out, _ = gpu.align((db_a, ds_a), (q, qs), api.make_params(n_threads=4))
    assert gpu_records(out) == want
    # (b) one query read is a piece of that contig: its first e-value-passing hit reaches NW with xlen = 5000
    q_b = q.copy()
    q_b[40 * 200:41 * 200] = contig[1000:1200]
    rc, _ = _oracle_rc(db_a, ds_a, q_b, qs, 4)
    assert rc == -5
    with pytest.raises(api.ImsameError) as e:
        gpu.align((db_a, ds_a), (q_b, qs), api.make_params(n_threads=4))
    assert e.value.code == -5
    # (c) the same piece also sits in a normal database read that the scan order reaches FIRST and accepts:
    #     list walks go in descending database position, so a copy placed after the contig wins
    db_c = db_a.copy()
    s_late = 3500
    lo = int(ds_a[s_late])
    db_c[lo:lo + 200] = contig[1000:1200]
    rc, want = _oracle_rc(db_c, ds_a, q_b, qs, 4)
    assert rc == 0 and want[40][0] == s_late
    out, _ = gpu.align((db_c, ds_a), (q_b, qs), api.make_params(n_threads=4))
    assert gpu_records(out) == want


def test_early_words_first_gives_the_same_records(gpu):
    """two-pass runs (imsame_gpu_set_passes: scan with the words that end in the first bands of their read, align,
    then scan again with the later words of the still unaccepted reads only): the records equal the oracle's and the
    one-pass run's, fewer seed hits are extended; fixed-length reads with both chunkings of the phantom word, ragged
    reads with word breaks and other thresholds, 300-base reads, a seed length other than 12"""
    from imsame_b200 import api
    rng = np.random.default_rng(5)
    cases = []
    db, ds, q, qs = sc.fixed_case(1001, 4, 100000, 150, 20000, 2000, 0.03)
    cases += [("fixed/1", db, ds, q, qs, None, dict(n_threads=1), {}), ("fixed/4", db, ds, q, qs, None, dict(n_threads=4), {})]
    db, ds, q, qs = sc.ragged_case(11, 3, 60000, 6000, 900, 0.05)
    brk = np.unique(rng.integers(1, len(db), size=300)).astype(np.uint64)
    brk = np.array([b for b in brk if b not in set(ds.tolist())], dtype=np.uint64)
    cases += [("ragged+breaks", db, ds, q, qs, brk, dict(n_threads=3), {}),
              ("ragged/thresholds", db, ds, q, qs, None, dict(n_threads=3, min_coverage=0.8, min_identity=0.9),
               dict(coverage=0.8, identity=0.9))]
    db, ds, q, qs = sc.fixed_case(3001, 3, 60000, 300, 5000, 800, 0.03)
    cases += [("wide", db, ds, q, qs, None, dict(n_threads=4), {})]
    try:
        for name, db, ds, q, qs, brk, gk, ok in cases:
            want, _ = oracle_records(db, ds, q, qs, gk["n_threads"], breaks=brk, **ok)
            p = api.make_params(**gk)
            gpu.set_passes(2)
            out2, st2 = gpu.align((db, ds), (q, qs), p, db_breaks=brk)
            gpu.set_passes(1)
            out1, st1 = gpu.align((db, ds), (q, qs), p, db_breaks=brk)
            assert st2["scan_passes"] == 2 and st1["scan_passes"] == 1, name
            assert gpu_records(out2) == want and gpu_records(out1) == want and len(want) > 100, name
            assert st2["n_hits"] < st1["n_hits"] and st2["n_db_kmers"] == st1["n_db_kmers"], name
            assert st2["k2_launches"] == 2 * st1["k2_launches"], name
        # the same context, device-resident inputs, a run of each kind after the other (tables of both kinds are kept)
        name, db, ds, q, qs, brk, gk, ok = cases[0]
        want, _ = oracle_records(db, ds, q, qs, gk["n_threads"])
        p = api.make_params(**gk)
        gpu.set_query((q, qs), p)
        gpu.set_db((db, ds))
        for mode in (2, 1, 2, 2):
            gpu.set_passes(mode)
            st = gpu.run(p)
            assert st["scan_passes"] == mode
            assert gpu_records(gpu.fetch()) == want
        gpu.set_kmer(9)
        gpu.set_passes(2)
        out, st = gpu.align((db, ds), (q, qs), p)
        assert st["scan_passes"] == 2 and gpu_records(out) == oracle_records(db, ds, q, qs, gk["n_threads"], k=9)[0]
    finally:
        gpu.set_kmer(12)
        gpu.set_passes(0)


def test_packed_and_generic_kernels_give_the_same_records(gpu):
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(41, 4, 100000, 250, 30000, 1500, 0.08)
    p = api.make_params(n_threads=4)
    gpu.set_nw_mode(1)
    gen, _ = gpu.align((db, ds), (q, qs), p)
    gpu.set_nw_mode(0)
    pk, _ = gpu.align((db, ds), (q, qs), p)
    assert gpu_records(pk) == gpu_records(gen)
    want, _ = oracle_records(db, ds, q, qs, 4)
    assert gpu_records(pk) == want and len(want) > 300


@pytest.mark.parametrize("n_threads", [1, 4])
def test_align_fixed_length_matches_oracle(gpu, n_threads):
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(1001, 4, 100000, 150, 20000, 2000, 0.03)
    want, st = oracle_records(db, ds, q, qs, n_threads)
    out, stats = gpu.align((db, ds), (q, qs), api.make_params(n_threads=n_threads))
    assert gpu_records(out) == want
    assert stats["n_accepted"] == len(want) > 500


def test_align_divergent_reads(gpu):
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(7, 4, 100000, 150, 20000, 2000, 0.10)
    want, _ = oracle_records(db, ds, q, qs, 4)
    out, _ = gpu.align((db, ds), (q, qs), api.make_params(n_threads=4))
    assert gpu_records(out) == want


def test_align_ragged_reads_and_thresholds(gpu):
    from imsame_b200 import api
    db, ds, q, qs = sc.ragged_case(11, 3, 60000, 6000, 900, 0.05)
    for kw in (dict(), dict(coverage=0.8, identity=0.9), dict(evalue=1e-5, igap=3, egap=1)):
        want, _ = oracle_records(db, ds, q, qs, 3, **kw)
        gk = dict(n_threads=3)
        if "coverage" in kw:
            gk.update(min_coverage=kw["coverage"], min_identity=kw["identity"])
        if "evalue" in kw:
            gk.update(min_e_value=kw["evalue"], igap=kw["igap"], egap=kw["egap"])
        out, _ = gpu.align((db, ds), (q, qs), api.make_params(**gk))
        assert gpu_records(out) == want, kw
        assert len(want) > 50


def test_align_with_word_breaks(gpu):
    """dropped non-ACGT characters reset database words but not query words (src/IMSAME.c:229-231)"""
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(21, 2, 50000, 120, 5000, 800, 0.02)
    rng = np.random.default_rng(3)
    brk = np.unique(rng.integers(1, len(db), size=3000)).astype(np.uint64)
    brk = np.array([b for b in brk if b % 120 != 0], dtype=np.uint64)  # not at read starts
    want, _ = oracle_records(db, ds, q, qs, 2, breaks=brk)
    want_nobrk, _ = oracle_records(db, ds, q, qs, 2)
    out, _ = gpu.align((db, ds), (q, qs), api.make_params(n_threads=2), db_breaks=brk)
    assert gpu_records(out) == want
    assert want != want_nobrk  # the breaks do change the answer


def test_sharded_database_equals_whole(gpu):
    """two shards + min-key reduction + owner payload == one shard"""
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(31, 4, 80000, 150, 16000, 1500, 0.03)
    whole, _ = gpu.align((db, ds), (q, qs), api.make_params(n_threads=4))
    nd = len(ds) - 1
    half = nd // 2
    keys, payloads = [], []
    ctx = api.Imsame(0)
    try:
        ctx.set_query((q, qs), api.make_params(n_threads=4))
        for lo, hi in ((0, half), (half, nd)):
            b0, b1 = int(ds[lo]), int(ds[hi])
            ctx.set_db((db[b0:b1], ds[lo:hi + 1] - ds[lo]))
            p = api.make_params(n_threads=4, db_total_len_global=len(db), db_pos_base=b0, db_seq_base=lo)
            ctx.run(p)
            rec = ctx.fetch()
            keys.append(rec)
        merged = {}
        for r in range(len(qs) - 1):
            cands = [k[r] for k in keys if k[r]["accepted"]]
            if cands:
                # scan order: qpos_end ascending, db_pos descending
                best = min(cands, key=lambda o: (int(o["qpos_end"]), -int(o["db_pos"])))
                merged[r] = (int(best["db_seq"]), int(best["qpos_end"]), int(best["db_pos"]), int(best["length"]),
                             int(best["identities"]))
    finally:
        ctx.close()
    assert merged == gpu_records(whole)


@pytest.mark.parametrize("divergence,identity", [(0.05, 0.95), (0.15, 0.80), (0.25, 0.70), (0.30, 0.70)])
def test_cfg5_low_identity_sweep(gpu, divergence, identity):
    """BASELINE.json configs[4] at k = 12 (the only k the reference has): divergent queries, identity
    thresholds 70-95 % -- dense candidate load, many near-threshold alignments with gaps"""
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(5005, 4, 100000, 150, 20000, 1500, divergence)
    want, _ = oracle_records(db, ds, q, qs, 4, identity=identity, coverage=0.5)
    out, _ = gpu.align((db, ds), (q, qs), api.make_params(n_threads=4, min_identity=identity, min_coverage=0.5))
    assert gpu_records(out) == want
    assert len(want) > (10 if divergence >= 0.25 else 300)


def test_cfg1_full_size(gpu):
    """BASELINE.json configs[0]: 10k x 150bp reads vs 100k-read metagenome, defaults"""
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(1001, 20, 500000, 150, 100000, 10000, 0.03)
    want, st = oracle_records(db, ds, q, qs, 4)
    out, stats = gpu.align((db, ds), (q, qs), api.make_params(n_threads=4))
    assert gpu_records(out) == want
    assert len(want) > 3000


def test_cfg2_shaped_properties_at_scale(gpu):
    """BASELINE.json configs[1] at 0.3 scale (300 k x 250 bp reads vs 3 M reads: two database segments,
    ~30 M candidate pairs) -- too large for the CPU oracle, so size-independent properties are checked:
    the packed-word and the generic NW kernel give identical records; three uneven database shards
    merged by the scan-order key equal the whole database; about half of the reads (those drawn from
    the genome pool) find a hit and every record satisfies the filter it was accepted by."""
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(2001, 300, 1000000, 250, 3000000, 300000, 0.03)
    p = api.make_params(n_threads=4)
    whole, st = gpu.align((db, ds), (q, qs), p)
    assert st["k3_packed_launches"] == st["k3_launches"] > 0
    gpu.set_nw_mode(1)
    generic, st1 = gpu.align((db, ds), (q, qs), p)
    gpu.set_nw_mode(0)
    assert st1["k3_packed_launches"] == 0
    for f in ("accepted", "db_seq", "qpos_end", "db_pos", "length", "identities"):
        assert np.array_equal(whole[f], generic[f]), f
    acc = whole["accepted"] == 1
    assert 0.45 < acc.mean() < 0.60
    assert np.all(whole["length"][acc] * 2 >= 250) and np.all(whole["identities"][acc] * 2 >= whole["length"][acc])
    assert np.all((whole["db_pos"][acc] - 1) // 250 == whole["db_seq"][acc])  # the seed lies inside the reported read
    # three uneven shards, merged on the host by (k-mer end ascending, database position descending)
    nd = len(ds) - 1
    cuts = [0, nd // 7, nd // 2 + 12345, nd]
    recs = []
    ctx = api.Imsame(0)
    try:
        ctx.set_query((q, qs), p)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            b0, b1 = int(ds[lo]), int(ds[hi])
            ctx.set_db((db[b0:b1], ds[lo:hi + 1] - ds[lo]))
            ctx.run(api.make_params(n_threads=4, db_total_len_global=len(db), db_pos_base=b0, db_seq_base=lo))
            recs.append(ctx.fetch())
    finally:
        ctx.close()
    key = np.stack([np.where(r["accepted"] == 1, r["qpos_end"].astype(np.int64) * (1 << 40) + ((1 << 40) - 1 - r["db_pos"].astype(np.int64)),
                             np.iinfo(np.int64).max) for r in recs])
    owner = key.argmin(axis=0)
    m_acc = key.min(axis=0) != np.iinfo(np.int64).max
    assert np.array_equal(m_acc, acc)
    for f in ("db_seq", "qpos_end", "db_pos", "length", "identities"):
        stacked = np.stack([r[f] for r in recs])
        merged = stacked[owner, np.arange(stacked.shape[1])]
        assert np.array_equal(np.where(m_acc, merged, 0), np.where(acc, whole[f], 0)), f


def _revcomp_reads(seq, start):
    """what revComp (src/reverseComplement.c:56-112) + the loader give for A/C/G/T-only reads: records in reverse
    order, each reverse-complemented = the whole concatenated array reversed and complemented"""
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    total = len(seq)
    return comp[seq[::-1]].copy(), (total - np.asarray(start, dtype=np.int64)[::-1]).astype(np.uint64)


def test_resident_samples_and_device_revcomp(gpu):
    """read sets uploaded once (imsame_gpu_sample_*), word table kept in the query sample,
    reverse complement made on the device from the packed form -- same records as imsame_gpu_align on host
    buffers and as the oracle; fixed-length and ragged reads with word breaks; 16-base word boundaries"""
    from imsame_b200 import api
    cases = [sc.fixed_case(88, 3, 40000, 150, 6000, 700, 0.04),
             sc.ragged_case(89, 3, 40000, 5000, 500, 0.05, lo=20, hi=300),
             sc.fixed_case(90, 2, 20000, 61, 3000, 400, 0.03)]  # total % 16 != 0
    p = api.make_params(n_threads=4)
    for ci, (db, ds, q, qs) in enumerate(cases):
        brk = None
        if ci == 1:
            rng = np.random.default_rng(5)
            brk = np.unique(rng.integers(1, len(db), size=600)).astype(np.uint64)
            brk = np.array([b for b in brk if b not in set(ds.tolist())], dtype=np.uint64)
        rdb, rds = _revcomp_reads(db, ds)
        rbrk = None if brk is None else np.sort(len(db) - brk.astype(np.int64)).astype(np.uint64)
        S_db, S_q = gpu.sample((db, ds), brk), gpu.sample((q, qs))
        S_rev = gpu.sample_revcomp(S_db)
        try:
            for _ in range(2):  # the second round reuses the word table kept in the query sample
                out, st = gpu.align_samples(S_db, S_q, p)
                want, _ = oracle_records(db, ds, q, qs, 4, breaks=brk)
                assert gpu_records(out) == want and len(want) > 30
                out_r, _ = gpu.align_samples(S_rev, S_q, p)
                host_r, _ = gpu.align((rdb, rds), (q, qs), p, db_breaks=rbrk)
                assert gpu_records(out_r) == gpu_records(host_r)
                want_r, _ = oracle_records(rdb, rds, q, qs, 4, breaks=rbrk)
                assert gpu_records(out_r) == want_r
            # a sample serves in both roles: the database sample as its own query
            self_out, _ = gpu.align_samples(S_db, S_db, p)
            self_host, _ = gpu.align((db, ds), (db, ds), p, db_breaks=brk)
            assert gpu_records(self_out) == gpu_records(self_host) and int(self_out["accepted"].sum()) > 0.8 * (len(ds) - 1)
        finally:
            for s_ in (S_db, S_q, S_rev):
                s_.free()
    # the context is still usable with host buffers afterwards
    db, ds, q, qs = cases[0]
    out, _ = gpu.align((db, ds), (q, qs), p)
    assert gpu_records(out) == oracle_records(db, ds, q, qs, 4)[0]


def _sampled_check(rec, db, ds, q, qs, reads, db_total_global=0, **kw):
    p = hp.default_params(n_threads=4, db_total_len_global=db_total_global, **kw)
    want, st = hp.oracle_align_sampled(hp.OracleSeqs(seq=db, start=ds), hp.OracleSeqs(seq=q, start=qs), p, reads)
    got = {int(r): (int(rec[r]["db_seq"]), int(rec[r]["qpos_end"]), int(rec[r]["db_pos"]), int(rec[r]["length"]),
                    int(rec[r]["identities"])) for r in reads if rec[int(r)]["accepted"]}
    return got, want, st


def test_cfg2_full_size_sampled_against_oracle(gpu):
    """BASELINE.json configs[1] at FULL size (1 M x 250 bp reads vs 10 M reads: five database segments, 32 bands,
    ~1e8 candidate pairs, pair-table growth): 1 500 randomly drawn query reads -- accepted and unaccepted --
    are re-derived by the index-free oracle (oracle/imsame_sampled.c: no database index, scan-order replay
    with the reference's early exit) and every record field must be equal."""
    from imsame_b200 import api, hostlib as H
    L, nd, nq = 250, 10_000_000, 1_000_000
    pool = H.SynthPool(2001, 1000, 1_000_000)
    db = pool.db_reads(0, nd, L)
    q = pool.query_reads(0, nq, L, 0.03)
    pool.close()
    ds = np.arange(nd + 1, dtype=np.uint64) * L
    qs = np.arange(nq + 1, dtype=np.uint64) * L
    rec, st = gpu.align((db, ds), (q, qs), api.make_params(n_threads=4))
    assert st["k2_launches"] >= 5 and st["n_pairs"] > 5e7
    reads = np.sort(np.random.default_rng(99).choice(nq, 1500, replace=False))
    got, want, ost = _sampled_check(rec, db, ds, q, qs, reads)
    assert got == want
    assert 600 < len(want) < 900 and ost.nw_calls > 20000


def test_ragged_multi_segment_sampled_against_oracle(gpu):
    """ragged reads (30..400 bases: generic NW kernel, read lookup through the block table) over a database of
    more than one segment, checked like the cfg2 test"""
    from imsame_b200 import api
    db, ds, q, qs = sc.ragged_case(4242, 300, 1000000, 3_000_000, 60000, 0.04, lo=30, hi=400)
    assert len(db) > (1 << 29)
    rec, st = gpu.align((db, ds), (q, qs), api.make_params(n_threads=4))
    assert st["k2_launches"] >= 2
    reads = np.sort(np.random.default_rng(7).choice(len(qs) - 1, 1200, replace=False))
    got, want, _ = _sampled_check(rec, db, ds, q, qs, reads)
    assert got == want and len(want) > 300


@pytest.mark.parametrize("k", [8, 10, 13, 14])
def test_runtime_seed_length_matches_generalised_oracle(gpu, k):
    """imsame_gpu_set_kmer.  The reference has FIXED_K = 12 only; for other k the checker is
    the oracle generalised to FIXED_K = k (pinned to the reference at k = 12).  Fixed-length and ragged reads,
    word breaks, phantom words (3 chunks); k = 12 afterwards gives the default result again."""
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(31 + k, 3, 40000, 120, 5000, 600, 0.05)
    rdb, rds, rq, rqs = sc.ragged_case(41 + k, 2, 30000, 2500, 400, 0.05, lo=20, hi=300)
    rng = np.random.default_rng(k)
    brk = np.unique(rng.integers(1, len(rdb), size=500)).astype(np.uint64)
    brk = np.array([b for b in brk if b not in set(rds.tolist())], dtype=np.uint64)
    try:
        gpu.set_kmer(k)
        want, st = oracle_records(db, ds, q, qs, 3, k=k)
        out, stats = gpu.align((db, ds), (q, qs), api.make_params(n_threads=3))
        assert gpu_records(out) == want and len(want) > 30
        assert stats["n_hits"] >= st.hits
        want, _ = oracle_records(rdb, rds, rq, rqs, 3, breaks=brk, k=k, evalue=1e-8)
        out, _ = gpu.align((rdb, rds), (rq, rqs), api.make_params(n_threads=3, min_e_value=1e-8), db_breaks=brk)
        assert gpu_records(out) == want and len(want) > 30
    finally:
        gpu.set_kmer(12)
    want, _ = oracle_records(db, ds, q, qs, 3)
    out, _ = gpu.align((db, ds), (q, qs), api.make_params(n_threads=3))
    assert gpu_records(out) == want


def test_seed_length_limits(gpu):
    from imsame_b200 import api
    for k in (3, 17, 0, -1):
        with pytest.raises(api.ImsameError):
            gpu.set_kmer(k)


@pytest.mark.parametrize("k", [15, 16])
def test_longest_seeds_through_properties(gpu, k):
    """k = 15 / 16 (2 x 4 GiB / 2 x 17 GiB of offsets, 64-bit table indices; the oracle's host index stops at 14):
    every record's seed is an exact k-mer match at the reported positions, inside both reads, and its
    (length, identities) are the oracle's NW result for that pair; the scan visits every database word once"""
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(5, 2, 20000, 100, 2000, 300, 0.01)
    try:
        gpu.set_kmer(k)
        out, st = gpu.align((db, ds), (q, qs), api.make_params(n_threads=1))
    finally:
        gpu.set_kmer(12)
    # words the scan saw: every position of every database read that ends a k-mer
    assert st["n_db_kmers"] == 2000 * (100 - k + 1)
    lib = hp.oracle()
    n = 0
    for r, o in enumerate(out):
        if not o["accepted"]:
            continue
        e, p, s = int(o["qpos_end"]), int(o["db_pos"]), int(o["db_seq"])
        assert int(ds[s]) <= p - k and p <= int(ds[s + 1])
        assert int(qs[r]) - 1 <= e - (k - 1) and e < int(qs[r + 1])  # - 1: the phantom word
        assert bytes(q[e - (k - 1):e + 1]) == bytes(db[p - k:p])
        x, y = db[int(ds[s]):int(ds[s + 1])], q[int(qs[r]):int(qs[r + 1])]
        assert (int(o["length"]), int(o["identities"])) == _oracle_nw(lib, x, y, 5, 2)[3:]
        n += 1
    assert n > 100

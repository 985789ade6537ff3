"""GPU: the drop-in binary (bin/IMSAME) and the traceback/render path against the reference's own
golden output (tests/golden, produced by the unmodified reference) and the oracle."""
import os
import subprocess

import numpy as np
import pytest

import helpers as hp
import synth_cases as sc

pytestmark = pytest.mark.gpu
G = os.path.join(hp.ROOT, "tests", "golden")
EXE = os.path.join(hp.ROOT, "bin", "IMSAME")
FLAGS = {"synth150": [], "dirty": ["-coverage", "0.3", "-identity", "0.6", "-evalue", "1e-10", "-igap", "4", "-egap", "1"],
         "few": ["-evalue", "1e-20"]}  # 3 query reads (fewer than threads in the t4 case) against synth150's database
DB_OF = {"few": "synth150"}


def db_path(name):
    return os.path.join(G, f"{DB_OF.get(name, name)}.db.fa")


def info_lines(stdout):
    return "".join(l + "\n" for l in stdout.splitlines() if "from the query were found" in l or "Jaccard" in l)


@pytest.mark.parametrize("name", sorted(FLAGS))
def test_cli_t1_output_is_byte_identical_to_reference(gpu, name, tmp_path):
    out = str(tmp_path / "o.align")
    r = subprocess.run([EXE, "-query", os.path.join(G, f"{name}.q.fa"), "-db", db_path(name),
                        "-out", out, "-n_threads", "1"] + FLAGS[name], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert open(out, "rb").read() == open(os.path.join(G, f"{name}.t1.align"), "rb").read()
    assert info_lines(r.stdout) == open(os.path.join(G, f"{name}.stdout")).read()


@pytest.mark.parametrize("name", sorted(FLAGS))
def test_cli_t4_headers_equal_reference(gpu, name, tmp_path):
    out = str(tmp_path / "o.align")
    subprocess.check_call([EXE, "-query", os.path.join(G, f"{name}.q.fa"), "-db", db_path(name),
                           "-out", out, "-n_threads", "4"] + FLAGS[name], stdout=subprocess.DEVNULL)
    got = sorted(l for l in open(out, "rb").read().split(b"\n") if hp.HEADER_RE.match(l))
    want = [l for l in open(os.path.join(G, f"{name}.t4.headers"), "rb").read().split(b"\n") if l]
    assert got == want


def test_cli_matches_oracle_file_on_fresh_input(gpu, tmp_path):
    """P2: whole .align file == oracle's (== reference's, test_oracle_golden) with -n_threads 1"""
    from imsame_b200 import hostlib as H
    pool = H.SynthPool(2024, 3, 60000)
    nd, nq, L = 8000, 700, 250
    dbf, qf, mine, orc = (str(tmp_path / n) for n in ("db.fa", "q.fa", "mine.align", "orc.align"))
    H.write_fasta(dbf, pool.db_reads(0, nd, L), nd, L, "d")
    H.write_fasta(qf, pool.query_reads(0, nq, L, 0.06), nq, L, "q")
    pool.close()
    subprocess.check_call([EXE, "-query", qf, "-db", dbf, "-out", mine, "-n_threads", "1"], stdout=subprocess.DEVNULL)
    hp.oracle_align(hp.OracleSeqs(dbf, True), hp.OracleSeqs(qf, False), hp.default_params(n_threads=1), out_path=orc)
    assert open(mine, "rb").read() == open(orc, "rb").read()
    assert len(hp.parse_align_headers(mine)) > 150


def test_cli_kmer_flag_matches_generalised_oracle(gpu, tmp_path):
    """-kmer (no such flag in the reference: FIXED_K = 12 is compiled in) against the oracle with FIXED_K = 10"""
    dbf, qf = os.path.join(G, "dirty.db.fa"), os.path.join(G, "dirty.q.fa")
    mine, orc = str(tmp_path / "mine.align"), str(tmp_path / "orc.align")
    subprocess.check_call([EXE, "-query", qf, "-db", dbf, "-out", mine, "-n_threads", "1", "-kmer", "10", "-evalue", "1e-10",
                           "-coverage", "0.3", "-identity", "0.6", "-igap", "4", "-egap", "1"], stdout=subprocess.DEVNULL)
    hp.oracle_align(hp.OracleSeqs(dbf, True), hp.OracleSeqs(qf, False),
                    hp.default_params(n_threads=1, evalue=1e-10, coverage=0.3, identity=0.6, igap=4, egap=1, k=10), out_path=orc)
    assert open(mine, "rb").read() == open(orc, "rb").read()
    assert len(hp.parse_align_headers(mine)) > 10


def test_cli_without_out_prints_summary_only(gpu):
    r = subprocess.run([EXE, "-query", os.path.join(G, "synth150.q.fa"), "-db", os.path.join(G, "synth150.db.fa"),
                        "-n_threads", "1"], capture_output=True, text=True)
    assert r.returncode == 0 and info_lines(r.stdout) == open(os.path.join(G, "synth150.stdout")).read()
    assert "Going from 0 to 90" in r.stdout


def test_cli_read_size_limit(gpu, tmp_path):
    """reads longer than MAX_READ_SIZE that reach NW -> the reference's terror text, exit 255"""
    rng = np.random.default_rng(1)
    s = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=3200)])
    f = tmp_path / "long.fa"
    f.write_bytes(b">a\n" + s + b"\n>b\n" + s[100:] + s[:100] + b"\n")
    r = subprocess.run([EXE, "-query", str(f), "-db", str(f), "-n_threads", "1"], capture_output=True, text=True)
    assert r.returncode == 255 and "ERR**** Read size reached for gapped alignment. ****" in r.stdout


def test_all_vs_all_script(gpu, tmp_path):
    """bin/all_vs_all_metagenomes_IMSAME.sh: 3 samples -> 3 pairs x (forward + reverse-complement) outputs"""
    from imsame_b200 import hostlib as H
    d, o = tmp_path / "samples", tmp_path / "out"
    d.mkdir(); o.mkdir()
    pool = H.SynthPool(4001, 3, 30000)
    for i in range(3):
        H.write_fasta(str(d / f"s{i}.fasta"), pool.db_reads(i * 1000, 600, 150), 600, 150, "r")
    pool.close()
    subprocess.run([os.path.join(hp.ROOT, "bin", "all_vs_all_metagenomes_IMSAME.sh"), str(d), "0.5", "0.5", "4", "fasta",
                    str(o)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    names = sorted(os.listdir(o))
    assert names == sorted(f"s{i}-s{j}{r}.align" for i in range(3) for j in range(i + 1, 3) for r in ("", ".r"))
    # forward comparisons find overlaps (same genome pool); reverse-complement ones find (almost) none
    assert len(hp.parse_align_headers(str(o / "s0-s1.align"))) > 50
    assert not any(f.endswith(".r.fasta") for f in os.listdir(d))
    if hp.have_reference():
        ref = str(tmp_path / "ref.align")
        hp.run_reference(str(d / "s0.fasta"), str(d / "s1.fasta"), ref, n_threads=4)
        assert hp.parse_align_headers(ref) == hp.parse_align_headers(str(o / "s0-s1.align"))


def test_in_process_all_vs_all_equals_the_script(gpu, tmp_path):
    """bin/IMSAME_allvsall (one process, samples parsed once, reverse complements built in memory) writes
    byte for byte the files the script produces with 2 IMSAME + 1 revComp process per pair, and keeps the
    script's resume rule.  Sample 2 is "dirty" (lower case, N, multi-line, '>' inside a header) so that
    revComp's record-order reversal and letter handling matter."""
    from imsame_b200 import hostlib as H
    d, o1, o2 = tmp_path / "samples", tmp_path / "out_script", tmp_path / "out_batch"
    d.mkdir(); o1.mkdir(); o2.mkdir()
    pool = H.SynthPool(4001, 3, 30000)
    for i in range(3):
        H.write_fasta(str(d / f"s{i}.fasta"), pool.db_reads(i * 1000, 500, 150), 500, 150, "r")
    reads = pool.db_reads(5000, 300, 150)
    pool.close()
    with open(d / "s3.fasta", "wb") as f:
        f.write(b"text before the first header\n")
        for r in range(300):
            b = bytes(reads[r * 150:(r + 1) * 150])
            if r % 3 == 0:
                b = b[:40].lower() + b"N" + b[40:]
            f.write(b">x%d with > inside\n" % r + b[:70] + b"\n" + b[70:] + b"\n")
    args = [str(d), "0.5", "0.5", "4", "fasta"]
    subprocess.run([os.path.join(hp.ROOT, "bin", "all_vs_all_metagenomes_IMSAME.sh")] + args + [str(o1)],
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(hp.ROOT, "bin", "IMSAME_allvsall")] + args + [str(o2)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    names = sorted(os.listdir(o1))
    assert names == sorted(os.listdir(o2)) and len(names) == 12
    for n in names:
        assert open(o1 / n, "rb").read() == open(o2 / n, "rb").read(), n
    assert os.path.getsize(o2 / "s0-s1.align") > 10000 and os.path.getsize(o2 / "s0-s3.align") > 1000
    assert "12 comparisons of 4 samples" in r.stdout
    # resume: existing outputs are kept, only the missing one is recomputed
    os.remove(o2 / "s1-s3.r.align")
    (o2 / "s0-s1.align").write_bytes(b"kept")
    r = subprocess.run([os.path.join(hp.ROOT, "bin", "IMSAME_allvsall")] + args + [str(o2)], capture_output=True, text=True)
    assert "1 comparisons of 4 samples" in r.stdout
    assert (o2 / "s0-s1.align").read_bytes() == b"kept"
    assert open(o1 / "s1-s3.r.align", "rb").read() == open(o2 / "s1-s3.r.align", "rb").read()


def test_traceback_api_matches_oracle_text(gpu):
    from imsame_b200 import api, hostlib as H
    db, ds, q, qs = sc.ragged_case(17, 2, 50000, 3000, 300, 0.08)
    out, _ = gpu.align((db, ds), (q, qs), api.make_params(n_threads=2))
    ops_off, ops, cell = gpu.traceback((db, ds), (q, qs), out, api.make_params(n_threads=2))
    lib = hp.oracle()
    import ctypes as C
    u8p = C.POINTER(C.c_ubyte)
    n_checked = 0
    for r in np.nonzero(out["accepted"])[0][:120]:
        s = int(out[r]["db_seq"])
        x = np.ascontiguousarray(db[int(ds[s]):int(ds[s + 1])])
        y = np.ascontiguousarray(q[int(qs[r]):int(qs[r + 1])])
        text = C.create_string_buffer(8 * (len(x) + len(y)) + 512)
        sc_, bx, by, ln, idn = C.c_int32(), C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib.orc_nw_traceback(x.ctypes.data_as(u8p), len(x), y.ctypes.data_as(u8p), len(y), -5, -2, C.byref(sc_),
                             C.byref(bx), C.byref(by), C.byref(ln), C.byref(idn), text, len(text))
        rec = H.render_record(int(r), s, int(out[r]["length"]), int(out[r]["identities"]), x, y, cell[r, 0],
                              cell[r, 1], ops[int(ops_off[r]):int(ops_off[r + 1])])
        hdr_len = rec.index(b" $$$$$$$ \n") + len(b" $$$$$$$ \n")
        assert rec[hdr_len:] == text.value
        assert (int(cell[r, 0]), int(cell[r, 1]), int(out[r]["length"]), int(out[r]["identities"])) == \
            (bx.value, by.value, ln.value, idn.value)
        n_checked += 1
    assert n_checked > 40

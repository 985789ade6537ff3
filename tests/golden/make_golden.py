#!/usr/bin/env python
"""Regenerates tests/golden/* by running the UNMODIFIED reference (oracle/_ref/IMSAME, compiled
from /root/reference/src by oracle/build_ref.sh) on small seeded inputs.  Only runs where the
reference exists (the authoring container); the fixtures are committed so that the oracle stays
pinned on machines without it.

Fixtures per case <name>:
  <name>.db.fa / <name>.q.fa      inputs (synthetic, imsame_b200/host/synth.c, or hand-written)
  <name>.t1.align                 reference output with -n_threads 1 (byte-exact target)
  <name>.t4.headers               sorted record header lines with -n_threads 4 (set target)
  <name>.stdout                   the two data-dependent [INFO] summary lines (t1)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from imsame_b200 import hostlib as H  # noqa: E402
import helpers as hp  # noqa: E402


def summary_lines(stdout):
    return "".join(l + "\n" for l in stdout.splitlines() if "from the query were found" in l or "Jaccard" in l)


def run_case(name, extra=(), db_name=None):
    db = os.path.join(HERE, f"{db_name or name}.db.fa")
    q = os.path.join(HERE, f"{name}.q.fa")
    out1 = os.path.join(HERE, f"{name}.t1.align")
    so = hp.run_reference(q, db, out1, n_threads=1, extra=extra)
    open(os.path.join(HERE, f"{name}.stdout"), "w").write(summary_lines(so))
    tmp = os.path.join(HERE, f"{name}.t4.tmp")
    hp.run_reference(q, db, tmp, n_threads=4, extra=extra)
    heads = sorted(l for l in open(tmp, "rb").read().split(b"\n") if hp.HEADER_RE.match(l))
    open(os.path.join(HERE, f"{name}.t4.headers"), "wb").write(b"\n".join(heads) + b"\n")
    os.remove(tmp)
    print(name, len(heads), "records")


def main():
    assert hp.have_reference(), "oracle/_ref/IMSAME missing: run oracle/build_ref.sh"
    # synthetic, fixed length 150, indel-bearing divergent reads
    pool = H.SynthPool(424242, 2, 20000)
    nd, nq, L = 700, 90, 150
    H.write_fasta(os.path.join(HERE, "synth150.db.fa"), pool.db_reads(0, nd, L), nd, L, "d")
    H.write_fasta(os.path.join(HERE, "synth150.q.fa"), pool.query_reads(0, nq, L, 0.08), nq, L, "q")
    pool.close()
    run_case("synth150")
    # hand-made ragged / dirty FASTA: multi-line records, lower case, N runs, CR line ends, junk before '>'
    pool = H.SynthPool(77, 1, 6000)
    reads = pool.db_reads(0, 160, 200)
    qreads = pool.query_reads(0, 40, 200, 0.05)
    pool.close()
    import numpy as np
    rng = np.random.default_rng(5)

    def dirty(arr, n, path, is_db):
        with open(path, "wb") as f:
            f.write(b"junk before the first header\nACGTACGT\n")
            for i in range(n):
                ln = int(rng.integers(40, 201))
                s = bytearray(arr[i * 200:i * 200 + ln].tobytes())
                if i % 3 == 0:
                    s = bytearray(bytes(s).lower())
                if i % 4 == 1 and ln > 60:
                    k = int(rng.integers(20, ln - 20))
                    s[k:k + 1] = b"NN" if i % 8 == 1 else b"n"
                f.write(b">r%d some description > with a bracket\n" % i)
                w = int(rng.integers(30, 81))
                eol = b"\r\n" if i % 5 == 2 else b"\n"
                for o in range(0, len(s), w):
                    f.write(bytes(s[o:o + w]) + eol)
    dirty(reads, 160, os.path.join(HERE, "dirty.db.fa"), True)
    dirty(qreads, 40, os.path.join(HERE, "dirty.q.fa"), False)
    run_case("dirty", extra=("-coverage", "0.3", "-identity", "0.6", "-evalue", "1e-10", "-igap", "4", "-egap", "1"))
    # 3 query reads: -n_threads 4 > n_seqs (reads_per_thread = 0, src/IMSAME.c:414,452) and an explicit
    # "-evalue 1e-20" (atof -> double -> long double, src/IMSAME.c:553: not the default's bit pattern, :44)
    lines = open(os.path.join(HERE, "synth150.q.fa"), "rb").read().split(b"\n")
    picked, n = [], 0
    for i in range(0, len(lines) - 1, 2):
        if i // 2 in (1, 4, 7):
            picked += [lines[i], lines[i + 1]]
    open(os.path.join(HERE, "few.q.fa"), "wb").write(b"\n".join(picked) + b"\n")
    run_case("few", extra=("-evalue", "1e-20"), db_name="synth150")


if __name__ == "__main__":
    main()

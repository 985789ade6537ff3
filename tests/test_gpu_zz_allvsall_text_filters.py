"""All-vs-all on samples whose reverse complement is NOT the mirror image of the sample (GPU; named to run after the
other GPU files).  revComp keeps letters only (src/reverseComplement.c:65-70) while the database loader restarts the
seed word at every dropped character but the newline (src/IMSAME.c:229-231): in a multi-line CRLF file, or one with
gap characters inside its records, the sample has word breaks its reverse complement does not have; 'U' comes back
as an 'A' the loader keeps.  bin/IMSAME_allvsall may derive a reverse-complemented read set on the device only when
imsame_revcomp_is_mirror says so (tests/test_host_cpu.py pins that predicate on the CPU); here the outputs of the
in-process driver are compared with the script's (one IMSAME + revComp process per comparison, text all the way) and,
when oracle/_ref travelled, with the unmodified reference workflow.  The samples hold reverse-strand reads so that
the .r.align files are not empty."""
import os
import subprocess

import pytest

import helpers as hp
import synth_cases as sc

pytestmark = pytest.mark.gpu


def test_all_vs_all_with_samples_that_are_not_their_mirror(gpu, tmp_path):
    d, o1, o2, o3 = (tmp_path / n for n in ("samples", "out_script", "out_batch", "out_reference"))
    for p in (d, o1, o2, o3):
        p.mkdir()
    sc.write_allvsall_filter_samples(d)
    from imsame_b200 import hostlib as H
    assert not H.revcomp_is_mirror((d / "s1.fasta").read_bytes()) and not H.revcomp_is_mirror((d / "s2.fasta").read_bytes())
    assert H.revcomp_is_mirror((d / "s0.fasta").read_bytes()) and H.revcomp_is_mirror((d / "s3.fasta").read_bytes())
    args = [str(d), "0.5", "0.5", "2", "fasta"]
    subprocess.run([os.path.join(hp.ROOT, "bin", "all_vs_all_metagenomes_IMSAME.sh")] + args + [str(o1)],
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(hp.ROOT, "bin", "IMSAME_allvsall")] + args + [str(o2)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    names = sorted(os.listdir(o1))
    assert names == sorted(os.listdir(o2)) and len(names) == 12
    for n in names:
        assert open(o1 / n, "rb").read() == open(o2 / n, "rb").read(), n
    # the reverse-strand samples match the forward one through their reverse complement
    for n in ("s0-s1.r.align", "s0-s2.r.align", "s0-s3.r.align"):
        assert os.path.getsize(o2 / n) > 5000, n
    ref_script = os.path.join(hp.ROOT, "oracle", "_ref", "all_vs_all_metagenomes_IMSAME.sh")
    if os.path.exists(ref_script) and hp.have_reference():
        subprocess.run(["bash", ref_script] + args + [str(o3)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert sorted(os.listdir(o3)) == names
        for n in names:  # two threads share the reference's output stream: header sets and sorted lines (tools/allvsall_bench.py)
            assert hp.parse_align_headers(str(o3 / n)) == hp.parse_align_headers(str(o2 / n)), n
            assert sorted(open(o3 / n, "rb").read().split(b"\n")) == sorted(open(o2 / n, "rb").read().split(b"\n")), n

#!/usr/bin/env python
"""End-to-end run of the drop-in CLI (bin/IMSAME) on a synthetic cfg2-shaped FASTA pair:
FASTA files on disk -> .align file.  usage: python tools/cli_e2e.py --scale 1.0 [--gpus N | --gpus-list 1,2]
With --gpus-list the command line runs once per entry (database sharded over that many GPUs inside the library,
imsame_gpu_align_sharded) and the output files are compared byte for byte."""
import argparse, os, subprocess, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imsame_b200 import hostlib as H  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--gpus-list", default=None)
ap.add_argument("--dir", default=None)
a = ap.parse_args()
nd, nq, g, L = int(10_000_000 * a.scale), int(1_000_000 * a.scale), max(2, int(1000 * a.scale)), 250
d = a.dir or tempfile.mkdtemp(prefix="imsame_cli_")
t = time.time()
pool = H.SynthPool(2001, g, 1_000_000)
db = pool.db_reads(0, nd, L); q = pool.query_reads(0, nq, L, 0.03); pool.close()
H.write_fasta(os.path.join(d, "db.fa"), db, nd, L, "d"); H.write_fasta(os.path.join(d, "q.fa"), q, nq, L, "q")
del db, q
print(f"generated + wrote FASTA in {time.time() - t:.1f}s: db {os.path.getsize(os.path.join(d, 'db.fa')) / 1e9:.2f} GB", flush=True)
env = dict(os.environ, IMSAME_TRACE="1")
outs = []
for gpus in ([int(x) for x in a.gpus_list.split(",")] if a.gpus_list else [a.gpus]):
    out = os.path.join(d, f"out_g{gpus}.align")
    outs.append(out)
    # run 0 warms the page cache; run 1 releases everything piece by piece at the end (IMSAME_FAST_EXIT=0); run 2
    # leaves the release to the operating system (the default) and must write the same file
    for rep, fast in enumerate(("0", "0", "1")):
        o = out + (".fast" if fast == "1" else "")
        t = time.time()
        r = subprocess.run([os.path.join(ROOT, "bin", "IMSAME"), "-query", os.path.join(d, "q.fa"), "-db", os.path.join(d, "db.fa"),
                            "-out", o, "-gpus", str(gpus)], capture_output=True, text=True, env=dict(env, IMSAME_FAST_EXIT=fast))
        wall = time.time() - t
        info = [l for l in r.stdout.splitlines() if l.startswith("[INFO]")]
        print(f"-gpus {gpus} run {rep} (IMSAME_FAST_EXIT={fast}): rc={r.returncode} wall {wall:.2f}s  -> {nq / wall:.0f} query reads/s "
              f"end to end, out {os.path.getsize(o) / 1e6:.0f} MB", flush=True)
        print("\n".join(info[-6:])); print(r.stderr[-900:], flush=True)
    import filecmp
    print("fast-exit output identical:", filecmp.cmp(out, out + ".fast", shallow=False), flush=True)
    os.remove(out + ".fast")
if len(outs) > 1:
    import filecmp
    print("output files identical across -gpus settings:", all(filecmp.cmp(outs[0], o, shallow=False) for o in outs[1:]))

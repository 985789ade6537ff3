"""GPU, >= 2 devices: database sharded over two ranks with the NCCL min-key / owner-payload
reduction (exactly what bench.py runs) equals the single-GPU result; CLI -gpus 2 equals -gpus 1."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as hp
import synth_cases as sc

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, os.path.join(hp.ROOT, "tests"))
    from imsame_b200 import api, sharding
    db, ds, q, qs = sc.fixed_case(31, 4, 80000, 150, 16000, 1500, 0.03)
    nd, nq = len(ds) - 1, len(qs) - 1
    lo, hi = sharding.shard_range(nd, rank, world)
    b0, b1 = int(ds[lo]), int(ds[hi])
    ctx = api.Imsame(rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_passes(int(os.environ.get("IMSAME_TEST_PASSES", "0")))
    p = api.make_params(n_threads=4, db_total_len_global=len(db), db_pos_base=b0, db_seq_base=lo)
    mode = os.environ.get("IMSAME_TEST_STEPPED", "1")
    keys = torch.empty(nq, dtype=torch.int64, device="cuda")
    payload = torch.empty(nq, dtype=torch.int64, device="cuda")
    if mode != "align_shard":
        ctx.set_query((q, qs), p)
        ctx.set_db((db[b0:b1], ds[lo:hi + 1] - ds[lo]))
    if mode == "align_shard":
        # the whole per-rank job in one collective call: uploads (a segment ahead of the scan) + table + sharded run
        box = [api.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], world, rank)
        st = ctx.align_shard((db[b0:b1], ds[lo:hi + 1] - ds[lo]), (q, qs), p, keys.data_ptr(), payload.data_ptr())
        assert st["ms_comm"] > 0 and st["h2d_bytes"] >= (b1 - b0) + len(q)
        assert st["scan_passes"] == max(1, int(os.environ.get("IMSAME_TEST_PASSES", "0")))
    elif mode == "nccl_in_library":
        # the product path: the library's own communicator (ncclCommInitRank from 128 bytes handed around by
        # torch.distributed) and its band-stepped run with ncclMin / ncclMax reductions inside
        box = [api.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], world, rank)
        st = ctx.run_sharded(p, keys.data_ptr(), payload.data_ptr())
        assert st["ms_comm"] > 0
        assert st["scan_passes"] == max(1, int(os.environ.get("IMSAME_TEST_PASSES", "0")))
    elif mode == "1":
        # keys exchanged between bands (what bench.py does), payload of the owner at the end
        ctx.run_stepped(p, keys.data_ptr(), payload.data_ptr(),
                        exchange=lambda: dist.all_reduce(keys, op=dist.ReduceOp.MIN), exchange_every=3)
        dist.all_reduce(payload, op=dist.ReduceOp.MAX)
    else:
        ctx.run(p, keys.data_ptr(), payload.data_ptr())

        def mask(kr, kl, pl):
            ctx.mask_payload(kr.data_ptr(), kl.data_ptr(), pl.data_ptr())
        sharding.reduce_best(keys, payload, dist, mask)
    rec = ctx.fetch(keys.data_ptr(), payload.data_ptr())
    if rank == 0:
        np.save(os.path.join(tmpdir, "rec.npy"), rec)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("stepped,passes", [("nccl_in_library", 0), ("align_shard", 0), ("1", 0), ("0", 0),
                                            ("nccl_in_library", 2), ("align_shard", 2), ("0", 2)])
def test_nccl_sharded_equals_single_gpu(gpu, tmp_path, stepped, passes):
    """passes = 2: early words first (imsame_gpu_set_passes) -- in the library's sharded run the keys are reduced
    between the two scans, so a read accepted in one shard drops out of the second scan of both"""
    import torch.multiprocessing as mp
    from imsame_b200 import api
    os.environ["IMSAME_TEST_STEPPED"] = stepped
    os.environ["IMSAME_TEST_PASSES"] = str(passes)
    port = 29700 + (os.getpid() % 1000) + len(stepped) + 7 * passes
    try:
        mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    finally:
        os.environ.pop("IMSAME_TEST_PASSES", None)
    rec = np.load(tmp_path / "rec.npy")
    db, ds, q, qs = sc.fixed_case(31, 4, 80000, 150, 16000, 1500, 0.03)
    whole, _ = gpu.align((db, ds), (q, qs), api.make_params(n_threads=4))
    for f in ("accepted", "db_seq", "qpos_end", "db_pos", "length", "identities"):
        assert np.array_equal(rec[f], whole[f]), f
    assert int(whole["accepted"].sum()) > 300


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_align_sharded_in_one_process_equals_oracle(gpu):
    """imsame_gpu_align_sharded: one process, one context per GPU, NCCL inside the library; ragged reads and
    word breaks so that shard boundaries, global coordinates and the break lists are all exercised"""
    from imsame_b200 import api
    db, ds, q, qs = sc.ragged_case(515, 3, 50000, 9000, 700, 0.05, lo=30, hi=300)
    rng = np.random.default_rng(1)
    brk = np.unique(rng.integers(1, len(db), size=800)).astype(np.uint64)
    brk = np.array([b for b in brk if b not in set(ds.tolist())], dtype=np.uint64)
    odb, oq = hp.OracleSeqs(seq=db, start=ds, brk=brk), hp.OracleSeqs(seq=q, start=qs)
    want = hp.best_to_records(hp.oracle_align(odb, oq, hp.default_params(n_threads=4))[0], len(qs) - 1)
    n = min(_n_gpus(), 4)
    ctxs = [gpu] + [api.Imsame(d) for d in range(1, n)]
    try:
        for it in range(3):  # the later calls reuse the communicator kept in the contexts; the last: early words first
            for c in ctxs:
                c.set_passes(2 if it == 2 else 0)
            out, stats = api.align_sharded(ctxs, (db, ds), (q, qs), api.make_params(n_threads=4), db_breaks=brk)
            assert all(s["scan_passes"] == (2 if it == 2 else 1) for s in stats)
            got = {int(r): (int(o["db_seq"]), int(o["qpos_end"]), int(o["db_pos"]), int(o["length"]), int(o["identities"]))
                   for r, o in enumerate(out) if o["accepted"]}
            assert got == want and len(want) > 150
            assert len(stats) == n and all(s["ms_comm"] > 0 and s["n_db_kmers"] > 0 for s in stats)
            assert sum(s["n_db_kmers"] for s in stats) == gpu.align((db, ds), (q, qs), api.make_params(n_threads=4),
                                                                     db_breaks=brk)[1]["n_db_kmers"]
    finally:
        gpu.set_passes(0)
        for c in ctxs[1:]:
            c.close()
        gpu.comm_free()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_align_sharded_read_size_error_is_a_property_of_the_whole_database(gpu):
    """"Read size reached for gapped alignment." (src/alignmentFunctions.c:155) in a sharded run: the long contig
    lies in ONE shard; every shard must come back with the same answer (the flag is reduced with the payload), and a
    later, normal database read that the scan order reaches first -- in the OTHER shard -- must still suppress it"""
    from imsame_b200 import api
    db, ds, q, qs = sc.fixed_case(71, 3, 50000, 200, 4000, 300, 0.03)
    rng = np.random.default_rng(3)
    contig = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=5000)]
    cut = 1000 * 200  # the contig sits in the first shard
    db_a = np.concatenate([db[:cut], contig, db[cut:]])
    ds_a = np.concatenate([ds[:1001], ds[1000:] + 5000]).astype(np.uint64)
    q_b = q.copy()
    q_b[40 * 200:41 * 200] = contig[1000:1200]
    p = api.make_params(n_threads=4)
    ctxs = [gpu, api.Imsame(1)]
    try:
        with pytest.raises(api.ImsameError) as e:
            api.align_sharded(ctxs, (db_a, ds_a), (q_b, qs), p)
        assert e.value.code == -5
        # the same piece in a normal read of the SECOND shard, later in the database: reached first, accepted
        db_c = db_a.copy()
        s_late = 3500
        lo = int(ds_a[s_late])
        db_c[lo:lo + 200] = contig[1000:1200]
        out, _ = api.align_sharded(ctxs, (db_c, ds_a), (q_b, qs), p)
        assert int(out[40]["accepted"]) == 1 and int(out[40]["db_seq"]) == s_late
        whole, _ = gpu.align((db_c, ds_a), (q_b, qs), p)
        for f in ("accepted", "db_seq", "qpos_end", "db_pos", "length", "identities"):
            assert np.array_equal(out[f], whole[f]), f
    finally:
        ctxs[1].close()
        gpu.comm_free()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_cli_two_gpus_equals_one(gpu, tmp_path):
    from imsame_b200 import hostlib as H
    pool = H.SynthPool(606, 3, 60000)
    nd, nq, L = 9000, 800, 150
    dbf, qf = str(tmp_path / "db.fa"), str(tmp_path / "q.fa")
    H.write_fasta(dbf, pool.db_reads(0, nd, L), nd, L, "d")
    H.write_fasta(qf, pool.query_reads(0, nq, L, 0.05), nq, L, "q")
    pool.close()
    exe = os.path.join(hp.ROOT, "bin", "IMSAME")
    outs = []
    for g in ("1", "2"):
        o = str(tmp_path / f"o{g}.align")
        subprocess.check_call([exe, "-query", qf, "-db", dbf, "-out", o, "-n_threads", "4", "-gpus", g],
                              stdout=subprocess.DEVNULL)
        outs.append(open(o, "rb").read())
    assert outs[0] == outs[1] and len(outs[0]) > 10000

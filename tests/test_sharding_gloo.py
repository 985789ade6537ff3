"""N > 1 host logic on CPU: two gloo ranks, each holding the oracle's per-shard best hits,
reduced exactly like bench.py does on NCCL (min over packed keys, owner mask, max over payload)."""
import os
import sys

import numpy as np
import pytest

import helpers as hp
import synth_cases as sc

KEY_NONE = 0x7FFFFFFFFFFFFFFF
POS_BITS = 40
POS_MASK = (1 << POS_BITS) - 1


def _worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(hp.ROOT, "tests"))
    from imsame_b200 import sharding
    db, ds, q, qs = sc.fixed_case(31, 3, 40000, 150, 3000, 400, 0.04)
    nd, nq = len(ds) - 1, len(qs) - 1
    lo, hi = nd * rank // world, nd * (rank + 1) // world
    b0, b1 = int(ds[lo]), int(ds[hi])
    odb = hp.OracleSeqs(seq=db[b0:b1], start=ds[lo:hi + 1] - ds[lo])
    oq = hp.OracleSeqs(seq=q, start=qs)
    best, _ = hp.oracle_align(odb, oq, hp.default_params(n_threads=4, db_total_len_global=len(db)))
    keys = np.full(nq, KEY_NONE, dtype=np.int64)
    payload = np.zeros(nq, dtype=np.int64)
    for r in range(nq):
        if best[r].accepted:
            keys[r] = sharding.make_key(best[r].qpos_end - int(qs[r]) + 1, best[r].db_pos + b0)
            payload[r] = sharding.make_payload(best[r].db_seq + lo, best[r].length, best[r].identities)
    k, p = torch.from_numpy(keys), torch.from_numpy(payload)
    sharding.reduce_best(k, p, dist, mask_fn=sharding.mask_payload_torch)
    if rank == 0:
        np.save(os.path.join(tmpdir, "keys.npy"), k.numpy())
        np.save(os.path.join(tmpdir, "payload.npy"), p.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_equals_whole_database(built, tmp_path):
    import torch.multiprocessing as mp
    from imsame_b200 import sharding
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    keys = np.load(tmp_path / "keys.npy")
    payload = np.load(tmp_path / "payload.npy")
    db, ds, q, qs = sc.fixed_case(31, 3, 40000, 150, 3000, 400, 0.04)
    whole, _ = hp.oracle_align(hp.OracleSeqs(seq=db, start=ds), hp.OracleSeqs(seq=q, start=qs),
                               hp.default_params(n_threads=4))
    want = hp.best_to_records(whole, len(qs) - 1)
    got = sharding.decode(keys, payload, qs)
    assert got == want and len(want) > 50

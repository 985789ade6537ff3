"""Database sharding across GPUs: contiguous read ranges per rank, the query
replicated, one min-reduction of the packed scan-order keys, then the owner's payload.

key     = (k-mer end inside the read + 1) << 40 | (2^40 - 1 - global db position)   (smaller = earlier
          in the reference's scan: src/alignmentFunctions.c:91-203, lists in descending position)
payload = global db read << 32 | length << 16 | identities
Both fit int64 (torch/NCCL reduce them as signed; values stay below 2^63).
"""
import numpy as np

KEY_NONE = 0x7FFFFFFFFFFFFFFF
POS_BITS = 40
POS_MASK = (1 << POS_BITS) - 1


def make_key(e_rel, db_pos_global):
    return (int(e_rel) << POS_BITS) | (POS_MASK - int(db_pos_global))


def make_payload(db_seq_global, length, identities):
    return (int(db_seq_global) << 32) | (int(length) << 16) | int(identities)


def shard_range(n_reads, rank, world):
    return n_reads * rank // world, n_reads * (rank + 1) // world


def mask_payload_torch(keys_reduced, keys_local, payload):
    """CPU/torch statement of imsame_gpu_mask_payload (tests only; GPUs use the kernel)"""
    payload[(keys_reduced != keys_local) | (keys_reduced == KEY_NONE)] = 0


def reduce_best(keys, payload, dist, mask_fn):
    """in place: keys <- min over ranks; payload <- payload of the rank that owns the winning key"""
    local = keys.clone()
    dist.all_reduce(keys, op=dist.ReduceOp.MIN)
    mask_fn(keys, local, payload)
    dist.all_reduce(payload, op=dist.ReduceOp.MAX)
    return keys, payload


def decode(keys, payload, q_start):
    """{read: (db_seq, qpos_end, db_pos, length, identities)}"""
    out = {}
    for r in np.nonzero(np.asarray(keys) != KEY_NONE)[0]:
        k, p = int(keys[r]), int(payload[r])
        out[int(r)] = (p >> 32, int(q_start[r]) + (k >> POS_BITS) - 1, POS_MASK - (k & POS_MASK), (p >> 16) & 0xFFFF,
                       p & 0xFFFF)
    return out

"""Host-side helpers for doc-range sharding: the top-k key format the
kernels exchange and the merge of per-shard lists (the device does the same merge in
topk_merge_kernel; this numpy version serves hosts that gather the lists themselves and the
CPU tests of the multi-rank logic)."""
import numpy as np


def shard_topk_keys(doc_ids, scores, k):
    """(global doc ids, scores) of one shard -> its k best as keys
    score_bits << 32 | (0xFFFFFFFF - doc), descending, zero-padded.  Only score > 0 counts."""
    doc_ids = np.asarray(doc_ids, dtype=np.uint64)
    scores = np.asarray(scores, dtype=np.float32)
    keep = scores > 0
    keys = (scores[keep].view(np.uint32).astype(np.uint64) << np.uint64(32)) | \
           (np.uint64(0xFFFFFFFF) - doc_ids[keep])
    keys = np.sort(keys)[::-1][:k]
    out = np.zeros(k, dtype=np.uint64)
    out[:len(keys)] = keys
    return out


def merge_topk(keys_by_rank, k):
    """[world, Q, k] per-shard keys -> [Q, k] global top-k keys (larger key = better)."""
    a = np.asarray(keys_by_rank, dtype=np.uint64)
    world, q, kk = a.shape
    flat = np.transpose(a, (1, 0, 2)).reshape(q, world * kk)
    return np.sort(flat, axis=1)[:, ::-1][:, :k]


def unpack_keys(keys):
    """keys -> (doc ids uint32 with 0xFFFFFFFF for empty slots, scores float32)."""
    keys = np.asarray(keys, dtype=np.uint64)
    docs = (np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))).astype(np.uint32)
    scores = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32)
    docs = np.where(keys == 0, np.uint32(0xFFFFFFFF), docs)
    return docs, scores


class ShardComm:
    """The few host-visible reductions a sharded query needs, over the NCCL communicator attached
    to one sa_index handle (sa_comm_init).  Everything exchanged is O(terms) or O(k) per query; the
    score vectors never leave their GPU."""

    def __init__(self, handle, rank, world):
        self.handle, self.rank, self.world = handle, rank, world

    def sum_u64(self, values):
        from . import _lib
        a = np.ascontiguousarray(values, dtype=np.uint64).copy()
        if self.world > 1 and a.size:
            _lib.check(_lib.lib().sa_comm_allreduce_sum_u64(self.handle, _lib.p_u64(a), a.size))
        return a

    def allgather_u64(self, values):
        """[n] per rank -> [world, n] on every rank (a sum over disjoint slots)."""
        v = np.ascontiguousarray(values, dtype=np.uint64)
        buf = np.zeros((self.world, v.size), dtype=np.uint64)
        buf[self.rank] = v
        return self.sum_u64(buf.reshape(-1)).reshape(self.world, v.size)

    def merge_topk_f64(self, docs, scores, k):
        """Per-shard (absolute doc ids, float64 scores) -> the global k best on every rank
        (score desc, doc asc), the single all-gather of per-shard top-k of SURVEY 8e."""
        d = self.allgather_u64(np.asarray(docs, dtype=np.uint64)).reshape(-1)
        s = self.allgather_u64(np.asarray(scores, dtype=np.float64).view(np.uint64)).reshape(-1).view(np.float64)
        keep = d != 0xFFFFFFFF
        d, s = d[keep], s[keep]
        order = np.lexsort((d, -s))[:k]
        out_d = np.full(k, 0xFFFFFFFF, dtype=np.uint32)
        out_s = np.zeros(k, dtype=np.float64)
        out_d[:len(order)] = d[order]
        out_s[:len(order)] = s[order]
        return out_d, out_s

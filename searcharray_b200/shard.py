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

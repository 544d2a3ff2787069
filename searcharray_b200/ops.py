"""Device implementations of the reference's native ops on raw host arrays (parity tests and
the Similarity protocol entry point).  Thin ctypes calls -- all arithmetic runs in CUDA."""
import ctypes

import numpy as np

from . import _lib


def bm25_score(term_freqs, doc_lens, avg_doc_lens, idf, k1, b, device=0):
    """searcharray.bm25.bm25_score (reference bm25/bm25.pyx:28-41): in place on term_freqs."""
    tf = term_freqs
    if tf.dtype != np.float32 or not tf.flags.c_contiguous:
        raise ValueError("term_freqs must be contiguous float32")
    dl = np.ascontiguousarray(doc_lens, dtype=np.float32)
    if len(dl) != len(tf):
        raise ValueError("doc_lens and term_freqs must have the same length")
    _lib.check(_lib.lib().sa_op_bm25_score(_lib.p_f32(tf), _lib.p_f32(dl), len(tf), float(avg_doc_lens),
                                           float(idf), float(k1), float(b), device))
    return tf


def popcount64_reduce(words, device=0):
    """reference roaringish/popcount.pyx:271-278 with key_shift=36, value_mask=0x3FFFF."""
    w = np.ascontiguousarray(words, dtype=np.uint64)
    if len(w) == 0:
        return np.array([]), np.array([])
    keys = np.empty(len(w), dtype=np.uint64)
    cnts = np.empty(len(w), dtype=np.float32)
    n = ctypes.c_uint64(0)
    _lib.check(_lib.lib().sa_op_popcount64_reduce(_lib.p_u64(w), len(w), device, _lib.p_u64(keys),
                                                  _lib.p_f32(cnts), ctypes.byref(n)))
    return keys[:n.value].copy(), cnts[:n.value].copy()


def bigram_freqs(lhs, rhs, cont_rhs=True, device=0):
    """reference phrase/bigram_freqs.py:213-307 -> ((doc ids, counts), continuation words)."""
    lhs = np.ascontiguousarray(lhs, dtype=np.uint64)
    rhs = np.ascontiguousarray(rhs, dtype=np.uint64)
    cap = 2 * min(len(lhs), len(rhs)) + 2
    ids = np.empty(cap, dtype=np.uint64)
    cnts = np.empty(cap, dtype=np.float32)
    nxt = np.empty(cap, dtype=np.uint64)
    n_ids, n_next = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _lib.check(_lib.lib().sa_op_bigram_freqs(_lib.p_u64(lhs), len(lhs), _lib.p_u64(rhs), len(rhs),
                                             1 if cont_rhs else 0, device, _lib.p_u64(ids), _lib.p_f32(cnts),
                                             ctypes.byref(n_ids), _lib.p_u64(nxt), ctypes.byref(n_next)))
    return (ids[:n_ids.value].copy(), cnts[:n_ids.value].copy()), nxt[:n_next.value].copy()

"""ctypes bindings for the CPU oracle's C kernels (oracle/sa_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(searcharray_b200) must never import this module.

Each wrapper mirrors the numpy-in / numpy-out signature of the Cython `def` it
restates (reference file:line is cited in sa_oracle.c).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsa_oracle.so")

ALL_BITS = 0xFFFFFFFFFFFFFFFF

_U64P = ctypes.POINTER(ctypes.c_uint64)
_F32P = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """Compile oracle/sa_oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "sa_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u64, f32 = ctypes.c_uint64, ctypes.c_float
        L.sao_bm25_score.argtypes = [_F32P, _F32P, f32, f32, f32, f32, ctypes.c_long]
        L.sao_bm25_score.restype = None
        L.sao_popcount64.argtypes = [_U64P, u64, _U64P]
        L.sao_popcount64.restype = None
        L.sao_popcount64_reduce.argtypes = [_U64P, u64, u64, u64, _U64P, _F32P]
        L.sao_popcount64_reduce.restype = u64
        L.sao_popcount_reduce_at.argtypes = [_U64P, _U64P, u64, _U64P, _F32P]
        L.sao_popcount_reduce_at.restype = u64
        L.sao_key_sum_over.argtypes = [_U64P, _U64P, u64, _U64P, _F32P]
        L.sao_key_sum_over.restype = u64
        L.sao_scatter.argtypes = [_F32P, _U64P, _F32P, u64]
        L.sao_scatter.restype = None
        L.sao_payload_slice.argtypes = [_U64P, u64, u64, u64, u64, _U64P]
        L.sao_payload_slice.restype = u64
        L.sao_unique.argtypes = [_U64P, u64, u64, _U64P]
        L.sao_unique.restype = u64
        L.sao_intersect_drop.argtypes = [_U64P, u64, _U64P, u64, u64, _U64P, _U64P]
        L.sao_intersect_drop.restype = u64
        L.sao_intersect_keep.argtypes = [_U64P, u64, _U64P, u64, u64, _U64P, _U64P, _U64P, _U64P]
        L.sao_intersect_keep.restype = None
        L.sao_adjacent.argtypes = [_U64P, u64, _U64P, u64, u64, _U64P, _U64P]
        L.sao_adjacent.restype = u64
        L.sao_intersect_with_adjacents.argtypes = [_U64P, u64, _U64P, u64, u64,
                                                   _U64P, _U64P, _U64P, _U64P, _U64P]
        L.sao_intersect_with_adjacents.restype = u64
        L.sao_merge.argtypes = [_U64P, u64, _U64P, u64, ctypes.c_int, _U64P]
        L.sao_merge.restype = u64
        L.sao_sort_merge_counts.argtypes = [_U64P, _F32P, u64, _U64P, _F32P, u64, _U64P, _F32P]
        L.sao_sort_merge_counts.restype = u64
        L.sao_span_freqs.argtypes = [_U64P, _U64P, u64, u64, u64, u64, u64, u64, _U64P, _F32P, _U64P]
        L.sao_span_freqs.restype = u64
        _lib = L
    return _lib


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p64(a):
    return a.ctypes.data_as(_U64P)


def _p32(a):
    return a.ctypes.data_as(_F32P)


# --------------------------------------------------------------------------- bm25
def bm25_score(term_freqs, doc_lens, avg_doc_lens, idf, k1, b):
    """In place, like searcharray.bm25.bm25_score (bm25.pyx:28-41)."""
    assert term_freqs.dtype == np.float32 and term_freqs.flags.c_contiguous
    dl = _f32(doc_lens)
    lib().sao_bm25_score(_p32(term_freqs), _p32(dl), float(avg_doc_lens), float(idf),
                         float(k1), float(b), term_freqs.shape[0])


# ----------------------------------------------------------------------- popcount
def popcount64(arr):
    arr = _u64(arr)
    out = np.empty(arr.shape[0], dtype=np.uint64)
    lib().sao_popcount64(_p64(arr), arr.shape[0], _p64(out))
    return out


def popcount64_reduce(arr, key_shift, value_mask):
    arr = _u64(arr)
    n = arr.shape[0]
    if n == 0:
        return np.array([]), np.array([])
    keys = np.empty(n, dtype=np.uint64)
    cnts = np.empty(n, dtype=np.float32)
    m = lib().sao_popcount64_reduce(_p64(arr), n, int(key_shift), int(value_mask), _p64(keys), _p32(cnts))
    return keys[:m].copy(), cnts[:m].copy()


def popcount_reduce_at(ids, payload):
    ids, payload = _u64(ids), _u64(payload)
    if len(ids) != len(payload):
        raise ValueError("ids and payload must have the same length")
    n = ids.shape[0]
    if n == 0:
        return np.array([]), np.array([])
    oi = np.empty(n, dtype=np.uint64)
    oc = np.empty(n, dtype=np.float32)
    m = lib().sao_popcount_reduce_at(_p64(ids), _p64(payload), n, _p64(oi), _p32(oc))
    return oi[:m].copy(), oc[:m].copy()


def key_sum_over(ids, count):
    ids, count = _u64(ids), _u64(count)
    if len(ids) != len(count):
        raise ValueError("ids and count must have the same length")
    n = ids.shape[0]
    if n == 0:
        return np.array([]), np.array([])
    oi = np.empty(n, dtype=np.uint64)
    oc = np.empty(n, dtype=np.float32)
    m = lib().sao_key_sum_over(_p64(ids), _p64(count), n, _p64(oi), _p32(oc))
    return oi[:m].copy(), oc[:m].copy()


# -------------------------------------------------------------------------- dense
def as_dense(indices, values, size):
    if len(indices) != len(values):
        raise ValueError("indices and values must have the same length")
    out = np.zeros(size, dtype=np.float32)
    indices, values = _u64(indices), _f32(values)
    if indices.shape[0]:
        lib().sao_scatter(_p32(out), _p64(indices), _p32(values), indices.shape[0])
    return out


def payload_slice(arr, payload_msb_mask, min_payload=0, max_payload=ALL_BITS):
    arr = _u64(arr)
    out = np.empty(arr.shape[0], dtype=np.uint64)
    m = lib().sao_payload_slice(_p64(arr), arr.shape[0], int(payload_msb_mask),
                                int(min_payload), int(max_payload), _p64(out))
    return out[:m].copy()


def unique(arr, rshift=0):
    arr = _u64(arr)
    out = np.empty(arr.shape[0], dtype=np.uint64)
    m = lib().sao_unique(_p64(arr), arr.shape[0], int(rshift), _p64(out))
    return out[:m].copy()


# ---------------------------------------------------------------------- intersect
def intersect(lhs, rhs, mask=ALL_BITS, drop_duplicates=True):
    if mask is None:
        mask = ALL_BITS
    if int(mask) == 0:
        raise ValueError("Mask cannot be zero")
    lhs, rhs = _u64(lhs), _u64(rhs)
    nl, nr = lhs.shape[0], rhs.shape[0]
    if drop_duplicates:
        cap = min(nl, nr)
        li = np.empty(cap, dtype=np.uint64)
        ri = np.empty(cap, dtype=np.uint64)
        m = lib().sao_intersect_drop(_p64(lhs), nl, _p64(rhs), nr, int(mask), _p64(li), _p64(ri))
        return li[:m], ri[:m]
    li = np.empty(nl, dtype=np.uint64)
    ri = np.empty(nr, dtype=np.uint64)
    ml, mr = ctypes.c_uint64(0), ctypes.c_uint64(0)
    lib().sao_intersect_keep(_p64(lhs), nl, _p64(rhs), nr, int(mask), _p64(li), _p64(ri),
                             ctypes.byref(ml), ctypes.byref(mr))
    return li[:ml.value], ri[:mr.value]


def adjacent(lhs, rhs, mask=ALL_BITS):
    if mask is None:
        mask = ALL_BITS
    if int(mask) == 0:
        raise ValueError("Mask cannot be zero")
    lhs, rhs = _u64(lhs), _u64(rhs)
    cap = min(lhs.shape[0], rhs.shape[0])
    li = np.empty(cap, dtype=np.uint64)
    ri = np.empty(cap, dtype=np.uint64)
    m = lib().sao_adjacent(_p64(lhs), lhs.shape[0], _p64(rhs), rhs.shape[0], int(mask), _p64(li), _p64(ri))
    return li[:m], ri[:m]


def intersect_with_adjacents(lhs, rhs, mask=ALL_BITS):
    if mask is None:
        mask = ALL_BITS
    if int(mask) == 0:
        raise ValueError("Mask cannot be zero")
    lhs, rhs = _u64(lhs), _u64(rhs)
    cap = min(lhs.shape[0], rhs.shape[0])
    li = np.empty(cap, dtype=np.uint64)
    ri = np.empty(cap, dtype=np.uint64)
    # an lhs element may pair once as "equal" and once as "adjacent": cap is enough for each list
    ali = np.empty(cap, dtype=np.uint64)
    ari = np.empty(cap, dtype=np.uint64)
    na = ctypes.c_uint64(0)
    m = lib().sao_intersect_with_adjacents(_p64(lhs), lhs.shape[0], _p64(rhs), rhs.shape[0], int(mask),
                                           _p64(li), _p64(ri), _p64(ali), _p64(ari), ctypes.byref(na))
    return li[:m], ri[:m], ali[:na.value], ari[:na.value]


# -------------------------------------------------------------------------- merge
def merge(lhs, rhs, drop_duplicates=False):
    lhs, rhs = _u64(lhs), _u64(rhs)
    out = np.empty(lhs.shape[0] + rhs.shape[0], dtype=np.uint64)
    m = lib().sao_merge(_p64(lhs), lhs.shape[0], _p64(rhs), rhs.shape[0], int(bool(drop_duplicates)), _p64(out))
    return out[:m].copy()


def sort_merge_counts(lhs_ids, lhs_counts, rhs_ids, rhs_counts):
    lhs_ids, rhs_ids = _u64(lhs_ids), _u64(rhs_ids)
    lhs_counts, rhs_counts = _f32(lhs_counts), _f32(rhs_counts)
    n = lhs_ids.shape[0] + rhs_ids.shape[0]
    oi = np.empty(n, dtype=np.uint64)
    oc = np.empty(n, dtype=np.float32)
    m = lib().sao_sort_merge_counts(_p64(lhs_ids), _p32(lhs_counts), lhs_ids.shape[0],
                                    _p64(rhs_ids), _p32(rhs_counts), rhs_ids.shape[0], _p64(oi), _p32(oc))
    return oi[:m].copy(), oc[:m].copy()


# -------------------------------------------------------------------------- spans
last_span_undefined = 0


def span_search(posns, lengths, slop, key_mask, header_mask, key_bits, lsb_bits):
    """Returns (keys u64[], counts f32[]) in Counter insertion order
    (roaringish/spans.pyx:322-330 + phrase/spans.py:186-187).  Sets the module global
    `last_span_undefined` to the number of table overflows the reference leaves undefined."""
    global last_span_undefined
    posns = _u64(posns)
    lengths = _u64(lengths)
    num_terms = lengths.shape[0] - 1
    padded = np.concatenate([posns, np.zeros(1, dtype=np.uint64)])
    n_docs_max = int(lengths[1] - lengths[0]) + 1
    keys = np.empty(n_docs_max, dtype=np.uint64)
    cnts = np.empty(n_docs_max, dtype=np.float32)
    undef = ctypes.c_uint64(0)
    m = lib().sao_span_freqs(_p64(padded), _p64(lengths), num_terms, int(slop), int(key_mask),
                             int(header_mask), int(key_bits), int(lsb_bits), _p64(keys), _p32(cnts),
                             ctypes.byref(undef))
    last_span_undefined = undef.value
    return keys[:m].copy(), cnts[:m].copy()

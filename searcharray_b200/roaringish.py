"""Roaringish posting-word layout and the index-time (host, numpy) encoder.

Layout (reference searcharray/roaringish/roaringish.py:30-35): one uint64 per
(doc, block = posn // 18):   doc id (28 b) | block (18 b) | bitmap of posn % 18 (18 b).
Index build stays on the host (BASELINE north_star: "pandas/numpy for the column type and
indexing"); the query-time operators on these words run on the GPU.
"""
import numpy as np

KEY_BITS = 28
LSB_BITS = 18
KEY_SHIFT = 36
MAX_POSN = (1 << 18) - 1          # reference roaringish.py:86
MAX_DOCS = 1 << KEY_BITS

_LSB = np.uint64(LSB_BITS)
_SHIFT = np.uint64(KEY_SHIFT)
_ONE = np.uint64(1)


def encode_postings(doc_ids, posns):
    """(doc, posn) pairs of ONE term, sorted by doc then posn -> sorted header-unique words.

    Same result as RoaringishEncoder.encode(payload=posns, keys=doc_ids)
    (reference roaringish.py:93-142)."""
    doc_ids = np.asarray(doc_ids).astype(np.uint64, copy=False)
    posns = np.asarray(posns).astype(np.uint64, copy=False)
    if posns.size == 0:
        return np.empty(0, dtype=np.uint64)
    if posns.max() > MAX_POSN:
        raise ValueError(f"Positions must be less than {1 << LSB_BITS}")
    header = (doc_ids << _SHIFT) | ((posns // _LSB) << _LSB)
    bit = _ONE << (posns % _LSB)
    first = np.flatnonzero(np.concatenate(([True], header[1:] != header[:-1])))
    payload = np.bitwise_or.reduceat(bit, first)
    return header[first] | payload


def encode_grouped(term_ids, doc_ids, posns):
    """Encodes many terms at once.  Input triples sorted by (term, doc, posn).
    Returns (words, term_ids_unique, offsets, lengths)."""
    term_ids = np.asarray(term_ids)
    if term_ids.size == 0:
        e = np.empty(0, dtype=np.uint64)
        return e, np.empty(0, dtype=np.int64), e.copy(), e.copy()
    doc_ids = np.asarray(doc_ids).astype(np.uint64, copy=False)
    posns = np.asarray(posns).astype(np.uint64, copy=False)
    if posns.max() > MAX_POSN:
        raise ValueError(f"Positions must be less than {1 << LSB_BITS}")
    header = (doc_ids << _SHIFT) | ((posns // _LSB) << _LSB)
    bit = _ONE << (posns % _LSB)
    new_word = np.concatenate(([True], (header[1:] != header[:-1]) | (term_ids[1:] != term_ids[:-1])))
    first = np.flatnonzero(new_word)
    words = header[first] | np.bitwise_or.reduceat(bit, first)
    word_terms = term_ids[first]
    tstart = np.flatnonzero(np.concatenate(([True], word_terms[1:] != word_terms[:-1])))
    uniq = word_terms[tstart]
    offsets = tstart.astype(np.uint64)
    lengths = np.diff(np.concatenate((tstart, [len(words)]))).astype(np.uint64)
    return words, uniq, offsets, lengths


def decode_positions(words):
    """word list of one (term, doc) -> sorted positions (reference roaringish.py:144-166)."""
    words = np.asarray(words, dtype=np.uint64)
    out = []
    for w in words:
        w = int(w)
        base = ((w >> LSB_BITS) & 0x3FFFF) * LSB_BITS
        bits = w & 0x3FFFF
        while bits:
            low = bits & -bits
            out.append(base + low.bit_length() - 1)
            bits ^= low
    return np.asarray(out, dtype=np.uint32)

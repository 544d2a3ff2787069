"""CPU oracle for SearchArray's scoring hot path -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the Python layer of the reference hot path, on top of the C
restatement of its Cython kernels (oracle/ops.py -> oracle/sa_oracle.c).  Paths are
relative to /root/reference.  Nothing under searcharray_b200/ may import this.

Pinned against the real reference by tests/golden/make_golden.py (golden vectors in
tests/golden/*.npz are produced by importing the reference itself).

Bit layout (roaringish/roaringish.py:30-35,66-86): word = doc(28b) | block(18b) | bitmap(18b).
"""
import numpy as np

from . import ops

U = np.uint64
_1 = U(1)
KEY_BITS = U(28)
PAYLOAD_LSB_BITS = U(18)
KEY_SHIFT = U(36)
KEY_MASK = U(0xFFFFFFF000000000)
PAYLOAD_MSB_MASK = U(0x0000000FFFFC0000)
PAYLOAD_LSB_MASK = U(0x000000000003FFFF)
HEADER_MASK = U(0xFFFFFFFFFFFC0000)
HEADER_BITS = U(46)
UPPER_BIT = U(1 << 17)
MAX_POSN = (1 << 18) - 1

RHS, LHS = "rhs", "lhs"


# ------------------------------------------------------------------ encoder bits
def encode(doc_ids, posns):
    """roaringish/roaringish.py:93-142 (encode, no boundaries): (doc, posn) pairs sorted
    by doc then posn -> one word per (doc, posn // 18)."""
    doc_ids = np.asarray(doc_ids, dtype=np.uint64)
    posns = np.asarray(posns, dtype=np.uint64)
    if len(posns) == 0:
        return np.array([], dtype=np.uint64)
    hdr = (doc_ids << KEY_SHIFT) | ((posns // PAYLOAD_LSB_BITS) << PAYLOAD_LSB_BITS)
    words = hdr | (_1 << (posns % PAYLOAD_LSB_BITS))
    starts = np.concatenate([[0], np.nonzero(np.diff(hdr))[0] + 1])
    return np.bitwise_or.reduceat(words.view(np.int64), starts).view(np.uint64)


def slice_words(words, keys=None, header=None, min_payload=None, max_payload=None):
    """roaringish/roaringish.py:245-282 (RoaringishEncoder.slice)."""
    if header is not None:
        if keys is not None:
            raise ValueError("Can't specify both keys and header")
        _, idx = ops.intersect(header, words & ~PAYLOAD_LSB_MASK, drop_duplicates=False)
        words = words[idx]
    if keys is not None:
        _, idx = ops.intersect(np.asarray(keys).view(np.uint64), words >> KEY_SHIFT, drop_duplicates=False)
        words = words[idx]
    if max_payload is None and min_payload is None:
        return words
    if min_payload is not None and min_payload % 18 != 0:
        raise ValueError("min_payload must be a multiple of 18")
    if max_payload is not None and max_payload % 18 != 17:
        raise ValueError("max_payload must be a multiple of 18 - 1")
    lo = 0 if min_payload is None else min_payload
    hi = 0xFFFFFFFFFFFFFFFF if max_payload is None else max_payload
    return ops.payload_slice(words, PAYLOAD_MSB_MASK, lo // 18, hi // 18)


# --------------------------------------------------------------------- term path
def docfreq(words):
    """phrase/middle_out.py:521-528 + roaringish.py:176-179 + unique.pyx:87-104."""
    return int(ops.unique(words, KEY_SHIFT).size)


def termfreqs_sparse(words):
    """phrase/middle_out.py:498-499 + roaringish.py:168-170 (num_values_per_key)."""
    return ops.popcount64_reduce(words, KEY_SHIFT, PAYLOAD_LSB_MASK)


def termfreqs_dense(words, num_docs):
    """postings.py:629-636: sparse (doc, tf) -> dense float32[num_docs]."""
    ids, tfs = termfreqs_sparse(words)
    if len(ids) == 0:
        return np.zeros(num_docs, dtype=np.float32)
    return ops.as_dense(ids, tfs, num_docs)


def compute_idf(num_docs, dfs):
    """similarity.py:19-21."""
    dfs = np.asarray(dfs)
    return np.sum(np.log(1 + (num_docs - dfs + 0.5) / (dfs + 0.5)))


def bm25(tfs, dfs, doc_lens, avg_doc_len, num_docs, k1=1.2, b=0.75):
    """similarity.py:24-38 (bm25_similarity closure): mutates and returns tfs."""
    if avg_doc_len == 0:
        return np.zeros_like(tfs)
    idf = compute_idf(num_docs, dfs)
    ops.bm25_score(tfs, doc_lens, avg_doc_len, idf, k1, b)
    return tfs


# ------------------------------------------------------------------- phrase path
def _adj_to_phrase_freq(overlap, adjacents):
    """phrase/bigram_freqs.py:48-62: a run of c+1 equal terms holds ceil(c/2)... pairs."""
    runs = ops.popcount64((overlap & (overlap << _1)) & PAYLOAD_LSB_MASK)
    adjacents -= -np.floor_divide(runs, -2, dtype=np.int64)
    return adjacents


def _inner_same_term(lhs_int, rhs_int, lhs_docs, cont):
    """phrase/bigram_freqs.py:65-101 (_inner_bigram_same_term), incl. the `lhs >> 1`
    header-bit leak into bit 17 of the LHS continuation (SURVEY quirk v)."""
    rhs_shift = rhs_int << _1
    overlap = lhs_int & rhs_shift
    adjacents = ops.popcount64(overlap & PAYLOAD_LSB_MASK).view(np.int64)
    adjusted = _adj_to_phrase_freq(overlap, adjacents).astype(np.uint64)
    ids, freqs = ops.key_sum_over(lhs_docs, adjusted)
    hdr = lhs_int & ~PAYLOAD_LSB_MASK
    rhs_cont = ((rhs_shift & rhs_int) & PAYLOAD_LSB_MASK) | hdr if cont == RHS else None
    lhs_cont = hdr | ((lhs_int & (lhs_int >> _1)) & PAYLOAD_LSB_MASK) if cont == LHS else None
    return (ids, freqs), (lhs_cont, rhs_cont)


def _inner(lhs_int, rhs_int, cont):
    """phrase/bigram_freqs.py:104-155 (_inner_bigram_freqs)."""
    lhs_docs = lhs_int >> KEY_SHIFT
    if len(lhs_int) == 0:
        empty = (np.array([], dtype=np.uint64), np.array([]))
        return (empty, (None, rhs_int)) if cont == RHS else (empty, (lhs_int, None))
    if len(lhs_int) == len(rhs_int) and np.all(lhs_int == rhs_int):
        return _inner_same_term(lhs_int, rhs_int, lhs_docs, cont)
    overlap = (lhs_int & PAYLOAD_LSB_MASK) & ((rhs_int & PAYLOAD_LSB_MASK) >> _1)
    lhs_next = rhs_next = None
    if cont == RHS:
        rhs_next = ((overlap << _1) & PAYLOAD_LSB_MASK) | (rhs_int & HEADER_MASK)
    else:
        lhs_next = overlap | (lhs_int & HEADER_MASK)
    ids, counts = ops.popcount_reduce_at(lhs_docs, overlap)
    return (ids, counts), (lhs_next, rhs_next)


def _adjacent(lhs_adj, rhs_adj, cont):
    """phrase/bigram_freqs.py:158-188 (_adjacent_bigram_freqs): lhs bit 17 & rhs bit 0
    on words whose headers differ by one block."""
    lhs_docs = lhs_adj >> KEY_SHIFT
    hit = ((lhs_adj & UPPER_BIT) != 0) & ((rhs_adj & _1) != 0)
    ids, counts = np.unique(lhs_docs[hit], return_counts=True)
    empty = np.asarray([], dtype=np.uint64)
    rhs_next = None if cont == LHS else empty
    lhs_next = None if cont == RHS else empty
    if np.any(hit):
        if cont == RHS:
            rhs_next = (rhs_adj[hit] & ~PAYLOAD_LSB_MASK) | _1
        else:
            lhs_next = (lhs_adj[hit] & ~PAYLOAD_LSB_MASK) | UPPER_BIT
    return (ids, counts), (lhs_next, rhs_next)


def _set_adjbit_at_header(next_inner, next_adj, cont):
    """phrase/bigram_freqs.py:191-210."""
    if len(next_inner) == 0:
        return next_adj
    if len(next_adj) == 0:
        return next_inner
    same_inner, same_adj = ops.intersect(next_inner, next_adj, mask=HEADER_MASK)
    keep = np.ones(len(next_adj), dtype=bool)
    keep[same_adj] = False
    if len(same_inner) > 0:
        next_inner[same_inner] |= (_1 if cont == RHS else UPPER_BIT)
        next_adj = next_adj[keep]
    return ops.merge(next_inner, next_adj)


def bigram_freqs(lhs, rhs, cont=RHS):
    """phrase/bigram_freqs.py:213-307: ((doc ids, counts f32), continuation words)."""
    li, ri, lai, rai = ops.intersect_with_adjacents(lhs, rhs, mask=HEADER_MASK)
    (ids, counts), (lhs_in, rhs_in) = _inner(lhs[li], rhs[ri], cont)
    (aids, acounts), (lhs_adj, rhs_adj) = _adjacent(lhs[lai], rhs[rai], cont)
    docs, cnts = ops.sort_merge_counts(ids, counts.astype(np.float32), aids, acounts.astype(np.float32))
    if cont == RHS:
        nxt = _set_adjbit_at_header(rhs_in, rhs_adj, RHS)
    else:
        nxt = _set_adjbit_at_header(lhs_in, lhs_adj, LHS)
    return (docs, cnts), nxt


def _and_min(ids, counts, new_ids, new_counts):
    """phrase/middle_out.py:73-93 (_intersect_bigram_matches)."""
    if ids is None:
        return new_ids, new_counts
    a, bidx = ops.intersect(ids, new_ids)
    return ids[a], np.minimum(counts[a], new_counts[bidx])


def _chain(enc, direction):
    """phrase/middle_out.py:96-151: left-to-right (cont=RHS) or right-to-left (cont=LHS)."""
    if len(enc) < 2:
        raise ValueError("phrase must have at least two terms")
    ids = counts = None
    if direction == RHS:
        carry = enc[0]
        for rhs in enc[1:]:
            (d, c), carry = bigram_freqs(carry, rhs, cont=RHS)
            ids, counts = _and_min(ids, counts, d, c)
    else:
        carry = enc[-1]
        for lhs in enc[-2::-1]:
            (d, c), carry = bigram_freqs(lhs, carry, cont=LHS)
            ids, counts = _and_min(ids, counts, d, c)
    return ids, counts


def compute_phrase_freqs(enc):
    """phrase/middle_out.py:154-168: direction chosen by where the shortest list sits;
    the middle-out split does NOT check adjacency across the split (reference quirk)."""
    shortest = min(range(len(enc)), key=lambda i: len(enc[i]))
    if shortest <= 1:
        return _chain(enc, RHS)
    if shortest >= len(enc) - 2:
        return _chain(enc, LHS)
    l_ids, l_counts = _chain(enc[:shortest], RHS)
    r_ids, r_counts = _chain(enc[shortest:], LHS)
    return _and_min(l_ids, l_counts, r_ids, r_counts)


# --------------------------------------------------------------------- slop path
def _span_candidates(enc):
    """phrase/spans.py:71-123 (_intersect_all): headers where every term occurs within
    +-1 block of term 0, then slice every term to {h-1, h, h+1}."""
    if len(enc) < 2:
        raise ValueError("Need at least two positions to intersect")
    last_l = last_r = None
    curr = enc[0]
    for nxt in enc[1:]:
        li, _ = ops.intersect(curr, nxt, mask=HEADER_MASK)
        int_hdr = curr[li] & ~PAYLOAD_LSB_MASK
        c2r, n2l = ops.adjacent(curr, nxt, mask=HEADER_MASK)
        lh = ops.merge(int_hdr, nxt[n2l])
        rh = ops.merge(int_hdr, curr[c2r])
        n2r, c2l = ops.adjacent(nxt, curr, mask=HEADER_MASK)
        lh = ops.merge(lh, curr[c2l])
        rh = ops.merge(rh, nxt[n2r])
        if last_l is not None:
            a, _ = ops.intersect(last_l, lh, mask=HEADER_MASK)
            b, _ = ops.intersect(last_r, rh, mask=HEADER_MASK)
            last_l, last_r = last_l[a], last_r[b]
        else:
            last_l, last_r = lh, rh
    one_block = _1 << (U(64) - HEADER_BITS)
    allh = ops.merge(last_r + one_block, last_l - one_block, drop_duplicates=True)
    allh = ops.merge(last_l, allh, drop_duplicates=True)
    allh = ops.merge(last_r, allh, drop_duplicates=True)
    allh = allh & HEADER_MASK
    sliced = [slice_words(e, header=allh) for e in enc]
    lengths = np.cumsum([0] + [len(s) for s in sliced], dtype=np.uint64)
    return np.concatenate(sliced, dtype=np.uint64), lengths


def span_search(enc, slop):
    """phrase/spans.py:171-187."""
    posns, lengths = _span_candidates(enc)
    return ops.span_search(posns, lengths, slop, KEY_MASK, HEADER_MASK, KEY_BITS, PAYLOAD_LSB_BITS)


# ------------------------------------------------------------------------ facade
class OracleIndex:
    """The slice of SearchArray / PosnBitArray state the hot path reads
    (postings.py:293-299, phrase/middle_out.py:320-328): per-term sorted words, doc_lens,
    avg_doc_length, corpus_size.  `rows` models a sliced array (FilteredPosns,
    middle_out.py:291-317; postings.py:344-358)."""

    def __init__(self, term_words, doc_lens, avg_doc_length=None, rows=None, corpus_size=None,
                 max_doc_id=None, cache=False):
        # cache=True reproduces PosnBitArray's docfreq_cache / termfreq_cache
        # (phrase/middle_out.py:326-327,501-528): the reference's steady ("warm") state.
        self.cache = cache
        self._df_cache = {}
        self._tf_cache = {}
        self.term_words = term_words          # {term_id: np.uint64[]}
        self.doc_lens = np.asarray(doc_lens, dtype=np.float32)
        self.rows = None if rows is None else np.asarray(rows, dtype=np.uint64)
        self.corpus_size = len(self.doc_lens) if corpus_size is None else corpus_size
        self.max_doc_id = (self.corpus_size - 1) if max_doc_id is None else max_doc_id
        self.avg_doc_length = (np.mean(self.doc_lens) if len(self.doc_lens) else 0) \
            if avg_doc_length is None else avg_doc_length

    def __len__(self):
        return len(self.doc_lens) if self.rows is None else len(self.rows)

    def sliced(self, key):
        """postings.py:344-358 (__getitem__ with a slice / mask / index array).

        Reference quirk (ix): `arr.doc_lens = self.doc_lens[key]` is a STRIDED VIEW for a
        stepped slice, and bm25_score (bm25.pyx:34-41) walks `&doc_lens[0]` contiguously,
        ignoring the stride -- so BM25 on arr[1::2] reads doc_lens[1], [2], [3]... of the
        parent.  Reproduced here (`_bm25_doc_lens`) for positive steps."""
        n = len(self)
        rows_local = np.arange(n)[key]
        doc_lens = self.doc_lens[key]
        base_rows = rows_local.astype(np.uint64) if self.rows is None else self.rows[rows_local]
        o = OracleIndex(self.term_words, doc_lens, self.avg_doc_length,
                        rows=base_rows, corpus_size=self.corpus_size, max_doc_id=self.max_doc_id)
        if isinstance(key, slice) and key.step not in (None, 1) and len(rows_local):
            first = int(rows_local[0])
            o._bm25_doc_lens = self.doc_lens[first:first + len(rows_local)]
            assert len(o._bm25_doc_lens) == len(rows_local), "reference would read out of bounds"
        return o

    def _words(self, term_id):
        w = self.term_words[term_id]
        if self.rows is not None:
            w = slice_words(w, keys=self.rows)      # FilteredPosns.__getitem__
        return w

    def docfreq(self, term_id):
        """postings.py:640-647; on a slice df comes from the filtered postings (quirk iii)."""
        if term_id is None or term_id not in self.term_words:
            return 0
        if self.cache and self.rows is None:
            if term_id not in self._df_cache:
                df = docfreq(self.term_words[term_id])
                if len(self.term_words[term_id]) <= 25:      # cache_gt_than (middle_out.py:517-519)
                    return df
                self._df_cache[term_id] = df
            return self._df_cache[term_id]
        return docfreq(self._words(term_id))

    def termfreqs(self, term_ids, slop=0, min_posn=None, max_posn=None):
        """postings.py:607-638, 689-708."""
        n = len(self)
        if isinstance(term_ids, (list, tuple)) and len(term_ids) == 1:
            term_ids = term_ids[0]
        if isinstance(term_ids, (list, tuple)):
            return self._phrase_freq(list(term_ids), slop, min_posn, max_posn)
        if term_ids is None or term_ids not in self.term_words:
            return np.zeros(n, dtype=np.float32)
        if self.rows is not None:
            w = slice_words(self._words(term_ids), keys=self.rows, min_payload=min_posn, max_payload=max_posn)
            ids, tfs = termfreqs_sparse(w)
            out = np.zeros(n, dtype=np.float32)
            out[np.isin(self.rows, ids)] = tfs
            return out
        w = self.term_words[term_ids]
        if min_posn is not None or max_posn is not None:
            w = slice_words(w, min_payload=min_posn, max_payload=max_posn)
        elif self.cache:
            # _termfreqs_with_cache (middle_out.py:501-509): cached iff the df is cached
            if term_ids not in self._tf_cache:
                ids, tfs = termfreqs_sparse(w)
                if term_ids not in self._df_cache:
                    return ops.as_dense(ids, tfs, n) if len(ids) else np.zeros(n, dtype=np.float32)
                self._tf_cache[term_ids] = (ids, tfs)
            ids, tfs = self._tf_cache[term_ids]
            return ops.as_dense(ids, tfs, n) if len(ids) else np.zeros(n, dtype=np.float32)
        return termfreqs_dense(w, n)

    def _phrase_freq(self, term_ids, slop, min_posn, max_posn):
        """postings.py:689-708 + phrase/middle_out.py:418-446."""
        n = len(self)
        if any(t is None or t not in self.term_words for t in term_ids):
            return np.zeros(n, dtype=np.float32)
        if len(term_ids) < 2:
            raise ValueError("Must have at least two terms")
        buf = np.zeros(int(self.max_doc_id) + 1, dtype=np.float32)
        if min_posn is None and max_posn is None:
            enc = [self._words(t) for t in term_ids]
        else:
            enc = [slice_words(self._words(t), min_payload=min_posn, max_payload=max_posn) for t in term_ids]
        if slop == 0:
            ids, counts = compute_phrase_freqs(enc)
        else:
            ids, counts = span_search(enc, slop)
        if ids is not None and len(ids):
            buf[ids.astype(np.int64)] = counts
        if self.rows is not None:
            return buf[self.rows.astype(np.int64)]
        return buf

    def score(self, term_ids, k1=1.2, b=0.75, slop=0, min_posn=None, max_posn=None):
        """postings.py:652-680 with the default bm25_similarity (similarity.py:24-38)."""
        toks = term_ids if isinstance(term_ids, (list, tuple)) else [term_ids]
        dfs = np.asarray([self.docfreq(t) for t in toks])
        tfs = self.termfreqs(term_ids, slop=slop, min_posn=min_posn, max_posn=max_posn)
        doc_lens = getattr(self, "_bm25_doc_lens", self.doc_lens)
        return bm25(tfs, dfs, doc_lens, self.avg_doc_length, self.corpus_size, k1, b)

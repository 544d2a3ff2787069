"""SearchArray -- the reference's search surface, backed by the B200 kernels.

Mirrors the Search API of reference searcharray/postings.py (`SearchArray.index` :249-300,
`termfreqs` :607-638, `docfreq` :640-647, `doclengths` :649-650, `score` :652-680,
`positions` :682-687, `_phrase_freq` :689-708): same names, argument meaning, return types
(dense float32[len(self)] numpy vectors) and error behaviour (TypeError / ValueError, unknown
terms -> zeros).  The postings live in HBM (one `sa_index` handle per array); every
`.score/.termfreqs` is one C-ABI call into libsearcharray_b200.so.  There is no CPU path.

Host-side pandas plumbing beyond what a DataFrame column needs is out of scope
(SURVEY.md section 2 row 2).
"""
import ctypes
import numbers
import threading
import weakref
from typing import List, Optional, Union

import numpy as np
import pandas as pd
from pandas.api.extensions import ExtensionArray, ExtensionDtype, register_extension_dtype
from pandas.api.types import is_list_like

from . import _lib
from .indexing import HostIndex, TermMissingError, build_index
from .roaringish import decode_positions
from .similarity import Bm25Similarity, Similarity, compute_idf, default_bm25


def ws_tokenizer(string):
    """reference postings.py:206-211"""
    if pd.isna(string):
        return []
    if not isinstance(string, str):
        raise ValueError("Expected a string")
    return string.split()


class Terms:
    """One indexed doc: term -> positions (a light stand-in for reference postings.py:57-160)."""

    def __init__(self, postings, doc_len=0):
        self.postings = postings
        self.doc_len = doc_len

    def terms(self):
        return ((t, len(p)) for t, p in self.postings.items())

    def termfreq(self, token):
        return len(self.postings.get(token, ()))

    def positions(self, token=None):
        if token is None:
            return self.postings
        return self.postings.get(token)

    def __len__(self):
        return len(self.postings)

    def __eq__(self, other):
        return isinstance(other, Terms) and self.doc_len == other.doc_len and \
            {k: list(v) for k, v in self.postings.items()} == {k: list(v) for k, v in other.postings.items()}

    def __hash__(self):
        return hash((self.doc_len, tuple(sorted(self.postings))))

    def __repr__(self):
        return f"Terms({ {k: list(map(int, v)) for k, v in self.postings.items()} })"


@register_extension_dtype
class TermsDtype(ExtensionDtype):
    """reference postings.py:163-203"""
    name = "tokenized_text_b200"
    type = Terms
    kind = "O"

    @classmethod
    def construct_from_string(cls, string):
        if not isinstance(string, str):
            raise TypeError("'construct_from_string' expects a string, got {}".format(type(string)))
        if string == cls.name:
            return cls()
        raise TypeError(f"Cannot construct a '{cls.__name__}' from '{string}'")

    @classmethod
    def construct_array_type(cls):
        return SearchArray

    @property
    def na_value(self):
        return Terms({})

    def __repr__(self):
        return "TermsDtype()"


class _PinnedPool:
    """Result vectors live in pinned host memory so the float32[N] D2H copy runs at PCIe
    rate; buffers are recycled when the numpy array that views them is garbage-collected.
    Sizes are bucketed to powers of two (so differently sized slices share buffers) and the idle
    pool is capped: beyond the cap a released buffer goes back to the driver (sa_host_free)."""

    MAX_IDLE_BYTES = 1 << 30

    def __init__(self):
        self._free = {}
        self._idle = 0
        self._lock = threading.Lock()

    @staticmethod
    def _bucket(nbytes):
        b = 4096
        while b < nbytes:
            b <<= 1
        return b

    def empty_f32(self, n):
        n = int(n)
        nbytes = self._bucket(max(n, 1) * 4)
        with self._lock:
            lst = self._free.get(nbytes)
            ptr = lst.pop() if lst else None
            if ptr is not None:
                self._idle -= nbytes
        if ptr is None:
            p = ctypes.c_void_p()
            _lib.check(_lib.lib().sa_host_alloc(ctypes.byref(p), nbytes))
            ptr = p.value
        buf = (ctypes.c_float * max(n, 1)).from_address(ptr)
        weakref.finalize(buf, self._release, nbytes, ptr)
        return np.frombuffer(buf, dtype=np.float32, count=n)

    def _release(self, nbytes, ptr):
        with self._lock:
            if self._idle + nbytes <= self.MAX_IDLE_BYTES:
                self._free.setdefault(nbytes, []).append(ptr)
                self._idle += nbytes
                return
        if _lib._lib is not None:
            _lib._lib.sa_host_free(ctypes.c_void_p(ptr))


_pool = _PinnedPool()


class DeviceIndex:
    """Owns one sa_index handle (one shard in one GPU's HBM)."""

    def __init__(self, host: HostIndex, device=0, doc_base=0):
        self.handle = ctypes.c_void_p()
        self.n_docs = host.n_docs
        self._rows_set = False
        words = host.words
        words_ptr = ctypes.cast(words.ctypes.data, _lib.P_u64) if len(words) else None   # (also a read-only memmap)
        _lib.check(_lib.lib().sa_index_create(
            words_ptr, len(words), _lib.p_u64(host.term_offsets),
            _lib.p_u64(host.term_lengths), host.n_terms, _lib.p_f32(host.doc_lens), host.n_docs,
            doc_base, device, ctypes.byref(self.handle)))
        self._finalizer = weakref.finalize(self, DeviceIndex._destroy, self.handle)

    @staticmethod
    def _destroy(handle):
        if handle and _lib._lib is not None:
            _lib._lib.sa_index_destroy(handle)

    def close(self):
        self._finalizer()


class SearchArray(ExtensionArray):
    """An ExtensionArray of tokenised text searchable with term / phrase queries on the GPU."""

    dtype = TermsDtype()

    def __init__(self, postings, tokenizer=ws_tokenizer, avoid_copies=True, device=0):
        if not is_list_like(postings):
            raise TypeError("Expected list-like object, got {}".format(type(postings)))
        self.tokenizer = tokenizer
        self.avoid_copies = avoid_copies
        self.device = device
        docs = []
        for p in postings:
            if isinstance(p, Terms):
                toks = [None] * int(p.doc_len)
                for t, posns in p.postings.items():
                    for x in posns:
                        if x >= len(toks):
                            toks.extend([None] * (x + 1 - len(toks)))
                        toks[x] = t
                docs.append(toks)
            elif isinstance(p, str) or p is None or (isinstance(p, float) and np.isnan(p)):
                docs.append(tokenizer(p))
            else:
                raise TypeError("Expected a Terms or a string")
        self._set_host(build_index(docs, lambda toks: [t for t in toks if t is not None]))

    # ------------------------------------------------------------------ construction
    def _set_host(self, host: HostIndex):
        self.host = host
        self.term_dict = host.term_dict
        self.doc_lens = host.doc_lens
        self.avg_doc_length = host.avg_doc_length
        self.corpus_size = host.n_docs
        self.rows = None                 # sliced view: local doc ids (postings.py:344-358)
        self._bm25_doc_lens = None
        # doc-range shard of a larger corpus: absolute id of row 0, and the GLOBAL
        # statistics idf / BM25 must use (set by from_host_index; None = this array is the corpus)
        self.doc_base = 0
        self.global_df = None            # uint64[n_terms] document frequencies over all shards
        self.comm = None                 # shard.ShardComm: sums over ranks where a query needs them
        self._shared = {"dev": None, "lock": threading.RLock()}   # shared by views/copies

    @classmethod
    def index(cls, array, tokenizer=ws_tokenizer, truncate=False, batch_size=100000, avoid_copies=True,
              workers=4, cache_gt_than=25, data_dir: Optional[str] = None, autowarm=True,
              device=0, gpu_build=False) -> "SearchArray":
        """Index an array of strings (reference postings.py:249-300).  batch_size / workers /
        cache_gt_than / data_dir / autowarm are accepted for signature compatibility: the
        per-term df table the reference warms lazily is computed on the device at upload.  data_dir:
        the posting words are written to `<data_dir>/<n>.dat` and memory-mapped (reference MemoryMappedArrays); the
        upload then DMAs straight from the mapping (cudaHostRegister), and a pickle carries the file name only."""
        if not is_list_like(array):
            raise TypeError("Expected list-like object, got {}".format(type(array)))
        host = build_index(list(array), tokenizer, truncate=truncate, gpu_build=device if gpu_build else None)
        if data_dir is not None:            # reference indexing.py:228-230, 291-293: memmap the bit positions
            host.memmap(data_dir)
        return cls.from_host_index(host, tokenizer=tokenizer, avoid_copies=avoid_copies, device=device)

    @classmethod
    def from_host_index(cls, host: HostIndex, tokenizer=ws_tokenizer, avoid_copies=True, device=0,
                        doc_base=0, corpus_size=None, avg_doc_length=None, global_df=None, comm=None):
        """Wraps a prebuilt HostIndex.  For one doc-range shard of a larger corpus pass the shard's
        first absolute doc id and the global corpus size / average doc length / per-term document
        frequencies (idf must not depend on the sharding)."""
        obj = cls.__new__(cls)
        obj.tokenizer = tokenizer
        obj.avoid_copies = avoid_copies
        obj.device = device
        obj._set_host(host)
        obj.doc_base = int(doc_base)
        if corpus_size is not None:
            obj.corpus_size = int(corpus_size)
        if avg_doc_length is not None:
            obj.avg_doc_length = avg_doc_length
        obj.global_df = None if global_df is None else np.asarray(global_df, dtype=np.uint64)
        obj.comm = comm
        return obj

    def _device(self) -> DeviceIndex:
        sh = self._shared
        if sh["dev"] is None:
            with sh["lock"]:
                if sh["dev"] is None:
                    sh["dev"] = DeviceIndex(self.host, device=self.device, doc_base=self.doc_base)
        return sh["dev"]

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_shared"] = None            # device handles never travel; re-uploaded lazily
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._shared = {"dev": None, "lock": threading.RLock()}

    # ------------------------------------------------------------ ExtensionArray bits
    @classmethod
    def _from_sequence(cls, scalars, dtype=None, copy=False):
        return cls(list(scalars))

    def __len__(self):
        return len(self.doc_lens) if self.rows is None else len(self.rows)

    @property
    def nbytes(self):
        return self.host.words.nbytes + self.host.doc_lens.nbytes

    def memory_usage(self, deep=False):
        return self.nbytes

    def isna(self):
        return self.doclengths() == 0

    def _row_terms(self, doc_id):
        post = {}
        key = np.uint64(doc_id + self.doc_base) << np.uint64(36)     # words carry ABSOLUTE doc ids (shards)
        nxt = np.uint64(doc_id + self.doc_base + 1) << np.uint64(36)
        for t in range(self.host.n_terms):
            w = self.host.term_words(t)
            a, b = np.searchsorted(w, key), np.searchsorted(w, nxt)
            if b > a:
                post[self.term_dict.get_term(t)] = decode_positions(w[a:b])
        return Terms(post, doc_len=int(self.host.doc_lens[doc_id]))

    def __getitem__(self, key):
        key = pd.api.indexers.check_array_indexer(self, key)
        if isinstance(key, numbers.Integral):
            n = len(self)
            if key < 0:
                key += n
            if not 0 <= key < n:
                raise IndexError("index out of bounds")
            doc = key if self.rows is None else int(self.rows[key])
            return self._row_terms(doc)
        rows = np.arange(len(self.doc_lens))[key] if self.rows is None else self.rows[key]
        view = SearchArray.__new__(SearchArray)
        view.__dict__.update(self.__dict__)
        view.rows = np.ascontiguousarray(rows, dtype=np.uint64)
        # Reference quirk: `arr.doc_lens = self.doc_lens[key]` (postings.py:353) is a STRIDED VIEW
        # for a stepped slice and bm25_score walks it contiguously (bm25.pyx:34-41), i.e. BM25 on
        # arr[a::s] sees the parent's doc_lens[a], [a+1], ...  Reproduced for parity.
        view._bm25_doc_lens = None
        if isinstance(key, slice) and key.step not in (None, 1) and len(rows):
            parent = self.doclengths()
            first = int(np.arange(len(parent))[key][0])
            if first + len(rows) <= len(parent):
                view._bm25_doc_lens = np.ascontiguousarray(parent[first:first + len(rows)])
        return view

    def take(self, indices, allow_fill=False, fill_value=None):
        idx = np.asarray(indices)
        if allow_fill and (idx < 0).any():
            raise NotImplementedError("take with fill is host-side pandas plumbing (out of scope)")
        return self[idx]

    def copy(self):
        c = SearchArray.__new__(SearchArray)
        c.__dict__.update(self.__dict__)
        return c

    @classmethod
    def _concat_same_type(cls, to_concat):
        if len(to_concat) == 1:
            return to_concat[0]
        raise NotImplementedError("concatenation re-indexes; out of scope for the scoring path")

    def __eq__(self, other):
        if isinstance(other, SearchArray):
            if len(self) != len(other):
                return False
            return np.asarray([a == b for a, b in zip(self, other)], dtype=bool)
        return NotImplemented

    # ******************************************************************************
    # Search API (reference postings.py:604-708)
    # ******************************************************************************
    def _check_token_arg(self, token):
        if isinstance(token, str):
            return token
        elif isinstance(token, list) and len(token) == 1:
            return token[0]
        elif isinstance(token, list):
            return token
        else:
            raise TypeError("Expected a string or list of strings for phrases")

    def _term_id(self, token):
        try:
            return self.term_dict.get_term_id(token)
        except TermMissingError:
            return _lib.NO_TERM

    @staticmethod
    def _payload_bounds(min_posn, max_posn):
        """RoaringishEncoder.slice argument checks (reference roaringish.py:267-282)."""
        if min_posn is None and max_posn is None:
            return 0, _lib.ALL_BITS
        if min_posn is not None and min_posn % 18 != 0:
            raise ValueError("min_payload must be a multiple of 18")
        if max_posn is not None and max_posn % 18 != 17:
            raise ValueError("max_payload must be a multiple of 18 - 1")
        lo = 0 if min_posn is None else min_posn
        hi = _lib.ALL_BITS if max_posn is None else max_posn
        return lo // 18, hi // 18

    def _apply_rows(self, dev):
        """Installs (or clears) this view's row filter on the shared device index."""
        if self.rows is None:
            if dev._rows_set:
                _lib.check(_lib.lib().sa_index_set_rows(dev.handle, None, 0))
                dev._rows_set = False
        else:
            _lib.check(_lib.lib().sa_index_set_rows(dev.handle, _lib.p_u64(self.rows), len(self.rows)))
            dev._rows_set = True

    def termfreqs(self, token: Union[List[str], str], slop: int = 0,
                  min_posn: Optional[int] = None, max_posn: Optional[int] = None) -> np.ndarray:
        token = self._check_token_arg(token)
        lo, hi = self._payload_bounds(min_posn, max_posn)
        dev = self._device()
        out = _pool.empty_f32(len(self))
        with self._shared["lock"]:
            self._apply_rows(dev)
            if isinstance(token, list):
                ids = np.asarray([self._term_id(t) for t in token], dtype=np.uint32)
                _lib.check(_lib.lib().sa_phrase_freqs(dev.handle, _lib.p_u32(ids), len(ids), int(slop),
                                                      lo, hi, _lib.p_f32(out)))
            else:
                _lib.check(_lib.lib().sa_termfreqs(dev.handle, self._term_id(token), lo, hi, _lib.p_f32(out)))
        return out

    def docfreq(self, token: str) -> int:
        if not isinstance(token, str):
            raise TypeError("Expected a string")
        tid = self._term_id(token)
        if tid == _lib.NO_TERM:
            return 0
        if self.global_df is not None and self.rows is None:
            return np.uint64(self.global_df[tid])
        dev = self._device()
        df = ctypes.c_uint64(0)
        with self._shared["lock"]:
            if self.rows is None:
                _lib.check(_lib.lib().sa_docfreq(dev.handle, tid, ctypes.byref(df)))
            else:
                self._apply_rows(dev)
                _lib.check(_lib.lib().sa_docfreq_rows(dev.handle, tid, ctypes.byref(df)))
        return np.uint64(df.value)

    def doclengths(self) -> np.ndarray:
        if self.rows is None:
            return self.doc_lens
        return self.doc_lens[self.rows.astype(np.int64)]

    def score(self, token: Union[str, List[str]], similarity: Similarity = default_bm25, slop: int = 0,
              min_posn: Optional[int] = None, max_posn: Optional[int] = None) -> np.ndarray:
        """Score each doc (reference postings.py:652-680).  With a bm25_similarity the whole
        chain postings -> tf -> BM25 is one fused kernel; any other callable receives the
        GPU-computed dense tf vector like the reference."""
        token = self._check_token_arg(token)
        tokens_l = [token] if isinstance(token, str) else token
        all_dfs = np.asarray([self.docfreq(t) for t in tokens_l])
        if not isinstance(similarity, Bm25Similarity):
            tfs = self.termfreqs(token, min_posn=min_posn, max_posn=max_posn, slop=slop)
            return similarity(tfs, all_dfs, self.doclengths(), self.avg_doc_length, self.corpus_size)
        lo, hi = self._payload_bounds(min_posn, max_posn)
        if self.rows is not None:
            # sliced array: tf on the filtered postings (FilteredPosns), then BM25 over the slice's
            # rows with the slice's doc lengths -- both on the GPU (reference postings.py:674-680)
            from . import ops
            tfs = self.termfreqs(token, min_posn=min_posn, max_posn=max_posn, slop=slop)
            if self.avg_doc_length == 0:
                return np.zeros_like(tfs)
            dl = getattr(self, "_bm25_doc_lens", None)
            if dl is None:
                dl = np.ascontiguousarray(self.doclengths())
            idf = compute_idf(self.corpus_size, all_dfs)
            return ops.bm25_score(tfs, dl, self.avg_doc_length, idf, similarity.k1, similarity.b, device=self.device)
        out = _pool.empty_f32(len(self))
        if self.avg_doc_length == 0:
            out[:] = 0
            return out
        idf = compute_idf(self.corpus_size, all_dfs)
        dev = self._device()
        with self._shared["lock"]:
            self._apply_rows(dev)
            if isinstance(token, list):
                ids = np.asarray([self._term_id(t) for t in token], dtype=np.uint32)
                _lib.check(_lib.lib().sa_score_phrase(dev.handle, _lib.p_u32(ids), len(ids), int(slop), idf,
                                                      self.avg_doc_length, similarity.k1, similarity.b,
                                                      lo, hi, _lib.p_f32(out)))
            else:
                _lib.check(_lib.lib().sa_score_term(dev.handle, self._term_id(token), idf, self.avg_doc_length,
                                                    similarity.k1, similarity.b, lo, hi, _lib.p_f32(out)))
        return out

    def positions(self, token: str, key=None) -> List[np.ndarray]:
        """Positions of a term per doc (reference postings.py:682-687): index-time decode, host."""
        tid = self.term_dict.get_term_id(token)
        docs = np.arange(len(self.doc_lens)) if self.rows is None else self.rows.astype(np.int64)
        if key is not None:
            docs = docs[key]
        w = self.host.term_words(tid)
        out = []
        for d in np.atleast_1d(docs):
            d = int(d) + self.doc_base                       # words carry ABSOLUTE doc ids (shards)
            a = np.searchsorted(w, np.uint64(d) << np.uint64(36))
            b = np.searchsorted(w, np.uint64(d + 1) << np.uint64(36))
            out.append(decode_positions(w[a:b]))
        return out

    # -------------------------------------------------- batched, HBM-resident path
    def search_topk(self, queries, k=10, similarity: Bm25Similarity = default_bm25, slop=0):
        """queries: list of str (term) or list[str] (phrase).  Returns (docs uint32[Q,k],
        scores float32[Q,k]); scores never leave HBM except the top-k (sa_score_batch_topk).
        Defined on unsliced arrays only: a sliced view scores through .score() (FilteredPosns semantics)."""
        if self.rows is not None:
            raise ValueError("search_topk on a sliced SearchArray is not supported: the batched path scores the "
                             "whole shard; use .score() on the slice")
        terms, starts, idfs = [], [0], []
        for q in queries:
            toks = [q] if isinstance(q, str) else list(q)
            terms.extend(self._term_id(t) for t in toks)
            starts.append(len(terms))
            idfs.append(compute_idf(self.corpus_size, np.asarray([self.docfreq(t) for t in toks])))
        terms = np.asarray(terms, dtype=np.uint32)
        starts = np.asarray(starts, dtype=np.uint32)
        idfs = np.asarray(idfs, dtype=np.float32)
        docs = np.empty((len(queries), k), dtype=np.uint32)
        scores = np.empty((len(queries), k), dtype=np.float32)
        dev = self._device()
        with self._shared["lock"]:
            _lib.check(_lib.lib().sa_score_batch_topk(dev.handle, _lib.p_u32(terms), _lib.p_u32(starts),
                                                      _lib.p_f32(idfs), len(queries), int(slop),
                                                      self.avg_doc_length, similarity.k1, similarity.b, k,
                                                      _lib.p_u32(docs), _lib.p_f32(scores)))
        return docs, scores

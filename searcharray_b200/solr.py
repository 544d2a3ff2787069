"""Solr-style edismax over a DataFrame of SearchArray columns -- host mirror of the reference's
`searcharray.solr` (solr.py:10-355) with the vector arithmetic on the GPU.

The query parsing, the mm mini-language and the explain string are host-side string work, as in
the reference.  Every score vector -- the per-(term, field) BM25 vectors, the phrase-phase vectors
of pf / pf2 / pf3 on the arrays sliced to the qf matches, and the combined vector -- is produced and
combined in HBM through the `sa_multi_*` entry points (include/searcharray_b200.h); only the final
vector (`edismax`) or its top-k (`edismax_topk`) comes back.  There is no CPU fallback: a custom
(non-BM25) similarity or a sliced frame still runs every `.score` on the GPU and only the final
element-wise combination in numpy, the way the reference composes it.
"""
import ctypes
import re
import threading
import weakref
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import pandas as pd

from . import _lib
from .postings import SearchArray
from .similarity import Bm25Similarity, Similarity, compute_idf, default_bm25


# --------------------------------------------------------------------------- parsing
def parse_min_should_match(num_clauses: int, spec: str) -> int:
    """Solr's `mm` (reference solr.py:10-59): ints, negatives, percentages, `n<spec` conditionals."""
    def as_int(text):
        try:
            return int(text)
        except ValueError:
            raise ValueError("Invalid 'mm' spec. Expecting an integer.")

    spec = spec.strip()
    if "<" in spec:
        result = num_clauses
        for clause in re.sub(r"\s*<\s*", "<", spec).split():
            bound, sep, rest = clause.partition("<")
            if not sep:
                raise ValueError("Invalid 'mm' spec: '" + clause + "'. Expecting values before and after '<'")
            if num_clauses <= as_int(bound):
                return result
            result = parse_min_should_match(num_clauses, rest)
        return result
    if "%" in spec:
        calc = (num_clauses * as_int(spec[:-1])) * (1 / 100)
        result = num_clauses + int(calc) if calc < 0 else int(calc)
    else:
        calc = as_int(spec)
        result = num_clauses + calc if calc < 0 else calc
    return min(num_clauses, max(result, 0))


def parse_field_boosts(field_lists: Optional[List[str]]) -> dict:
    """`["title^2", "body"]` -> {"title": 2.0, "body": None} (reference solr.py:62-74)."""
    out = {}
    for spec in field_lists or []:
        parts = spec.split("^")
        out[parts[0]] = None if len(parts) == 1 else float(parts[1])
    return out


def get_field(frame, field) -> SearchArray:
    if field not in frame.columns:
        raise ValueError(f"Field {field} not in dataframe")
    if not isinstance(frame[field].array, SearchArray):
        raise ValueError(f"Field {field} is not a searcharray field")
    return frame[field].array


def parse_query_terms(frame: pd.DataFrame, query: str, query_fields: List[str]):
    """Tokenise the query with every field's own tokenizer (reference solr.py:85-114)."""
    search_terms: Dict[str, List[str]] = {}
    num_search_terms, term_centric = 0, True
    for field in query_fields:
        toks = list(get_field(frame, field).tokenizer(query))
        search_terms[field] = toks
        if num_search_terms == 0:
            num_search_terms = len(toks)
        elif len(toks) != num_search_terms:
            term_centric = False
    return num_search_terms, search_terms, term_centric


def _boost_text(boost):
    return f"{boost}" if boost is not None else "1"


# ------------------------------------------------------------------- device multi handle
class _Multi:
    """An sa_multi over the device indexes of some fields (cached per field tuple)."""

    def __init__(self, arrays: List[SearchArray]):
        self.devs = [a._device() for a in arrays]          # keeps the field handles alive
        handles = (ctypes.c_void_p * len(self.devs))(*[d.handle for d in self.devs])
        self.handle = ctypes.c_void_p()
        _lib.check(_lib.lib().sa_multi_create(handles, len(self.devs), ctypes.byref(self.handle)))
        self.lock = threading.Lock()
        self._finalizer = weakref.finalize(self, _Multi._destroy, self.handle)

    @staticmethod
    def _destroy(handle):
        if handle and _lib._lib is not None:
            _lib._lib.sa_multi_destroy(handle)


# optional per-call wall-clock accounting (tools / bench diagnostics): set to a dict to collect
_TIMING = None


def _timed(name, fn, *args):
    if _TIMING is None:
        return fn(*args)
    import time
    t0 = time.perf_counter()
    rc = fn(*args)
    _TIMING[name] = _TIMING.get(name, 0.0) + time.perf_counter() - t0
    return rc


_multis: Dict[tuple, _Multi] = {}
_multis_lock = threading.Lock()


def _multi_for(arrays: List[SearchArray]) -> _Multi:
    key = tuple(id(a._device()) for a in arrays)
    with _multis_lock:
        m = _multis.pop(key, None)
        if m is None:
            while len(_multis) >= 16:                 # evict the least recently used; a caller still
                _multis.pop(next(iter(_multis)))      # holding it keeps it alive until it is done
            m = _Multi(arrays)
        _multis[key] = m                              # most recently used last
        return m


class _locked:
    """Holds the multi's lock AND every participating array's device lock (in a fixed order) for a whole
    edismax evaluation: the sa_multi_* calls keep intermediate state in the field indexes' scratch
    buffers, which a concurrent .score()/.termfreqs() on the same array would overwrite."""

    def __init__(self, multi: _Multi, arrays: List[SearchArray]):
        uniq = {id(a._shared["lock"]): a._shared["lock"] for a in arrays}
        self.locks = [multi.lock] + [uniq[k] for k in sorted(uniq)]

    def __enter__(self):
        for lk in self.locks:
            lk.acquire()
        return self

    def __exit__(self, *exc):
        for lk in reversed(self.locks):
            lk.release()
        return False


def _u32(values):
    return np.asarray(values, dtype=np.uint32)


def _f32(values):
    return np.asarray(values, dtype=np.float32)


class _Plan:
    """Everything edismax derives from its arguments before any scoring."""

    def __init__(self, frame, q, qf, mm, pf, pf2, pf3, tie, q_op, similarity):
        listify = lambda x: x if isinstance(x, list) else [x]
        self.query_fields = parse_field_boosts(listify(qf))
        self.phrase_fields = parse_field_boosts(listify(pf)) if pf else {}
        self.bigram_fields = parse_field_boosts(pf2) if pf2 else {}
        self.trigram_fields = parse_field_boosts(pf3) if pf3 else {}
        mm = "1" if mm is None else (f"{mm}" if isinstance(mm, int) else mm)
        self.mm = "100%" if q_op == "AND" else mm
        self.tie = tie
        if not isinstance(similarity, dict):
            similarity = {field: similarity for field in self.query_fields}
        for field in self.query_fields:
            similarity.setdefault(field, default_bm25)
        self.similarity = similarity
        self.names = list(self.query_fields)
        self.arrays = [get_field(frame, f) for f in self.names]
        self.num_terms, self.search_terms, self.term_centric = parse_query_terms(frame, q, self.names)

    def device_ok(self):
        return (all(isinstance(self.similarity[f], Bm25Similarity) for f in self.names)
                and all(a.rows is None for a in self.arrays)
                and len({len(a) for a in self.arrays}) == 1 and len(self.names) <= 8
                and all(len(t) <= 16 for t in self.search_terms.values()))

    # explain strings, reference solr.py:133-147, 160-178, 199-200, 219-220, 241-242
    def explain_qf(self):
        if self.term_centric:
            need = parse_min_should_match(self.num_terms, spec=self.mm)
            groups = ["(" + " | ".join(f"{f}:{self.search_terms[f][i]}^{_boost_text(b)}"
                                         for f, b in self.query_fields.items()) + ")"
                      for i in range(self.num_terms)]
            return "(" + " ".join(groups) + f")~{need}"
        parts = []
        for f, b in self.query_fields.items():
            toks = self.search_terms[f]
            need = min(parse_min_should_match(len(toks), spec=self.mm), len(toks))
            parts.append("((" + " ".join(f"{f}:{t}" for t in toks) + f")~{need})^{_boost_text(b)}")
        return " | ".join(parts)

    def phases(self):
        """[(phase name, [(field, boost, [phrase token lists], repeat_last)])] in reference order."""
        def grams(toks, n):
            return [toks[i:i + n] for i in range(len(toks) - n + 1)]
        out = []
        pf = [(f, b, [self.search_terms[f]], False) for f, b in self.phrase_fields.items()
              if len(self.search_terms[f]) >= 2]
        pf2 = [(f, b, grams(self.search_terms[f], 2), True) for f, b in self.bigram_fields.items()
               if len(self.search_terms[f]) >= 2]
        pf3 = [(f, b, grams(self.search_terms[f], 3), False) for f, b in self.trigram_fields.items()
               if len(self.search_terms[f]) >= 3]
        for name, items in (("pf", pf), ("pf2", pf2), ("pf3", pf3)):
            out.append((name, items))
        return out

    def explain_phases(self):
        text = ""
        for _, items in self.phases():
            for f, b, phrases, _ in items:
                for ph in phrases:
                    text += f" ({f}:\"{' '.join(ph)}\")^{_boost_text(b)}"
        return text


def _run_device(plan: _Plan, multi: _Multi) -> _Multi:
    """qf phase + phrase phases into the multi's HBM-resident combined vector (caller holds `_locked`)."""
    L = _lib.lib()
    for arr in plan.arrays:                    # a sliced view of the same column may have left its row filter installed
        arr._apply_rows(arr._device())
    F = len(plan.names)
    n_terms, tids, idfs, boosts, has_boost, avgdl, k1, b, mms = [], [], [], [], [], [], [], [], []
    for f, arr in zip(plan.names, plan.arrays):
        toks = plan.search_terms[f]
        sim = plan.similarity[f]
        n_terms.append(len(toks))
        tids.extend(arr._term_id(t) for t in toks)
        idfs.extend(compute_idf(arr.corpus_size, np.asarray([arr.docfreq(t)])) for t in toks)
        boost = plan.query_fields[f]
        boosts.append(0.0 if boost is None else boost)
        has_boost.append(0 if boost is None else 1)
        avgdl.append(arr.avg_doc_length)
        k1.append(sim.k1)
        b.append(sim.b)
        if plan.term_centric:
            mms.append(parse_min_should_match(plan.num_terms, spec=plan.mm))
        else:
            mms.append(min(parse_min_should_match(len(toks), spec=plan.mm), len(toks)))
    n_matches = ctypes.c_uint64(0)
    a_nt, a_tid, a_idf = _u32(n_terms), _u32(tids if tids else [0]), _f32(idfs if idfs else [0])
    a_boost, a_hb, a_avgdl, a_k1, a_b, a_mm = _f32(boosts), _u32(has_boost), _f32(avgdl), _f32(k1), _f32(b), _u32(mms)
    _lib.check(_timed("qf", L.sa_multi_qf, multi.handle, 0 if plan.term_centric else 1, _lib.p_u32(a_nt), _lib.p_u32(a_tid),
                             _lib.p_f32(a_idf), _lib.p_f32(a_boost), _lib.p_u32(a_hb), _lib.p_f32(a_avgdl),
                             _lib.p_f32(a_k1), _lib.p_f32(a_b), _lib.p_u32(a_mm), float(plan.tie),
                             ctypes.byref(n_matches)))
    phases = plan.phases()
    comm = plan.arrays[0].comm            # doc-range shards: match counts and filtered dfs are global
    total_matches = n_matches.value if comm is None else int(comm.sum_u64([n_matches.value])[0])
    if total_matches == 0 or not any(items for _, items in phases):
        return multi

    # every phrase of one field runs in one launch on that field's lists filtered to qf > 0
    field_index = {f: i for i, f in enumerate(plan.names)}
    per_field: Dict[str, List[Tuple[str, int]]] = {}          # field -> [(phase, phrase no)] in row order
    for name, items in phases:
        for f, _, phrases, _ in items:
            if f not in field_index:
                raise KeyError(f)                              # like the reference: pf fields must be in qf
            rows = per_field.setdefault(f, [])
            rows.extend((name, i) for i in range(len(phrases)))
    row_of: Dict[Tuple[str, str, int], int] = {}
    for f, rows in per_field.items():
        fi, arr, sim = field_index[f], plan.arrays[field_index[f]], plan.similarity[f]
        toks = plan.search_terms[f]
        uniq = list(dict.fromkeys(toks))
        slot = {t: i for i, t in enumerate(uniq)}
        u_ids = _u32([arr._term_id(t) for t in uniq])
        dfs = np.zeros(len(uniq), dtype=np.uint64)
        _lib.check(_timed("filter", L.sa_multi_filter, multi.handle, fi, _lib.p_u32(u_ids), len(uniq), _lib.p_u64(dfs)))
        if comm is not None:
            dfs = comm.sum_u64(dfs)
        starts, slots, ids, p_idf = [0], [], [], []
        phrase_lists = {name: phrases for name, items in phases for ff, _, phrases, _ in items if ff == f}
        for r, (name, i) in enumerate(rows):
            ph = phrase_lists[name][i]
            row_of[(f, name, i)] = r
            slots.extend(slot[t] for t in ph)
            ids.extend(int(u_ids[slot[t]]) for t in ph)
            starts.append(len(slots))
            p_idf.append(compute_idf(arr.corpus_size, np.asarray([dfs[slot[t]] for t in ph])))
        a_st, a_sl, a_id, a_pi = _u32(starts), _u32(slots), _u32(ids), _f32(p_idf)
        _lib.check(_timed("phrases", L.sa_multi_phrases, multi.handle, fi, len(rows), _lib.p_u32(a_st), _lib.p_u32(a_sl),
                                      _lib.p_u32(a_id), _lib.p_f32(a_pi), arr.avg_doc_length, sim.k1, sim.b))
    for name, items in phases:
        e_field, e_row, e_boost, e_hb = [], [], [], []
        for f, boost, phrases, repeat_last in items:
            order = list(range(len(phrases))) + ([len(phrases) - 1] if repeat_last else [])
            for i in order:
                e_field.append(field_index[f])
                e_row.append(row_of[(f, name, i)])
                e_boost.append(0.0 if boost is None else boost)
                e_hb.append(0 if boost is None else 1)
        if e_field:
            a_f, a_r, a_bo, a_h = _u32(e_field), _u32(e_row), _f32(e_boost), _u32(e_hb)
            _lib.check(_timed("add_phase", L.sa_multi_add_phase, multi.handle, len(e_field), _lib.p_u32(a_f), _lib.p_u32(a_r),
                                            _lib.p_f32(a_bo), _lib.p_u32(a_h)))
    return multi


def _run_composed(plan: _Plan) -> np.ndarray:
    """Custom similarity callables / sliced frames: every `.score` still runs on the GPU, the
    element-wise combination follows the reference's numpy composition (solr.py:117-355)."""
    arrays = dict(zip(plan.names, plan.arrays))
    n = len(plan.arrays[0])

    def boosted(vec, boost):
        return vec * (1 if boost is None else boost)

    if plan.term_centric:
        term_vecs = []
        for i in range(plan.num_terms):
            hi, tot = np.zeros(n), np.zeros(n)
            for f, boost in plan.query_fields.items():
                s = boosted(arrays[f].score(plan.search_terms[f][i], similarity=plan.similarity[f]), boost)
                tot += s
                hi = np.maximum(hi, s)
            term_vecs.append(hi + (tot - hi) * plan.tie)
        need = parse_min_should_match(plan.num_terms, spec=plan.mm)
        ok = np.sum(np.asarray(term_vecs) > 0, axis=0) >= need
        scores = np.sum(term_vecs, axis=0)
        scores[~ok] = 0
    else:
        field_vecs = []
        for f, boost in plan.query_fields.items():
            toks = plan.search_terms[f]
            ts = np.array([arrays[f].score(t, similarity=plan.similarity[f]) for t in toks])
            need = min(parse_min_should_match(len(toks), spec=plan.mm), len(toks))
            ok = np.sum(ts > 0, axis=0) >= need
            tot = np.sum(ts, axis=0)
            tot[~ok] = 0
            field_vecs.append(boosted(tot, boost))
        stacked = np.asarray(field_vecs)
        tot, hi = np.sum(stacked, axis=0), np.max(stacked, axis=0)
        scores = hi + (tot - hi) * plan.tie
    sliced = {f: arrays[f][scores > 0] for f in plan.names}
    for _, items in plan.phases():
        parts = []
        for f, boost, phrases, repeat_last in items:
            vec = None
            for ph in phrases:
                vec = boosted(sliced[f].score(ph, similarity=plan.similarity[f]), boost)
                parts.append(vec)
            if repeat_last:
                parts.append(vec)
        if parts:
            scores[np.where(scores)[0]] += np.sum(parts, axis=0)
    return scores


# ------------------------------------------------------------------------------ API
def edismax(frame: pd.DataFrame, q: str, qf: List[str], mm: Optional[Union[str, int]] = None,
            pf: Optional[List[str]] = None, pf2: Optional[List[str]] = None, pf3: Optional[List[str]] = None,
            ps2: int = 0, ps3: int = 0, ps: int = 0, tie: float = 0.0, q_op: str = "OR",
            similarity: Union[Similarity, Dict[str, Similarity]] = default_bm25) -> Tuple[np.ndarray, str]:
    """Same signature and result as the reference's `edismax` (solr.py:251-355): the score vector
    over the frame's rows (float64 term-centric, float32 field-centric) and the explain string.
    ps / ps2 / ps3 are accepted and ignored, as in the reference (quirk vii)."""
    plan = _Plan(frame, q, qf, mm, pf, pf2, pf3, tie, q_op, similarity)
    explain = plan.explain_qf() + plan.explain_phases()
    if not plan.device_ok():
        return _run_composed(plan), explain
    n = len(plan.arrays[0])
    multi = _multi_for(plan.arrays)
    with _locked(multi, plan.arrays):
        _run_device(plan, multi)
        out = np.empty(n, dtype=np.float64 if plan.term_centric else np.float32)
        _lib.check(_lib.lib().sa_multi_download(multi.handle, out.ctypes.data_as(ctypes.c_void_p),
                                                0 if plan.term_centric else 1))
    return out, explain


def edismax_topk(frame: pd.DataFrame, q: str, qf: List[str], k: int = 10, mm: Optional[Union[str, int]] = None,
                 pf: Optional[List[str]] = None, pf2: Optional[List[str]] = None, pf3: Optional[List[str]] = None,
                 tie: float = 0.0, q_op: str = "OR",
                 similarity: Union[Similarity, Dict[str, Similarity]] = default_bm25):
    """The k best rows of `edismax(...)` (score desc, row asc) without moving the score vector
    off the GPU.  Returns (rows uint32[k], scores float64[k]); unused slots are 0xFFFFFFFF / 0."""
    plan = _Plan(frame, q, qf, mm, pf, pf2, pf3, tie, q_op, similarity)
    if not plan.device_ok():
        raise NotImplementedError("edismax_topk needs BM25 similarities on unsliced SearchArray columns")
    docs = np.empty(k, dtype=np.uint32)
    scores = np.empty(k, dtype=np.float64)
    multi = _multi_for(plan.arrays)
    with _locked(multi, plan.arrays):
        _run_device(plan, multi)
        _lib.check(_timed("topk", _lib.lib().sa_multi_topk, multi.handle, k, _lib.p_u32(docs),
                                            scores.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    comm = plan.arrays[0].comm
    if comm is not None:                  # one all-gather of the per-shard top-k, merged on every rank
        return comm.merge_topk_f64(docs, scores, k)
    return docs, scores

"""Seeded synthetic MSMARCO-shaped corpora, generated as postings (not text).

SURVEY.md section 8d: a Python tokenizer over ~5e8 tokens would take hours, so the benchmark
corpus is generated directly in the index's upload format and injected the way
SearchArray.index injects its build (reference postings.py:293-299).  The generator itself is
plain C + pthreads (csrc/sa_synth.c -> libsa_synth.so, counter-based random streams): data
infrastructure for bench.py and the tests, not part of the scoring path.

  doc_lens ~ floor(clip(lognormal(3.9, 0.45), 8, 400))   (mean ~ 55 tokens, MSMARCO-passage-like)
  vocabulary: 1,024 query terms, df/N in {3e-1, 1e-1, 3e-2, 1e-2, 1e-3, 1e-4} (~171 per bucket)
  per (term, doc): tf ~ Geometric(0.6) capped at 8 and at doc_len, positions uniform in the doc
  phrases over that vocabulary, planted as exact (and, every second plant, gapped) occurrences:
     "rare"   4 terms, one of df/N <= 1e-3 at a random slot (exercises L->R / R->L / middle-out)
     "hard"   4 terms, ALL of df/N >= 1e-2 (no short list to drive the intersection)
     "bigram" common (df/N 3e-1) x mid (df/N 3e-2): BASELINE.md's 4.5M x 0.45M-word case

The doc-id space is cut into `N_BLOCKS` fixed blocks, each (term, block) drawn from its own
seeded stream, so any rank of a 1/2/4/8-GPU run can generate exactly its doc range of the SAME
global corpus (doc-range sharding, section 8e) without generating the rest.
"""
import ctypes
import os

import numpy as np

from .indexing import HostIndex, TermDict

SEED = 20260924
N_BLOCKS = 64
DF_BUCKETS = (3e-1, 1e-1, 3e-2, 1e-2, 1e-3, 1e-4)
MAX_TF = 8

_HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH_LIB_PATH = os.path.join(_HERE, "libsa_synth.so")
_synth_lib = None


def _lib():
    global _synth_lib
    if _synth_lib is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise RuntimeError(f"{SYNTH_LIB_PATH} not found: build it with `python -m searcharray_b200.build`")
        L = ctypes.CDLL(SYNTH_LIB_PATH)
        c = ctypes
        L.sa_synth_run.restype = c.c_void_p
        L.sa_synth_run.argtypes = [c.c_uint64, c.c_uint64, c.c_uint64, c.c_uint32, c.c_uint32, c.c_uint32,
                                   c.c_double, c.c_double, c.c_float, c.c_float,
                                   c.c_uint32, c.c_void_p, c.c_uint32, c.c_void_p, c.c_void_p, c.c_void_p,
                                   c.c_void_p, c.c_int]
        L.sa_synth_dims.restype = None
        L.sa_synth_dims.argtypes = [c.c_void_p, c.POINTER(c.c_uint64), c.POINTER(c.c_uint64), c.c_void_p]
        L.sa_synth_copy.restype = None
        L.sa_synth_copy.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int]
        L.sa_synth_free.restype = None
        L.sa_synth_free.argtypes = [c.c_void_p]
        L.sa_synth_doc_len_sum.restype = c.c_double
        L.sa_synth_doc_len_sum.argtypes = [c.c_uint64, c.c_uint64, c.c_uint64, c.c_uint32, c.c_double, c.c_double,
                                           c.c_float, c.c_float, c.c_int]
        _synth_lib = L
    return _synth_lib


def _rng(*key):
    return np.random.default_rng([SEED, *key])


def block_bounds(n_docs):
    return [(n_docs * b) // N_BLOCKS for b in range(N_BLOCKS + 1)]


# Fields of the two-field (edismax) corpus: the body is the MSMARCO-passage-like field above; the
# title is a short field (mean ~ 6 tokens) over the SAME vocabulary with rarer terms (SURVEY 8d,
# config 5).  `key` separates the fields' random streams.
FIELDS = {
    "body": {"key": 0, "len": (3.9, 0.45, 8, 400), "df_scale": 1.0, "plant_scale": 1.0},
    "title": {"key": 101, "len": (1.7, 0.4, 1, 30), "df_scale": 0.15, "plant_scale": 0.25},
}


class SynthSpec:
    """Names + generation parameters of every term and phrase of the synthetic vocabulary."""

    def __init__(self, n_docs, terms_per_bucket=None, n_phrases=256, n_hard=None, n_bigrams=32, field="body",
                 n_terms=1024):
        self.n_docs = int(n_docs)
        self.field = field
        nb = len(DF_BUCKETS)
        if terms_per_bucket is None:
            per = [n_terms // nb + (1 if b < n_terms % nb else 0) for b in range(nb)]
        else:
            per = [int(terms_per_bucket)] * nb
        self.per_bucket = per
        self.terms = []            # (name, df_fraction, bucket)
        self.bucket_terms = []
        for bi, p in enumerate(DF_BUCKETS):
            names = [f"b{bi}_{j}" for j in range(per[bi])]
            self.bucket_terms.append(names)
            self.terms.extend((nm, p, bi) for nm in names)
        self.term_index = {t[0]: i for i, t in enumerate(self.terms)}
        # stratified order: round-robin over the df buckets, every term exactly once
        self.query_terms = []
        for j in range(max(per)):
            for bi in range(nb):
                if j < per[bi]:
                    self.query_terms.append(self.bucket_terms[bi][j])

        # ---- phrases over the vocabulary
        r = _rng(7)
        if n_hard is None:
            n_hard = n_phrases // 4
        n_rare = n_phrases - n_hard
        common = [nm for bi in range(4) for nm in self.bucket_terms[bi]]
        self.phrases = []

        def pick_common(k, exclude=()):
            out = []
            while len(out) < k:
                nm = common[int(r.integers(0, len(common)))]
                if nm not in out and nm not in exclude:
                    out.append(nm)
            return out

        can_phrase = len(common) >= 4 and per[4] > 0 and per[5] > 0
        for i in range(n_rare if can_phrase else 0):
            rb = 4 + (i % 2)
            rare = self.bucket_terms[rb][(i // 2) % per[rb]]
            rare_slot = int(r.integers(0, 4))
            others = pick_common(3)
            names = others[:rare_slot] + [rare] + others[rare_slot:]
            plant_frac = (1e-2, 1e-1)[(i // 2) % 2]
            self.phrases.append({"terms": names, "kind": "rare", "rare_slot": rare_slot, "rare_p": DF_BUCKETS[rb],
                                 "plant_p": DF_BUCKETS[rb] * plant_frac, "gapped": 1})
        for i in range(n_hard if can_phrase else 0):
            self.phrases.append({"terms": pick_common(4), "kind": "hard", "plant_p": 2e-5, "gapped": 1})
        for i in range(n_bigrams if can_phrase else 0):
            a = self.bucket_terms[0][i % per[0]]
            b = self.bucket_terms[2][i % per[2]]
            self.phrases.append({"terms": [a, b] if i % 2 == 0 else [b, a], "kind": "bigram", "plant_p": 1e-4,
                                 "gapped": 0})


def _threads(n_threads):
    if n_threads:
        return int(n_threads)
    world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
    return max(1, min(64, (os.cpu_count() or 1) // world))


def generate_shard(spec: SynthSpec, rank=0, world=1, n_threads=None):
    """HostIndex for the doc range owned by `rank` of `world` (absolute doc ids kept).
    Returns (host_index, doc_lo, doc_hi)."""
    assert N_BLOCKS % world == 0, "world size must divide the number of blocks"
    per = N_BLOCKS // world
    fld = FIELDS[spec.field]
    mu, sigma, lo_len, hi_len = fld["len"]
    term_p = np.ascontiguousarray([t[1] * fld["df_scale"] for t in spec.terms], dtype=np.float64)
    ph_start = np.zeros(len(spec.phrases) + 1, dtype=np.uint32)
    ph_terms = []
    for g, ph in enumerate(spec.phrases):
        ph_terms.extend(spec.term_index[t] for t in ph["terms"])
        ph_start[g + 1] = len(ph_terms)
    ph_terms = np.ascontiguousarray(ph_terms, dtype=np.uint32)
    ph_p = np.ascontiguousarray([ph["plant_p"] * fld["plant_scale"] for ph in spec.phrases], dtype=np.float64)
    ph_gap = np.ascontiguousarray([ph["gapped"] for ph in spec.phrases], dtype=np.uint32)
    L = _lib()

    def ptr(a):
        return a.ctypes.data_as(ctypes.c_void_p) if len(a) else None

    h = L.sa_synth_run(SEED, fld["key"], spec.n_docs, N_BLOCKS, rank * per, (rank + 1) * per,
                       mu, sigma, lo_len, hi_len, len(spec.terms), ptr(term_p),
                       len(spec.phrases), ptr(ph_start), ptr(ph_terms), ptr(ph_p), ptr(ph_gap), _threads(n_threads))
    try:
        lo, hi = ctypes.c_uint64(0), ctypes.c_uint64(0)
        lens = np.zeros(len(spec.terms), dtype=np.uint64)
        L.sa_synth_dims(h, ctypes.byref(lo), ctypes.byref(hi), lens.ctypes.data_as(ctypes.c_void_p))
        words = np.empty(int(lens.sum()), dtype=np.uint64)
        doc_lens = np.empty(hi.value - lo.value, dtype=np.float32)
        L.sa_synth_copy(h, words.ctypes.data_as(ctypes.c_void_p), doc_lens.ctypes.data_as(ctypes.c_void_p),
                        _threads(n_threads))
    finally:
        L.sa_synth_free(h)
    td = TermDict()
    for name, _, _ in spec.terms:
        td.add_term(name)
    offs = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.uint64) if len(lens) else lens
    # avg_doc_length must be the GLOBAL value on every shard (section 8e): callers pass
    # global_avg_doc_length(spec); the shard mean is not used.
    host = HostIndex(words, offs, lens, doc_lens, td, avg_doc_length=None)
    return host, int(lo.value), int(hi.value)


def global_avg_doc_length(spec: SynthSpec):
    """float32(exact float64 sum of every block's doc lengths / n_docs): identical on every rank."""
    fld = FIELDS[spec.field]
    mu, sigma, lo_len, hi_len = fld["len"]
    total = _lib().sa_synth_doc_len_sum(SEED, fld["key"], spec.n_docs, N_BLOCKS, mu, sigma, lo_len, hi_len,
                                        _threads(None))
    return np.float32(total / spec.n_docs)


def stratified_term_queries(spec: SynthSpec, n_queries):
    """`n_queries` single-term queries cycling over the df buckets (stratified); with
    the default vocabulary the first 1,024 are all DISTINCT terms."""
    names = spec.query_terms
    return [names[i % len(names)] for i in range(n_queries)]


def phrase_queries(spec: SynthSpec, n_queries, kinds=("rare", "hard")):
    """4-term phrase queries: the spec's distinct planted phrases of the given kinds, rare and hard
    interleaved 3:1 like they were created; cycles when n_queries exceeds them."""
    rare = [ph for ph in spec.phrases if ph["kind"] == "rare" and "rare" in kinds]
    hard = [ph for ph in spec.phrases if ph["kind"] == "hard" and "hard" in kinds]
    pool, i, j = [], 0, 0
    while i < len(rare) or j < len(hard):
        for _ in range(3):
            if i < len(rare):
                pool.append(rare[i]); i += 1
        if j < len(hard):
            pool.append(hard[j]); j += 1
    if not pool:
        return []
    return [list(pool[q % len(pool)]["terms"]) for q in range(n_queries)]


def phrase_kinds(spec: SynthSpec, queries):
    """kind ("rare" / "hard" / "bigram") of each query produced by phrase_queries / bigram_queries."""
    kind = {tuple(ph["terms"]): ph["kind"] for ph in spec.phrases}
    return [kind[tuple(q)] for q in queries]


def bigram_queries(spec: SynthSpec, n_queries):
    pool = [ph for ph in spec.phrases if ph["kind"] == "bigram"]
    return [list(pool[q % len(pool)]["terms"]) for q in range(n_queries)] if pool else []


def edismax_queries(spec: SynthSpec, n_queries, seed=13):
    """Mixed 2-5 term queries for the two-field edismax workload: a run of
    2-4 terms of a planted phrase (so pf / pf2 / pf3 find matches), optionally followed by a
    term of the single-term vocabulary."""
    r = _rng(5, seed)
    pool = [ph for ph in spec.phrases if ph["kind"] in ("rare", "hard")]
    out = []
    for _ in range(n_queries):
        ph = pool[int(r.integers(0, len(pool)))]["terms"]
        n = int(r.integers(2, 5))
        at = int(r.integers(0, 4 - n + 1))
        toks = list(ph[at:at + n])
        if r.random() < 0.5:
            toks.append(spec.query_terms[int(r.integers(0, len(spec.query_terms)))])
        out.append(" ".join(toks))
    return out

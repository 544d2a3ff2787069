"""Seeded synthetic MSMARCO-shaped corpora, generated as postings (not text).

SURVEY.md section 8d: a Python tokenizer over ~5e8 tokens would take hours, so the benchmark
corpus is generated directly in the index's upload format and injected the way
SearchArray.index injects its build (reference postings.py:293-299).

  doc_lens ~ clip(lognormal(3.9, 0.45), 8, 400)   (mean ~ 55 tokens, MSMARCO-passage-like)
  query vocabulary: df/N in {3e-1, 1e-1, 3e-2, 1e-2, 1e-3, 1e-4}, `terms_per_bucket` each
  per (term, doc): tf ~ Geometric(0.6) capped at 8 and at doc_len, positions uniform in the doc
  phrase groups: 4 terms, planted as an exact phrase in a fraction of the rarest term's docs

The doc-id space is cut into `N_BLOCKS` fixed blocks, each (term, block) drawn from its own
seeded stream, so any rank of a 1/2/4/8-GPU run can generate exactly its doc range of the SAME
global corpus (doc-range sharding, section 8e) without generating the rest.
"""
import numpy as np

from .indexing import HostIndex, TermDict
from .roaringish import encode_postings

SEED = 20260924
N_BLOCKS = 8
DF_BUCKETS = (3e-1, 1e-1, 3e-2, 1e-2, 1e-3, 1e-4)
MAX_TF = 8


def _rng(*key):
    return np.random.default_rng([SEED, *key])


def block_bounds(n_docs):
    return [(n_docs * b) // N_BLOCKS for b in range(N_BLOCKS + 1)]


# Fields of the two-field (edismax) corpus: the body is the MSMARCO-passage-like field above; the
# title is a short field (mean ~ 6 tokens) over the SAME vocabulary with rarer terms (SURVEY 8d,
# config 5).  `key` separates the fields' random streams; the body keeps key () so that the
# single-field corpus is unchanged.
FIELDS = {
    "body": {"key": (), "len": (3.9, 0.45, 8, 400), "df_scale": 1.0, "plant_scale": 1.0},
    "title": {"key": (101,), "len": (1.7, 0.4, 1, 30), "df_scale": 0.15, "plant_scale": 0.25},
}


def gen_doc_lens(n_docs, block, field="body"):
    lo, hi = block_bounds(n_docs)[block], block_bounds(n_docs)[block + 1]
    mu, sigma, lo_len, hi_len = FIELDS[field]["len"]
    r = _rng(*FIELDS[field]["key"], 0, block)
    return np.clip(r.lognormal(mu, sigma, hi - lo), lo_len, hi_len).astype(np.float32)


def _postings_for(rng, doc_lens, doc0, p, planted=None):
    """Random (doc, posn) pairs for one term in one block, merged with planted pairs."""
    n = len(doc_lens)
    docs = np.flatnonzero(rng.random(n, dtype=np.float32) < p)
    tf = np.minimum(np.minimum(rng.geometric(0.6, size=len(docs)), MAX_TF), doc_lens[docs].astype(np.int64))
    u = rng.random((len(docs), MAX_TF), dtype=np.float32)
    pos = (u * doc_lens[docs][:, None]).astype(np.int64)
    keep = np.arange(MAX_TF)[None, :] < tf[:, None]
    key = ((docs[:, None] + doc0) << 18 | pos)[keep]
    if planted is not None and len(planted):
        key = np.concatenate([key, planted])
    key = np.unique(key)                       # sorted by (doc, posn), duplicates dropped
    return key >> 18, key & 0x3FFFF


class SynthSpec:
    """Names + generation parameters of every term of the synthetic vocabulary."""

    def __init__(self, n_docs, terms_per_bucket=8, n_phrase_groups=8, field="body"):
        self.n_docs = n_docs
        self.field = field
        self.terms = []            # (name, df_fraction, phrase_group or -1, slot in group)
        for bi, p in enumerate(DF_BUCKETS):
            for j in range(terms_per_bucket):
                self.terms.append((f"b{bi}_{j}", p, -1, 0))
        self.query_terms = [t[0] for t in self.terms]
        # phrase groups: one rare term at a random slot, three commoner ones
        self.phrases = []
        r = _rng(7)
        for g in range(n_phrase_groups):
            rare_slot = int(r.integers(0, 4))
            rare_p = (1e-3, 1e-4)[g % 2]
            plant_frac = (1e-2, 1e-1)[(g // 2) % 2]
            names = []
            for s in range(4):
                p = rare_p if s == rare_slot else float(r.choice([1e-2, 3e-2, 1e-1, 3e-1]))
                name = f"p{g}_{s}"
                self.terms.append((name, p, g, s))
                names.append(name)
            self.phrases.append({"terms": names, "rare_slot": rare_slot, "rare_p": rare_p,
                                 "plant_frac": plant_frac})
        self.term_index = {t[0]: i for i, t in enumerate(self.terms)}


def generate_shard(spec: SynthSpec, rank=0, world=1, progress=None):
    """HostIndex for the doc range owned by `rank` of `world` (absolute doc ids kept).
    Returns (host_index, doc_lo, doc_hi)."""
    assert N_BLOCKS % world == 0, "world size must divide the number of blocks"
    per = N_BLOCKS // world
    blocks = range(rank * per, (rank + 1) * per)
    bounds = block_bounds(spec.n_docs)
    doc_lo, doc_hi = bounds[blocks[0]], bounds[blocks[-1] + 1]
    fld = FIELDS[spec.field]
    fkey = fld["key"]
    dl_blocks = {b: gen_doc_lens(spec.n_docs, b, spec.field) for b in blocks}
    doc_lens = np.concatenate([dl_blocks[b] for b in blocks])

    # planted phrase occurrences per (group, block): docs that hold the rare term get the
    # exact phrase at a random start with probability plant_frac
    plants = {}
    for g, ph in enumerate(spec.phrases):
        for b in blocks:
            dl = dl_blocks[b]
            r = _rng(*fkey, 2, g, b)
            docs = np.flatnonzero(r.random(len(dl), dtype=np.float32) <
                                  ph["rare_p"] * ph["plant_frac"] * fld["plant_scale"])
            if spec.field != "body":
                docs = docs[dl[docs] >= 4]                 # the phrase must fit
            start = (r.random(len(docs)) * np.maximum(dl[docs] - 4, 1)).astype(np.int64)
            plants[(g, b)] = (docs + bounds[b], start)

    td = TermDict()
    word_lists, offs, lens, total = [], [], [], 0
    for ti, (name, p, g, slot) in enumerate(spec.terms):
        parts = []
        for b in blocks:
            planted = None
            if g >= 0:
                pdocs, pstart = plants[(g, b)]
                planted = (pdocs << 18) | (pstart + slot)
            d, pos = _postings_for(_rng(*fkey, 1, ti, b), dl_blocks[b], bounds[b], p * fld["df_scale"], planted)
            parts.append(encode_postings(d, pos))
        w = np.concatenate(parts) if len(parts) > 1 else parts[0]
        td.add_term(name)
        word_lists.append(w)
        offs.append(total)
        lens.append(len(w))
        total += len(w)
        if progress:
            progress(ti, len(spec.terms))
    words = np.concatenate(word_lists)
    # avg_doc_length must be the GLOBAL value on every shard (section 8e): a fixed constant of
    # the generator (float32 mean of the full corpus is rank-dependent to compute), so use the
    # analytic-free approach: mean over this shard is NOT used; callers pass the global value.
    host = HostIndex(words, offs, lens, doc_lens, td, avg_doc_length=None)
    return host, doc_lo, doc_hi


def stratified_term_queries(spec: SynthSpec, n_queries, seed=11):
    """`n_queries` single-term queries cycling over the df buckets (stratified)."""
    r = _rng(3, seed)
    names = spec.query_terms
    per_bucket = len(names) // len(DF_BUCKETS)
    out = []
    for i in range(n_queries):
        b = i % len(DF_BUCKETS)
        out.append(names[b * per_bucket + int(r.integers(0, per_bucket))])
    return out


def phrase_queries(spec: SynthSpec, n_queries, seed=12):
    r = _rng(4, seed)
    return [list(spec.phrases[int(r.integers(0, len(spec.phrases)))]["terms"]) for _ in range(n_queries)]


def edismax_queries(spec: SynthSpec, n_queries, seed=13):
    """Mixed 2-5 term queries for the two-field edismax workload: a run of
    2-4 terms of a planted phrase group (so pf / pf2 / pf3 find matches), optionally followed by a
    term of the single-term vocabulary."""
    r = _rng(5, seed)
    out = []
    for _ in range(n_queries):
        ph = spec.phrases[int(r.integers(0, len(spec.phrases)))]["terms"]
        n = int(r.integers(2, 5))
        at = int(r.integers(0, 4 - n + 1))
        toks = list(ph[at:at + n])
        if r.random() < 0.5:
            toks.append(spec.query_terms[int(r.integers(0, len(spec.query_terms)))])
        out.append(" ".join(toks))
    return out

"""CPU-only checks of the host side: index builder vs the reference's index (golden), and that
the C-ABI library loads and exports every symbol include/searcharray_b200.h declares."""
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import search as osearch


@pytest.fixture(scope="module")
def api():
    return (np.load(os.path.join(GOLDEN, "api.npz")), json.load(open(os.path.join(GOLDEN, "api.json"))))


def test_library_exports_every_declared_symbol():
    from searcharray_b200 import _lib
    from searcharray_b200.build import build
    build()
    header = open(os.path.join(ROOT, "include", "searcharray_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = _lib.lib()                      # raises if any SIGNATURES symbol is missing
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_host_indexer_matches_reference_index(api):
    g, meta = api
    from searcharray_b200.indexing import build_index
    from searcharray_b200.postings import ws_tokenizer
    host = build_index(meta["docs"], ws_tokenizer)
    assert host.term_dict.id_to_terms == meta["terms"]      # first-seen term ids
    assert np.array_equal(host.doc_lens, g["doc_lens"])
    assert np.float32(host.avg_doc_length) == g["avg_doc_length"][0]
    lens = g["index_lens"].astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)])
    for t in range(len(lens)):
        assert np.array_equal(host.term_words(t), g["index_words"][offs[t]:offs[t + 1]]), meta["terms"][t]


def test_encoder_matches_oracle_encode():
    from searcharray_b200.roaringish import encode_postings, decode_positions
    rng = np.random.default_rng(5)
    docs = np.sort(rng.integers(0, 50, 400))
    posns = np.concatenate([np.sort(rng.choice(3000, size=(docs == d).sum(), replace=False)) for d in np.unique(docs)])
    w = encode_postings(docs, posns)
    assert np.array_equal(w, osearch.encode(docs, posns))
    assert np.all(np.diff(w.astype(np.uint64)) > 0)
    d0 = np.unique(docs)[0]
    sel = w[(w >> np.uint64(36)) == d0]
    assert np.array_equal(decode_positions(sel), posns[docs == d0])


def test_too_long_doc_raises_and_truncate():
    from searcharray_b200.indexing import build_index
    from searcharray_b200.roaringish import MAX_POSN
    big = ["x"] * (MAX_POSN + 5)
    with pytest.raises(ValueError):
        build_index([big], lambda d: d)
    host = build_index([big], lambda d: d, truncate=True)
    assert host.doc_lens[0] == MAX_POSN


def test_shard_partition_covers_index(api):
    g, meta = api
    from searcharray_b200.indexing import build_index
    from searcharray_b200.postings import ws_tokenizer
    host = build_index(meta["docs"], ws_tokenizer)
    n = host.n_docs
    cuts = [0, n // 3, 2 * n // 3, n]
    shards = [host.shard(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    for t in range(host.n_terms):
        joined = np.concatenate([s.term_words(t) for s in shards])
        assert np.array_equal(joined, host.term_words(t))
    assert sum(s.n_docs for s in shards) == n


def test_shard_comm_merge_topk_f64():
    """ShardComm.merge_topk_f64 (the host merge of the all-gathered per-shard edismax top-k): fake a
    2-rank all-gather and compare with a sort of the union."""
    from searcharray_b200.shard import ShardComm
    rng = np.random.default_rng(3)
    k = 10
    per_rank = []
    for r in range(2):
        n = 7 if r == 0 else 10
        docs = np.full(k, 0xFFFFFFFF, dtype=np.uint32)
        scores = np.zeros(k)
        docs[:n] = rng.choice(1000, size=n, replace=False) + 1000 * r
        scores[:n] = np.sort(rng.random(n))[::-1]
        per_rank.append((docs, scores))
    per_rank[1][1][2] = per_rank[0][1][1]          # a tie across shards: lower doc id first

    class Fake(ShardComm):
        """merge_topk_f64 gathers the doc ids first, then the score bit patterns."""

        def __init__(self, rank):
            super().__init__(None, rank, 2)
            self.calls = 0

        def allgather_u64(self, values):
            assert np.array_equal(np.asarray(values, dtype=np.uint64),
                                  per_rank[self.rank][self.calls].astype(np.uint64) if self.calls == 0
                                  else per_rank[self.rank][1].view(np.uint64))
            col = self.calls
            self.calls += 1
            return np.stack([per_rank[r][0].astype(np.uint64) if col == 0 else per_rank[r][1].view(np.uint64)
                             for r in range(2)])

    all_d = np.concatenate([p[0] for p in per_rank]).astype(np.int64)
    all_s = np.concatenate([p[1] for p in per_rank])
    keep = all_d != 0xFFFFFFFF
    order = np.lexsort((all_d[keep], -all_s[keep]))[:k]
    for rank in range(2):
        d, s = Fake(rank).merge_topk_f64(*per_rank[rank], k)
        assert np.array_equal(d, all_d[keep][order].astype(np.uint32))
        assert np.array_equal(s, all_s[keep][order])


def test_synth_title_field_and_edismax_queries():
    """The second (title-like) field of the two-field corpus: same vocabulary, short docs, rarer
    terms; shards of a 2-rank split concatenate to the 1-rank corpus; the body field is unchanged by
    the field parameter."""
    from searcharray_b200 import synth
    n = 40_000
    body = synth.SynthSpec(n)
    title = synth.SynthSpec(n, field="title")
    assert [t[0] for t in body.terms] == [t[0] for t in title.terms]
    hb, _, _ = synth.generate_shard(body)
    ht, lo, hi = synth.generate_shard(title)
    assert (lo, hi) == (0, n) and ht.n_terms == hb.n_terms
    assert 1 <= ht.doc_lens.min() and ht.doc_lens.max() <= 30 and 4 < ht.doc_lens.mean() < 8
    assert len(ht.words) < len(hb.words) / 4
    parts = [synth.generate_shard(title, r, 2) for r in range(2)]
    assert parts[0][2] == parts[1][1]
    for t in range(ht.n_terms):
        assert np.array_equal(np.concatenate([p[0].term_words(t) for p in parts]), ht.term_words(t))
    assert np.array_equal(np.concatenate([p[0].doc_lens for p in parts]), ht.doc_lens)
    qs = synth.edismax_queries(body, 20)
    assert all(2 <= len(q.split()) <= 5 for q in qs)
    assert all(tok in body.term_index for q in qs for tok in q.split())


def test_pickle_round_trip_keeps_the_index_and_drops_device_state():
    """reference test/test_search.py:62-73 (pickle round trip): the host index travels, device
    handles never do (they are re-created lazily on first use)."""
    import pickle
    from searcharray_b200 import SearchArray
    arr = SearchArray.index(["foo bar bar baz", "data2", "data3 bar", "bunny funny wunny"] * 5)
    arr._shared["dev"] = object()                  # stand-in for a live device handle
    clone = pickle.loads(pickle.dumps(arr))
    assert clone._shared["dev"] is None
    assert np.array_equal(clone.host.words, arr.host.words)
    assert np.array_equal(clone.doc_lens, arr.doc_lens) and clone.avg_doc_length == arr.avg_doc_length
    assert clone.term_dict.term_to_ids == arr.term_dict.term_to_ids
    assert len(clone) == len(arr) and clone.corpus_size == arr.corpus_size
    view = pickle.loads(pickle.dumps(arr[1::2]))
    assert np.array_equal(view.rows, np.arange(len(arr))[1::2])

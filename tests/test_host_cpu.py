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

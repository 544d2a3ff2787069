"""GPU parity: term path (termfreqs / docfreq / BM25 score / top-k) vs the CPU oracle and the
golden vectors of the real reference.  Integer results bit-exact; BM25 scores compared
bit-for-bit first and within 1e-5 relative (BASELINE north_star tolerance) as the contract."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5


def assert_scores(got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    if np.array_equal(got.view(np.uint32), want.view(np.uint32)):
        return
    assert np.array_equal(got > 0, want > 0), "match mask differs"
    np.testing.assert_allclose(got, want, rtol=REL_TOL, atol=0)


@pytest.fixture(scope="module")
def api():
    return (np.load(os.path.join(GOLDEN, "api.npz")), json.load(open(os.path.join(GOLDEN, "api.json"))))


@pytest.fixture(scope="module")
def arr(api):
    from searcharray_b200 import SearchArray
    return SearchArray.index(api[1]["docs"])


def test_golden_single_term(api, arr):
    from searcharray_b200 import bm25_similarity
    g, meta = api
    n = 0
    for rec in meta["queries"]:
        if len(rec["tokens"]) != 1:
            continue
        qi, tok = rec["idx"], rec["tokens"][0]
        assert np.array_equal(arr.termfreqs(tok), g[f"q{qi}_tf"])
        assert int(arr.docfreq(tok)) == int(g[f"q{qi}_df"][0])
        assert_scores(arr.score(tok), g[f"q{qi}_score"])
        assert_scores(arr.score(tok, similarity=bm25_similarity(k1=0.9, b=0.4)), g[f"q{qi}_score_k1b"])
        assert np.array_equal(arr.termfreqs(tok, max_posn=17), g[f"q{qi}_tf_max17"])
        assert np.array_equal(arr.termfreqs(tok, min_posn=18), g[f"q{qi}_tf_min18"])
        n += 1
    assert n >= 6


def test_reference_known_answers():
    """reference test/test_search.py:19-33,91-95,121-124."""
    from searcharray_b200 import SearchArray
    data = SearchArray.index(["foo bar bar baz", "data2", "data3 bar", "bunny funny wunny"] * 25)
    assert np.array_equal(data.termfreqs("bar"), np.asarray([2, 0, 1, 0] * 25, dtype=np.float32))
    assert int(data.docfreq("bar")) == 50 and int(data.docfreq("nope")) == 0
    assert np.array_equal(data.doclengths(), np.asarray([4, 1, 2, 3] * 25, dtype=np.float32))
    assert np.allclose(data.score("bar"), np.asarray([0.37066694, 0., 0.34314217, 0.] * 25))
    assert np.array_equal(data.score("nope"), np.zeros(100, dtype=np.float32))
    with pytest.raises(TypeError):
        data.score(5)
    with pytest.raises(ValueError):
        data.termfreqs("bar", min_posn=5)


def _random_index(rng, n_docs, n_terms, max_df_frac=0.4):
    from searcharray_b200.indexing import index_from_term_postings
    from searcharray_b200.roaringish import encode_postings
    doc_lens = rng.integers(0, 300, n_docs).astype(np.float32)
    words = []
    for t in range(n_terms):
        df = max(1, int(n_docs * max_df_frac * rng.random() ** 3))
        docs = np.sort(rng.choice(n_docs, size=df, replace=False))
        tf = np.minimum(1 + rng.geometric(0.5, size=df), 40)
        d = np.repeat(docs, tf)
        p = np.concatenate([np.sort(rng.choice(700, size=k, replace=False)) for k in tf])
        words.append(encode_postings(d, p))
    return index_from_term_postings([f"t{i}" for i in range(n_terms)], words, doc_lens)


@pytest.mark.parametrize("n_docs", [1, 7, 4096, 4097, 50_001])
def test_random_index_vs_oracle(n_docs):
    from oracle import search as osearch
    from searcharray_b200 import SearchArray, bm25_similarity
    rng = np.random.default_rng(n_docs)
    host = _random_index(rng, n_docs, 12)
    arr = SearchArray.from_host_index(host)
    oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                               avg_doc_length=host.avg_doc_length)
    for t in range(host.n_terms):
        name = f"t{t}"
        assert np.array_equal(arr.termfreqs(name), oidx.termfreqs(t))
        assert int(arr.docfreq(name)) == oidx.docfreq(t)
        assert_scores(arr.score(name), oidx.score(t))
    # exotic BM25 parameters: every doc goes through the formula (NaN / inf parity, quirk i)
    for k1, b in [(0.0, 0.75), (1.2, 1.0), (2.0, 0.0)]:
        got = arr.score("t0", similarity=bm25_similarity(k1=k1, b=b))
        want = oidx.score(0, k1=k1, b=b)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert_scores(got[ok], want[ok])


def test_topk_matches_sorted_oracle():
    from oracle import search as osearch
    from searcharray_b200 import SearchArray
    rng = np.random.default_rng(11)
    host = _random_index(rng, 200_000, 10, max_df_frac=0.6)
    arr = SearchArray.from_host_index(host)
    oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                               avg_doc_length=host.avg_doc_length)
    names = [f"t{t}" for t in range(host.n_terms)] + ["missing"]
    for k in (1, 10, 32):
        docs, scores = arr.search_topk(names, k=k)
        for qi, name in enumerate(names):
            if name == "missing":
                assert np.all(docs[qi] == 0xFFFFFFFF) and np.all(scores[qi] == 0)
                continue
            s = oidx.score(qi)
            order = np.lexsort((np.arange(len(s)), -s.astype(np.float64)))[:k]
            order = order[s[order] > 0]
            assert np.array_equal(docs[qi][:len(order)], order.astype(np.uint32)), (name, k)
            assert_scores(scores[qi][:len(order)], s[order])
            assert np.all(docs[qi][len(order):] == 0xFFFFFFFF)

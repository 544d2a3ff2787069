"""GPU parity: phrase path (slop == 0).  Phrase counts, match masks and doc-id sets must be
bit-exact against the golden vectors of the real reference and against the CPU oracle."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g_bi():
    return np.load(os.path.join(GOLDEN, "bigram.npz"))


@pytest.fixture(scope="module")
def api():
    return (np.load(os.path.join(GOLDEN, "api.npz")), json.load(open(os.path.join(GOLDEN, "api.json"))))


def test_op_bigram_freqs_matches_reference(g_bi):
    from searcharray_b200 import ops
    g = g_bi
    for c in range(int(g["n_bigram"][0])):
        k = f"b{c}_"
        for cname, cont_rhs in (("R", True), ("L", False)):
            (ids, cnt), nxt = ops.bigram_freqs(g[k + "lhs"], g[k + "rhs"], cont_rhs=cont_rhs)
            assert np.array_equal(ids, g[k + cname + "_ids"]), (c, cname)
            assert np.array_equal(cnt, g[k + cname + "_cnt"]), (c, cname)
            assert np.array_equal(nxt, g[k + cname + "_next"]), (c, cname)


def _index_from_lists(lists):
    from searcharray_b200 import SearchArray
    from searcharray_b200.indexing import index_from_term_postings
    n_docs = int(max(int(w[-1] >> np.uint64(36)) for w in lists)) + 1
    doc_lens = np.full(n_docs, 10, dtype=np.float32)
    names = [f"t{i}" for i in range(len(lists))]
    return SearchArray.from_host_index(index_from_term_postings(names, lists, doc_lens)), names, n_docs


def test_golden_phrase_lists(g_bi):
    """compute_phrase_freqs on raw lists (2..7 terms; left-to-right, right-to-left, middle-out;
    duplicated lists exercise the same-term branch and its speculation retry)."""
    g = g_bi
    for c in range(int(g["n_phrase"][0])):
        k = f"p{c}_"
        n = int(g[k + "n"][0])
        arr, names, n_docs = _index_from_lists([g[k + f"t{i}"] for i in range(n)])
        want = np.zeros(n_docs, dtype=np.float32)
        ids = g[k + "ids"].astype(np.int64)
        want[ids] = g[k + "cnt"]
        got = arr.termfreqs(names)
        assert np.array_equal(got, want), (c, n)


def test_golden_api_phrases(api):
    from searcharray_b200 import SearchArray, bm25_similarity
    g, meta = api
    arr = SearchArray.index(meta["docs"])
    n = 0
    for rec in meta["queries"]:
        toks = rec["tokens"]
        if len(toks) < 2:
            continue
        qi = rec["idx"]
        assert np.array_equal(arr.termfreqs(toks), g[f"q{qi}_tf"]), toks
        got, want = arr.score(toks), g[f"q{qi}_score"]
        assert np.array_equal(got > 0, want > 0), toks
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)
        np.testing.assert_allclose(arr.score(toks, similarity=bm25_similarity(k1=0.9, b=0.4)),
                                   g[f"q{qi}_score_k1b"], rtol=1e-5, atol=0)
        n += 1
    assert n >= 15


# known answers from the reference's own test table (test/test_phrase_matches.py:17-194)
SCENARIOS = [
    (["foo bar bar baz", "data2", "data3 bar", "bunny funny wunny"] * 25, ["foo", "bar"], [1, 0, 0, 0] * 25),
    (["foo bear bar baz", "data2", "data3 bar", "bunny funny wunny"] * 25, ["foo", "bar"], [0, 0, 0, 0] * 25),
    (["foo foo bar bar baz", "data2", "data3 bar", "bunny funny wunny"] * 25, ["foo", "bar"], [1, 0, 0, 0] * 25),
    (["foo bar bar bar foo", "data2", "data3 bar", "bunny funny wunny"] * 25, ["foo", "bar"], [1, 0, 0, 0] * 25),
    (["foo bar baz baz", "data2", "data3 bar", "bunny funny wunny"] * 25, ["foo", "bar", "baz"], [1, 0, 0, 0] * 25),
    (["foo bar bar baz", "data2", "data3 bar", "bunny funny wunny"] * 25, ["term_does", "not_exist"], [0, 0, 0, 0] * 25),
]


@pytest.mark.parametrize("docs,phrase,expected", SCENARIOS)
def test_reference_scenarios(docs, phrase, expected):
    from searcharray_b200 import SearchArray
    arr = SearchArray.index(docs)
    assert np.array_equal(arr.termfreqs(phrase), np.asarray(expected, dtype=np.float32))


@pytest.mark.parametrize("phrase", ["foo bar baz", "foo bar", "foo foo foo", "foo foo bar", "foo bar bar",
                                    "foo bar bar baz buz foo bar", "foo bar bar baz buz foo foo", "foo foo"])
def test_phrase_across_block_boundaries(phrase):
    """reference test/test_phrase_matches.py:249-299: the phrase slid over 18-position blocks,
    all offsets in one corpus; checked against the CPU oracle."""
    from oracle import search as osearch
    from searcharray_b200 import SearchArray
    docs = []
    for off in range(0, 60):
        docs.append(" ".join(["dummy"] * off) + " " + phrase)
        docs.append("not match")
    arr = SearchArray.index(docs)
    host = arr.host
    oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                               avg_doc_length=host.avg_doc_length)
    toks = phrase.split()
    tids = [host.term_dict.get_term_id(t) for t in toks]
    want = oidx.termfreqs(tids)
    got = arr.termfreqs(toks)
    assert np.array_equal(got, want)
    assert np.all(got[0::2] >= 1) and np.all(got[1::2] == 0)


def test_random_corpus_vs_oracle():
    """Zipf-ish random text, phrases of 2..6 terms incl. repeated terms, vs the oracle."""
    from oracle import search as osearch
    from searcharray_b200 import SearchArray
    rng = np.random.default_rng(99)
    vocab = [f"v{i}" for i in range(12)]
    p = 1.0 / np.arange(1, 13)
    p /= p.sum()
    docs = [" ".join(rng.choice(vocab, size=int(rng.integers(1, 120)), p=p)) for _ in range(6000)]
    arr = SearchArray.index(docs)
    host = arr.host
    oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                               avg_doc_length=host.avg_doc_length)
    tid = host.term_dict.term_to_ids
    checked = 0
    for _ in range(60):
        n = int(rng.integers(2, 7))
        toks = [str(x) for x in rng.choice(vocab, size=n, p=p)]
        if rng.random() < 0.3:
            toks[1] = toks[0]
        want = oidx.termfreqs([tid[t] for t in toks])
        got = arr.termfreqs(toks)
        assert np.array_equal(got, want), toks
        sw = oidx.score([tid[t] for t in toks])
        sg = arr.score(toks)
        assert np.array_equal(sg > 0, sw > 0)
        np.testing.assert_allclose(sg, sw, rtol=1e-5, atol=0)
        checked += int(want.sum() > 0)
    assert checked > 10


def test_batch_topk_mixed_terms_and_phrases():
    """sa_score_batch_topk with term and phrase queries interleaved, vs the oracle's full sort."""
    from oracle import search as osearch
    from searcharray_b200 import SearchArray
    rng = np.random.default_rng(5)
    vocab = [f"v{i}" for i in range(10)]
    p = 1.0 / np.arange(1, 11)
    p /= p.sum()
    docs = [" ".join(rng.choice(vocab, size=int(rng.integers(1, 90)), p=p)) for _ in range(20_000)]
    arr = SearchArray.index(docs)
    host = arr.host
    oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                               avg_doc_length=host.avg_doc_length)
    tid = host.term_dict.term_to_ids
    queries = ["v0", ["v0", "v1"], "v7", ["v1", "v0", "v2"], ["v3", "v3"], "nope", ["v0", "nope"],
               ["v2", "v1", "v0", "v3"], "v9", ["v0", "v0", "v0"], ["v5", "v1"]]
    for k in (3, 10):
        got_docs, got_scores = arr.search_topk(queries, k=k)
        for qi, q in enumerate(queries):
            toks = [q] if isinstance(q, str) else q
            ids = [tid.get(t) for t in toks]
            s = oidx.score(ids[0] if len(ids) == 1 else ids)
            order = np.lexsort((np.arange(len(s)), -s.astype(np.float64)))[:k]
            order = order[s[order] > 0]
            assert np.array_equal(got_docs[qi][:len(order)], order.astype(np.uint32)), (q, k)
            np.testing.assert_allclose(got_scores[qi][:len(order)], s[order], rtol=1e-5, atol=0)
            assert np.all(got_docs[qi][len(order):] == 0xFFFFFFFF)


SWEEP_PHRASES = ["foo bar baz", "foo bar", "foo foo foo", "foo foo bar", "foo bar bar", "foo bar bar baz buz foo bar",
                 "foo bar bar baz buz foo foo", "foo foo", "foo foo bar", "foo bar bar"]


def _sweep_shapes(prefix_and_phrase):
    """the three corpus shapes of reference test/test_phrase_matches.py:249-299 and what each must return"""
    return [([prefix_and_phrase, "not match"], [1, 0]),
            (["not match"] * 100 + [prefix_and_phrase], [0] * 100 + [1]),
            ((["not match"] + [prefix_and_phrase]) * 100, ([0] + [1]) * 100)]


@pytest.mark.parametrize("phrase", SWEEP_PHRASES)
def test_block_boundary_sweep_at_reference_size(phrase):
    """reference test/test_phrase_matches.py:249-299 at its own size: 10 phrases x 100 position offsets x 3 corpus
    shapes, one small index per case, with the reference's two follow-up assertions: every bigram of the phrase
    matches wherever the phrase does (:205-212), and slop 1..3 match wherever slop 0 does (:215-222)."""
    from searcharray_b200 import SearchArray
    toks = phrase.split()
    for off in range(100):
        text = " ".join(["dummy"] * off) + " " + phrase
        for docs, expected in _sweep_shapes(text):
            arr = SearchArray.index(docs)
            got = arr.termfreqs(toks)
            assert np.array_equal(got, np.asarray(expected, dtype=np.float32)), (phrase, off, len(docs))
            if off % 9 == 0 or len(docs) == 2:
                hit = got > 0
                for a, b in zip(toks[:-1], toks[1:]):
                    assert np.all(arr.termfreqs([a, b])[hit] > 0), (phrase, off, a, b)
                for slop in (1, 2, 3):
                    assert np.all(arr.termfreqs(toks, slop=slop)[hit] > 0), (phrase, off, slop)

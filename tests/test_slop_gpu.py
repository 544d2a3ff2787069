"""GPU parity: phrase search with slop > 0 (span search) vs golden vectors of the real
reference and vs the CPU oracle.  Counts bit-exact (docs where the reference overflows its
512-slot span table -- undefined behaviour there -- are excluded)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _index_from_lists(lists):
    from searcharray_b200 import SearchArray
    from searcharray_b200.indexing import index_from_term_postings
    n_docs = int(max(int(w[-1] >> np.uint64(36)) for w in lists)) + 1
    names = [f"t{i}" for i in range(len(lists))]
    return SearchArray.from_host_index(index_from_term_postings(names, lists, np.full(n_docs, 10, dtype=np.float32))), names, n_docs


def test_golden_span_lists():
    from oracle import ops as oops, search as osearch
    g = np.load(os.path.join(GOLDEN, "bigram.npz"))
    checked = 0
    for c in range(int(g["n_phrase"][0])):
        k = f"p{c}_"
        n = int(g[k + "n"][0])
        if k + "s1_ids" not in g:
            continue
        enc = [g[k + f"t{i}"] for i in range(n)]
        arr, names, n_docs = _index_from_lists(enc)
        for slop in (1, 2, 4):
            osearch.span_search([e.copy() for e in enc], slop)
            if oops.last_span_undefined:
                continue                      # reference behaviour undefined for this input
            want = np.zeros(n_docs, dtype=np.float32)
            want[g[k + f"s{slop}_ids"].astype(np.int64)] = g[k + f"s{slop}_cnt"]
            got = arr.termfreqs(names, slop=slop)
            assert np.array_equal(got, want), (c, n, slop)
            checked += 1
    assert checked > 30


def test_golden_api_slop():
    from searcharray_b200 import SearchArray
    g = np.load(os.path.join(GOLDEN, "api.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "api.json")))
    arr = SearchArray.index(meta["docs"])
    n = 0
    for rec in meta["queries"]:
        qi, toks = rec["idx"], rec["tokens"]
        for slop in (1, 2, 3):
            key = f"q{qi}_tf_slop{slop}"
            if key in g:
                assert np.array_equal(arr.termfreqs(toks, slop=slop), g[key]), (toks, slop)
                n += 1
        if f"q{qi}_score_slop2" in g:
            got, want = arr.score(toks, slop=2), g[f"q{qi}_score_slop2"]
            assert np.array_equal(got > 0, want > 0)
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)
    assert n >= 30


def test_reference_slop_scenarios():
    """reference test/test_slop_matches.py:7-88 flavour: slop k finds what exact search misses,
    and counts/masks grow monotonically with slop (test_phrase_matches.py:206-221)."""
    from oracle import search as osearch
    from searcharray_b200 import SearchArray
    docs = ["foo bar baz", "foo x bar", "bar foo", "foo x y bar", "nothing here", "foo foo bar bar foo bar"] * 40
    arr = SearchArray.index(docs)
    host = arr.host
    oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                               avg_doc_length=host.avg_doc_length)
    tid = host.term_dict.term_to_ids
    exact = arr.termfreqs(["foo", "bar"])
    prev = exact
    for slop in range(1, 6):
        got = arr.termfreqs(["foo", "bar"], slop=slop)
        want = oidx.termfreqs([tid["foo"], tid["bar"]], slop=slop)
        assert np.array_equal(got, want), slop
        assert np.all(got >= exact)
        assert np.all((got > 0) >= (prev > 0))
        prev = got
    assert arr.termfreqs(["foo", "bar"], slop=1)[1] > 0 and exact[1] == 0

"""The reference's non-default similarities (similarity.py:41-89): oracle vs golden on the CPU,
device kernels (sa_op_similarity) vs golden on the GPU -- bit-exact, dtype included."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _cases():
    g = np.load(os.path.join(GOLDEN, "similarity.npz"))
    for c in range(int(g["n_cases"][0])):
        k1, b, avgdl, nd = g[f"c{c}_params"]
        avg = np.float32(avgdl) if c < 3 else float(avgdl)     # the last case passes a Python float
        yield g, c, float(k1), float(b), avg, int(nd), g[f"c{c}_dfs"]


def test_oracle_similarities_match_reference():
    from oracle import similarity as osim
    for g, c, k1, b, avg, nd, dfs in _cases():
        tf, dl = g["tf"], g["doc_lens"]
        for name, got in (("impact", osim.bm25_impact(tf, dfs, dl, avg, nd, k1, b)),
                          ("legacy", osim.bm25_legacy(tf, dfs, dl, avg, nd, k1, b)),
                          ("classic", osim.classic(tf, dfs, dl, avg, nd))):
            want = g[f"c{c}_{name}"]
            assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=True), (c, name)


def test_lucene_known_answers_impact():
    """reference test/test_similarity.py:16-78: bm25_impact * idf reproduces Lucene's BM25 values."""
    from oracle import similarity as osim
    for tf, df, dl, avgdl, nd, want in ((2, 14, 4, 2.7322686, 8516, 3.52482), (1, 5, 35, 50.580456, 8514, 3.8199246),
                                        (2, 7, 44, 50.580456, 8514, 4.5636616), (25, 7823, 152, 119.18542, 8516, 0.08028283)):
        imp = osim.bm25_impact(np.asarray([tf], np.float32), [df], np.asarray([dl], np.float32), avgdl, nd)
        assert np.isclose(imp * osim.idf_bm25(nd, [df]), want).all()


@pytest.mark.gpu
def test_device_similarities_match_reference():
    from searcharray_b200 import bm25_impact, bm25_legacy_similarity, classic_similarity
    for g, c, k1, b, avg, nd, dfs in _cases():
        tf, dl = g["tf"], g["doc_lens"]
        for name, sim in (("impact", bm25_impact(k1, b)), ("legacy", bm25_legacy_similarity(k1, b)),
                          ("classic", classic_similarity())):
            got = sim(tf.copy(), dfs, dl, avg, nd)
            want = g[f"c{c}_{name}"]
            assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=True), (c, name)
    assert not bm25_impact()(g["tf"], dfs, g["doc_lens"], 0, nd).any()


@pytest.mark.gpu
def test_score_with_other_similarities():
    """SearchArray.score(token, similarity=...) with the device similarities vs the oracle composition."""
    from oracle import search as osearch, similarity as osim
    from searcharray_b200 import SearchArray, bm25_impact, bm25_legacy_similarity, classic_similarity
    docs = ["foo bar bar baz", "data2", "data3 bar", "bunny funny wunny", "bar", "foo foo foo bar"] * 30
    arr = SearchArray.index(docs)
    host = arr.host
    oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                               avg_doc_length=host.avg_doc_length)
    tid = host.term_dict.term_to_ids
    for q in ("bar", ["foo", "bar"]):
        ids = tid[q] if isinstance(q, str) else [tid[t] for t in q]
        tf = oidx.termfreqs(ids)
        dfs = np.asarray([oidx.docfreq(t) for t in (ids if isinstance(ids, list) else [ids])])
        n, avg = len(docs), host.avg_doc_length
        for sim, want in ((bm25_impact(), osim.bm25_impact(tf, dfs, host.doc_lens, avg, n)),
                          (bm25_legacy_similarity(0.9, 0.4), osim.bm25_legacy(tf, dfs, host.doc_lens, avg, n, 0.9, 0.4)),
                          (classic_similarity(), osim.classic(tf, dfs, host.doc_lens, avg, n))):
            got = arr.score(q, similarity=sim)
            assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=True), (q, sim)

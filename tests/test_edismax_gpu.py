"""GPU parity: searcharray_b200.solr.edismax (HBM-resident combination, sa_multi_*) against golden
vectors of the real reference (tests/golden/edismax.npz): score vectors BIT-exact (float64 /
float32, same dtype), explain strings identical; top-k against a full sort of the golden vector."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def lower_one_token(text):
    return [text.lower()]


@pytest.fixture(scope="module")
def frame_and_golden():
    from searcharray_b200 import SearchArray
    g = np.load(os.path.join(GOLDEN, "edismax.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "edismax.json")))
    frame = pd.DataFrame({"title": SearchArray.index(meta["title"]),
                          "body": SearchArray.index(meta["body"]),
                          "tag": SearchArray.index(meta["tag"], tokenizer=lower_one_token)})
    return frame, g, meta


def _similarity(case):
    from searcharray_b200 import bm25_similarity
    if "similarity" not in case:
        return {}
    return {"similarity": {f: bm25_similarity(k1=k1, b=b) for f, (k1, b) in case["similarity"].items()}}


def test_edismax_matches_reference(frame_and_golden):
    from searcharray_b200.solr import edismax
    frame, g, meta = frame_and_golden
    for case in meta["cases"]:
        got, explain = edismax(frame, **case["kwargs"], **_similarity(case))
        want = g[case["name"]]
        assert str(got.dtype) == case["dtype"], case["name"]
        assert np.array_equal(got, want), (case["name"], float(np.abs(got - want).max()), int((got != want).sum()))
        assert explain == case["explain"], case["name"]


def test_edismax_topk(frame_and_golden):
    from searcharray_b200.solr import edismax_topk
    frame, g, meta = frame_and_golden
    for case in meta["cases"]:
        want = g[case["name"]].astype(np.float64)
        for k in (1, 10, 32):
            docs, scores = edismax_topk(frame, k=k, **case["kwargs"], **_similarity(case))
            order = np.lexsort((np.arange(len(want)), -want))[:k]
            order = order[want[order] > 0]
            assert np.array_equal(docs[:len(order)], order.astype(np.uint32)), (case["name"], k)
            assert np.array_equal(scores[:len(order)], want[order]), (case["name"], k)
            assert np.all(docs[len(order):] == 0xFFFFFFFF)


def test_edismax_composed_path_matches_device_path(frame_and_golden):
    """A similarity the device path does not recognise takes the composed path (GPU .score calls,
    numpy combination): same numbers."""
    from searcharray_b200 import bm25_similarity
    from searcharray_b200.solr import edismax
    frame, g, meta = frame_and_golden
    inner = bm25_similarity()

    def custom(term_freqs, doc_freqs, doc_lens, avg_doc_lens, num_docs):
        return inner(term_freqs, doc_freqs, doc_lens, avg_doc_lens, num_docs)

    for name in ("pf_all", "mm2", "field_centric_mm"):
        case = next(c for c in meta["cases"] if c["name"] == name)
        got, _ = edismax(frame, similarity=custom, **case["kwargs"])
        want = g[name]
        assert got.dtype == want.dtype
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=0)
        assert np.array_equal(got > 0, want > 0)


def test_topk_with_thousands_of_exact_ties():
    """more than 2,048 docs share the top score exactly (uniform short docs): edismax_topk must still
    return the k best by (score desc, row asc) instead of raising."""
    from searcharray_b200 import SearchArray
    from searcharray_b200.solr import edismax, edismax_topk
    docs = ["foo"] * 3000 + ["foo bar"] * 2500 + ["bar baz"] * 100 + ["foo"] * 2600
    frame = pd.DataFrame({"t": SearchArray.index(docs)})
    for q, k in (("foo", 10), ("foo", 32), ("bar", 7)):
        want, _ = edismax(frame, q, qf=["t"])
        order = np.lexsort((np.arange(len(want)), -want))[:k]
        order = order[want[order] > 0]
        d, s = edismax_topk(frame, q, qf=["t"], k=k)
        assert np.array_equal(d[:len(order)], order.astype(np.uint32)), (q, k, d, order)
        assert np.array_equal(s[:len(order)], want[order])

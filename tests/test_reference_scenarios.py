"""The known-answer tables of the reference's OWN tests (test/test_phrase_matches.py:17-194,
test/test_slop_matches.py:7-72, test/test_minmax_posns.py:5-42; extracted by
tests/golden/make_golden_scenarios.py) against the CPU oracle (not gpu) and against the CUDA path
through the public API (gpu).  Counts bit-exact; slop scores within 1e-5 relative."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

S = json.load(open(os.path.join(GOLDEN, "reference_scenarios.json")))


def expand(c):
    return c["base"] * c["times"]


def ws(text):
    return text.split()


def oracle_for(docs):
    from oracle import search as osearch
    from searcharray_b200.indexing import build_index
    host = build_index(docs, ws)
    idx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                              avg_doc_length=host.avg_doc_length)
    return idx, host.term_dict.term_to_ids


def ids_of(tid, phrase):
    ids = [tid.get(t) for t in phrase]
    return ids[0] if len(ids) == 1 else ids


PHRASE_IDS = [sc["name"] for sc in S["phrase"]]


@pytest.mark.parametrize("sc", S["phrase"], ids=PHRASE_IDS)
def test_oracle_phrase_scenarios(sc):
    from oracle import ops as oops
    docs = expand(sc["docs"])
    idx, tid = oracle_for(docs)
    ids = ids_of(tid, sc["phrase"])
    assert np.array_equal(idx.termfreqs(ids), np.asarray(expand(sc["expected"]), dtype=np.float32))
    for s, want in sc.get("slop", {}).items():
        got = idx.termfreqs(ids, slop=int(s))
        if not oops.last_span_undefined:
            assert np.array_equal(got, np.asarray(expand(want), dtype=np.float32)), s
    if "odd_slice" in sc:
        assert np.array_equal(idx.sliced(slice(1, None, 2)).termfreqs(ids),
                              np.asarray(expand(sc["odd_slice"]), dtype=np.float32))


@pytest.mark.parametrize("sc", S["minmax"], ids=[sc["name"] for sc in S["minmax"]])
def test_oracle_minmax_scenarios(sc):
    idx, tid = oracle_for(expand(sc["docs"]))
    got = idx.termfreqs(ids_of(tid, sc["phrase"]), min_posn=sc["min_posn"], max_posn=sc["max_posn"])
    assert np.array_equal(got, np.asarray(expand(sc["expected"]), dtype=np.float32))


@pytest.mark.parametrize("sc", S["slop"], ids=[sc["name"] for sc in S["slop"]])
def test_oracle_slop_scenarios(sc):
    docs = [sc["doc"], " empty ", sc["doc"] + " " + sc["doc"], " empty"] * 100
    idx, tid = oracle_for(docs)
    ids = ids_of(tid, sc["phrase"])
    for s, first4 in sc["scores_first4"].items():
        got = idx.score(ids, slop=int(s))
        np.testing.assert_allclose(got[:4], np.asarray(first4, dtype=np.float32), rtol=1e-5, atol=0)
        assert np.all((got[::2] > 0) == sc["match"]) and np.all(got[1::2] == 0)


# ------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("sc", S["phrase"], ids=PHRASE_IDS)
def test_gpu_phrase_scenarios(sc):
    """reference test_phrase_api + test_phrase_on_slice (test_phrase_matches.py:224-246)."""
    from searcharray_b200 import SearchArray
    arr = SearchArray.index(expand(sc["docs"]))
    phrase = sc["phrase"] if len(sc["phrase"]) > 1 else sc["phrase"][0]
    expected = np.asarray(expand(sc["expected"]), dtype=np.float32)
    before = arr.copy()
    tf = arr.termfreqs(phrase)
    assert np.array_equal(tf, expected)
    assert len(arr) == len(before)
    for s, want in sc.get("slop", {}).items():
        assert np.array_equal(arr.termfreqs(phrase, slop=int(s)), np.asarray(expand(want), dtype=np.float32)), s
    if "odd_slice" in sc:
        sl = arr[1::2]
        got = sl.termfreqs(phrase)
        assert len(got) == len(sl) and np.array_equal(got, np.asarray(expand(sc["odd_slice"]), dtype=np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("sc", S["minmax"], ids=[sc["name"] for sc in S["minmax"]])
def test_gpu_minmax_scenarios(sc):
    from searcharray_b200 import SearchArray
    arr = SearchArray.index(expand(sc["docs"]))
    got = arr.termfreqs(sc["phrase"], min_posn=sc["min_posn"], max_posn=sc["max_posn"])
    assert np.array_equal(got, np.asarray(expand(sc["expected"]), dtype=np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("sc", S["slop"], ids=[sc["name"] for sc in S["slop"]])
def test_gpu_slop_scenarios(sc):
    """reference test_phrase_slop (test_slop_matches.py:75-88)."""
    from searcharray_b200 import SearchArray
    arr = SearchArray.index([sc["doc"], " empty ", sc["doc"] + " " + sc["doc"], " empty"] * 100)
    toks = sc["phrase"]                       # already tokenised by the reference's tokenizer
    assert arr.tokenizer(" ".join(toks)) == toks
    for s in range(sc["slop"], max(sc["slop"], 10)):
        scores = arr.score(toks, slop=s)
        assert np.all((scores[::2] > 0) == sc["match"]) and np.all(scores[1::2] == 0)
        if str(s) in sc["scores_first4"]:
            np.testing.assert_allclose(scores[:4], np.asarray(sc["scores_first4"][str(s)], dtype=np.float32),
                                       rtol=1e-5, atol=0)

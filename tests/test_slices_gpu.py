"""GPU parity: sliced arrays (FilteredPosns semantics incl. df-on-slice and the strided doc_lens
quirk) and min_posn / max_posn on terms, phrases and slop, vs the golden vectors of the reference."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    return (np.load(os.path.join(GOLDEN, "api.npz")), json.load(open(os.path.join(GOLDEN, "api.json"))))


@pytest.fixture(scope="module")
def arr(api):
    from searcharray_b200 import SearchArray
    return SearchArray.index(api[1]["docs"])


def close(got, want):
    assert got.shape == want.shape
    assert np.array_equal(got > 0, want > 0)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)


def test_golden_slices(api, arr):
    g, meta = api
    odd = arr[1::2]
    mid = arr[100:700]
    assert len(odd) == 750 and len(mid) == 600
    for rec in meta["queries"]:
        qi, toks = rec["idx"], rec["tokens"]
        q = toks[0] if len(toks) == 1 else toks
        assert np.array_equal(odd.termfreqs(q), g[f"q{qi}_tf_odd"]), toks
        close(odd.score(q), g[f"q{qi}_score_odd"])
        close(mid.score(q), g[f"q{qi}_score_mid"])
    # the unsliced array still answers for all docs afterwards
    assert np.array_equal(arr.termfreqs("w0"), g["q0_tf"])


def test_golden_min_max_posn(api, arr):
    g, meta = api
    for rec in meta["queries"]:
        qi, toks = rec["idx"], rec["tokens"]
        q = toks[0] if len(toks) == 1 else toks
        assert np.array_equal(arr.termfreqs(q, max_posn=17), g[f"q{qi}_tf_max17"]), toks
        assert np.array_equal(arr.termfreqs(q, min_posn=18), g[f"q{qi}_tf_min18"]), toks


def test_slice_vs_oracle_mask_and_fancy():
    from oracle import search as osearch
    from searcharray_b200 import SearchArray
    rng = np.random.default_rng(3)
    vocab = [f"v{i}" for i in range(8)]
    docs = [" ".join(rng.choice(vocab, size=int(rng.integers(1, 60)))) for _ in range(3000)]
    arr = SearchArray.index(docs)
    host = arr.host
    oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                               avg_doc_length=host.avg_doc_length)
    tid = host.term_dict.term_to_ids
    mask = rng.random(3000) < 0.3
    sl, osl = arr[mask], oidx.sliced(mask)
    for q in ["v0", ["v1", "v2"], ["v3", "v3"], ["v0", "v1", "v2"]]:
        ids = tid[q] if isinstance(q, str) else [tid[t] for t in q]
        assert np.array_equal(sl.termfreqs(q), osl.termfreqs(ids)), q
        close(sl.score(q), osl.score(ids))
        assert np.array_equal(sl.termfreqs(q, slop=2) if not isinstance(q, str) else sl.termfreqs(q),
                              osl.termfreqs(ids, slop=2) if not isinstance(q, str) else osl.termfreqs(ids)), q
    assert int(sl.docfreq("v0")) == osl.docfreq(tid["v0"])

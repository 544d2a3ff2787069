"""Pins the CPU oracle (oracle/) against golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  CPU-only."""
import json
import os

import numpy as np
import pytest

from oracle import ops, search

U = np.uint64
HM = search.HEADER_MASK


@pytest.fixture(scope="module")
def g_ops(golden_dir):
    return np.load(os.path.join(golden_dir, "ops.npz"))


@pytest.fixture(scope="module")
def g_bi(golden_dir):
    return np.load(os.path.join(golden_dir, "bigram.npz"))


@pytest.fixture(scope="module")
def g_api(golden_dir):
    return (np.load(os.path.join(golden_dir, "api.npz")),
            json.load(open(os.path.join(golden_dir, "api.json"))))


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b)


def test_native_ops_match_reference(g_ops):
    g = g_ops
    for c in range(int(g["n_cases"][0])):
        k = f"c{c}_"
        lhs, rhs = g[k + "lhs"], g[k + "rhs"]
        a, b, cc, d = ops.intersect_with_adjacents(lhs, rhs, mask=HM)
        eq(a, g[k + "iwa_li"]); eq(b, g[k + "iwa_ri"]); eq(cc, g[k + "iwa_lai"]); eq(d, g[k + "iwa_rai"])
        a, b = ops.intersect(lhs, rhs, mask=HM)
        eq(a, g[k + "int_li"]); eq(b, g[k + "int_ri"])
        a, b = ops.adjacent(lhs, rhs, mask=HM)
        eq(a, g[k + "adj_li"]); eq(b, g[k + "adj_ri"])
        a, b = ops.intersect(lhs >> U(36), rhs >> U(36), drop_duplicates=False)
        eq(a, g[k + "keep_li"]); eq(b, g[k + "keep_ri"])
        a, b = ops.intersect(lhs >> U(36), rhs >> U(36), drop_duplicates=True)
        eq(a, g[k + "dropk_li"]); eq(b, g[k + "dropk_ri"])
        eq(ops.merge(lhs, rhs), g[k + "merge"])
        eq(ops.merge(lhs, rhs, drop_duplicates=True), g[k + "merge_drop"])
        eq(ops.unique(lhs, 36), g[k + "uniq_keys"])
        eq(ops.unique(np.sort(np.concatenate([lhs, lhs[::2]]))), g[k + "uniq"])
        ids, cnt = ops.popcount64_reduce(lhs, 36, 0x3FFFF)
        eq(ids, g[k + "pcr_ids"]); eq(cnt, g[k + "pcr_cnt"])
        assert cnt.dtype == np.float32
        eq(ops.popcount64(lhs), g[k + "pc64"])
        i2, c2 = ops.popcount_reduce_at(lhs >> U(36), lhs & U(0x3FFFF) & U(0x15555))
        eq(i2, g[k + "pra_ids"]); eq(c2, g[k + "pra_cnt"])
        i3, c3 = ops.key_sum_over(lhs >> U(36), ops.popcount64(lhs & U(0xFF)))
        eq(i3, g[k + "kso_ids"]); eq(c3, g[k + "kso_cnt"])
        rids, rcnt = ops.popcount64_reduce(rhs, 36, 0x3FFFF)
        mi, mc = ops.sort_merge_counts(ids, cnt, rids, rcnt)
        eq(mi, g[k + "smc_ids"]); eq(mc, g[k + "smc_cnt"])
        n_docs = g[k + "dense"].shape[0]
        eq(ops.as_dense(ids, cnt, n_docs), g[k + "dense"])
        eq(ops.payload_slice(lhs, 0x0000000FFFFC0000, 1, 2), g[k + "pslice"])
        tf = ops.as_dense(ids, cnt, n_docs)
        dl = g[k + "bm25_dl"]
        ops.bm25_score(tf, dl, float(np.mean(dl)), 2.345, 1.2, 0.75)
        # bit-exact incl. NaN/inf positions (dl == 0 docs exist)
        assert np.array_equal(tf.view(np.uint32), g[k + "bm25"].view(np.uint32))


def test_bigram_freqs_match_reference(g_bi):
    g = g_bi
    for c in range(int(g["n_bigram"][0])):
        k = f"b{c}_"
        for cname, cont in (("R", search.RHS), ("L", search.LHS)):
            (ids, cnt), nxt = search.bigram_freqs(g[k + "lhs"].copy(), g[k + "rhs"].copy(), cont)
            eq(np.asarray(ids, dtype=np.uint64), g[k + cname + "_ids"])
            eq(np.asarray(cnt, dtype=np.float32), g[k + cname + "_cnt"])
            eq(np.asarray(nxt, dtype=np.uint64), g[k + cname + "_next"])


def test_phrase_and_span_match_reference(g_bi):
    g = g_bi
    n_span = 0
    for c in range(int(g["n_phrase"][0])):
        k = f"p{c}_"
        n = int(g[k + "n"][0])
        enc = [g[k + f"t{i}"] for i in range(n)]
        ids, cnt = search.compute_phrase_freqs([e.copy() for e in enc])
        eq(np.asarray(ids, dtype=np.uint64), g[k + "ids"])
        eq(np.asarray(cnt, dtype=np.float32), g[k + "cnt"])
        for slop in (1, 2, 4):
            if k + f"s{slop}_ids" in g:
                sids, scnt = search.span_search([e.copy() for e in enc], slop)
                eq(sids, g[k + f"s{slop}_ids"])
                if ops.last_span_undefined == 0:
                    # (the reference writes past its 512-slot span table otherwise: undefined)
                    eq(scnt, g[k + f"s{slop}_cnt"])
                    n_span += 1
    assert n_span > 30


def _oracle_index(g, meta):
    lens = g["index_lens"].astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)])
    words = g["index_words"]
    tw = {i: words[offs[i]:offs[i + 1]] for i in range(len(lens))}
    tid = {t: i for i, t in enumerate(meta["terms"])}
    idx = search.OracleIndex(tw, g["doc_lens"], avg_doc_length=g["avg_doc_length"][0])
    return idx, tid


def test_api_matches_reference(g_api):
    g, meta = g_api
    idx, tid = _oracle_index(g, meta)
    n = len(idx)
    odd = idx.sliced(slice(1, None, 2))
    mid = idx.sliced(slice(100, 700))
    for rec in meta["queries"]:
        qi = rec["idx"]
        toks = [tid.get(t) for t in rec["tokens"]]
        q = toks[0] if len(toks) == 1 else toks
        eq(idx.termfreqs(q), g[f"q{qi}_tf"])
        assert np.array_equal(idx.score(q).view(np.uint32), g[f"q{qi}_score"].view(np.uint32))
        assert np.array_equal(idx.score(q, k1=0.9, b=0.4).view(np.uint32), g[f"q{qi}_score_k1b"].view(np.uint32))
        if len(toks) == 1:
            assert idx.docfreq(toks[0]) == int(g[f"q{qi}_df"][0])
        eq(idx.termfreqs(q, max_posn=17), g[f"q{qi}_tf_max17"])
        eq(idx.termfreqs(q, min_posn=18), g[f"q{qi}_tf_min18"])
        for slop in (1, 2, 3):
            key = f"q{qi}_tf_slop{slop}"
            if key in g:
                eq(idx.termfreqs(q, slop=slop), g[key])
        if f"q{qi}_score_slop2" in g:
            assert np.array_equal(idx.score(q, slop=2).view(np.uint32), g[f"q{qi}_score_slop2"].view(np.uint32))
        eq(odd.termfreqs(q), g[f"q{qi}_tf_odd"])
        assert np.array_equal(odd.score(q).view(np.uint32), g[f"q{qi}_score_odd"].view(np.uint32))
        assert np.array_equal(mid.score(q).view(np.uint32), g[f"q{qi}_score_mid"].view(np.uint32))


def test_lucene_known_answers():
    """Known answers from the reference's own test (test/test_similarity.py:16-49)."""
    # tf=2, df=14, doc_len=4, avgdl=2.7322686, N=8516 -> 3.52482 (Lucene explain output)
    tf = np.asarray([2.0], dtype=np.float32)
    out = search.bm25(tf, np.asarray([14]), np.asarray([4.0], dtype=np.float32), 2.7322686, 8516)
    assert np.isclose(out[0], 3.52482)

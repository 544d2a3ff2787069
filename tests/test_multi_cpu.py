"""world_size-2 `gloo` run on CPU of the multi-rank host logic (doc-range shards, global df,
per-shard top-k, all-gather, merge)."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_two_rank_gloo_sharded_topk():
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "2"
    port = str(29000 + (os.getpid() % 2000))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tests", "_mp_worker.py")]
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:]
    assert "MP_OK 2" in out.stdout, out.stdout[-3000:]


def test_shards_tile_the_corpus():
    from searcharray_b200 import synth
    spec = synth.SynthSpec(40_000, terms_per_bucket=2, n_phrases=8, n_bigrams=2)
    full, lo, hi = synth.generate_shard(spec, 0, 1)
    assert (lo, hi) == (0, 40_000)
    for world in (2, 4, 8):
        parts = [synth.generate_shard(spec, r, world) for r in range(world)]
        assert [p[1] for p in parts] == [40_000 * r // world for r in range(world)]
        assert np.array_equal(np.concatenate([p[0].doc_lens for p in parts]), full.doc_lens)
        for t in range(full.n_terms):
            assert np.array_equal(np.concatenate([p[0].term_words(t) for p in parts]), full.term_words(t))
    # deterministic whatever the number of generator threads; global avgdl = mean of the whole corpus
    one, _, _ = synth.generate_shard(spec, 0, 1, n_threads=1)
    assert np.array_equal(one.words, full.words) and np.array_equal(one.doc_lens, full.doc_lens)
    assert synth.global_avg_doc_length(spec) == np.float32(np.sum(full.doc_lens, dtype=np.float64) / 40_000)


def test_synth_corpus_is_well_formed():
    """Every term's list is sorted and header-unique (the index's upload format), df follows the
    bucket, planted phrases really occur, and the query sets are distinct."""
    from oracle import search as osearch
    from searcharray_b200 import synth
    n = 200_000
    spec = synth.SynthSpec(n, terms_per_bucket=6, n_phrases=16, n_bigrams=4)
    host, _, _ = synth.generate_shard(spec)
    for t, (name, p, _) in enumerate(spec.terms):
        w = host.term_words(t)
        assert np.all(np.diff((w >> np.uint64(18)).astype(np.int64)) > 0)
        assert np.all((w & np.uint64(0x3FFFF)) != 0)
        docs = (w >> np.uint64(36)).astype(np.int64)
        assert docs.min() >= 0 and docs.max() < n
        posn_ok = ((w >> np.uint64(18)) & np.uint64(0x3FFFF)).astype(np.int64) * 18 < host.doc_lens[docs]
        assert posn_ok.all()
        df = len(np.unique(docs))
        assert abs(df - p * n) < 6 * np.sqrt(p * n) + 0.2 * p * n + 30, (name, df, p * n)
    oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens)
    for ph in spec.phrases:
        ids = [spec.term_index[t] for t in ph["terms"]]
        exact = np.count_nonzero(oidx.termfreqs(ids))
        sloppy = np.count_nonzero(oidx.termfreqs(ids, slop=2))
        assert sloppy >= exact
        if ph["plant_p"] * n >= 20:
            assert exact >= 1, ph
            if ph["gapped"]:
                assert sloppy > exact, ph
    q = synth.stratified_term_queries(spec, len(spec.terms))
    assert len(set(q)) == len(spec.terms)
    pq = synth.phrase_queries(spec, 16)
    assert len({tuple(x) for x in pq}) == 16 and set(synth.phrase_kinds(spec, pq)) == {"rare", "hard"}
    assert all(len(x) == 2 for x in synth.bigram_queries(spec, 4))


def test_key_roundtrip_and_merge():
    from searcharray_b200.shard import merge_topk, shard_topk_keys, unpack_keys
    rng = np.random.default_rng(0)
    docs = np.arange(1000, dtype=np.uint64)
    scores = rng.random(1000).astype(np.float32)
    scores[::3] = 0
    scores[10] = scores[20]                      # a tie: lower doc id wins
    k = 7
    full = shard_topk_keys(docs, scores, k)
    a = shard_topk_keys(docs[:500], scores[:500], k)
    b = shard_topk_keys(docs[500:], scores[500:], k)
    merged = merge_topk(np.stack([a, b])[:, None, :], k)[0]
    assert np.array_equal(merged, full)
    d, s = unpack_keys(full)
    order = np.lexsort((docs, -scores.astype(np.float64)))[:k]
    assert np.array_equal(d, order.astype(np.uint32)) and np.array_equal(s, scores[order])

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
    spec = synth.SynthSpec(40_000, terms_per_bucket=1, n_phrase_groups=2)
    full, lo, hi = synth.generate_shard(spec, 0, 1)
    assert (lo, hi) == (0, 40_000)
    for world in (2, 4, 8):
        parts = [synth.generate_shard(spec, r, world) for r in range(world)]
        assert [p[1] for p in parts] == [40_000 * r // world for r in range(world)]
        assert np.array_equal(np.concatenate([p[0].doc_lens for p in parts]), full.doc_lens)
        for t in range(full.n_terms):
            assert np.array_equal(np.concatenate([p[0].term_words(t) for p in parts]), full.term_words(t))


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

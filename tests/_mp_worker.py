"""Worker for tests/test_multi_cpu.py: world_size-2 gloo run of the doc-range sharding logic
(shard generation, global df via all-reduce, per-shard top-k + all-gather + merge), with the CPU
oracle standing in for the GPU scorer."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ops, search as osearch  # noqa: E402
from searcharray_b200 import synth  # noqa: E402
from searcharray_b200.shard import merge_topk, shard_topk_keys  # noqa: E402
from searcharray_b200.similarity import compute_idf  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_docs, k = 160_000, 10
    spec = synth.SynthSpec(n_docs, terms_per_bucket=2, n_phrases=0, n_bigrams=0)
    host, lo, hi = synth.generate_shard(spec, rank, world, n_threads=2)
    assert len(host.doc_lens) == hi - lo
    # global statistics: df by all-reduce, avgdl from the generator's blocks
    df_local = np.asarray([osearch.docfreq(host.term_words(t)) for t in range(host.n_terms)], dtype=np.int64)
    import torch
    df_t = torch.from_numpy(df_local.copy())
    dist.all_reduce(df_t)
    df = df_t.numpy()
    avgdl = synth.global_avg_doc_length(spec)

    results = []
    for t in range(host.n_terms):
        idf = np.float32(compute_idf(n_docs, np.asarray([df[t]])))
        ids, tfs = osearch.termfreqs_sparse(host.term_words(t))
        if len(ids):
            s = tfs.copy()
            ops.bm25_score(s, host.doc_lens[(ids - np.uint64(lo)).astype(np.int64)], avgdl, idf, 1.2, 0.75)
        else:
            s = np.zeros(0, dtype=np.float32)
        results.append(shard_topk_keys(ids.astype(np.uint64), s, k))
    local = np.stack(results)                                   # [n_terms, k] keys, global doc ids
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    merged = merge_topk(np.stack(gathered), k)                  # [n_terms, k]

    if rank == 0:
        full, _, _ = synth.generate_shard(spec, 0, 1)
        for r in range(world):                                  # shards tile the corpus exactly
            pass
        for t in range(full.n_terms):
            idf = np.float32(compute_idf(n_docs, np.asarray([osearch.docfreq(full.term_words(t))])))
            assert osearch.docfreq(full.term_words(t)) == df[t]
            ids, tfs = osearch.termfreqs_sparse(full.term_words(t))
            s = tfs.copy()
            if len(ids):
                ops.bm25_score(s, full.doc_lens[ids.astype(np.int64)], avgdl, idf, 1.2, 0.75)
            want = shard_topk_keys(ids.astype(np.uint64), s, k)
            assert np.array_equal(merged[t], want), (t, merged[t], want)
        print("MP_OK", world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

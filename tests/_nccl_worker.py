"""Worker for tests/test_nccl_gpu.py: one process per GPU, doc-range shards of one synthetic corpus,
sa_score_batch_topk_allgather (kernels + ncclAllGather + topk_merge_kernel) -- rank 0 compares the merged
global top-k (doc ids and score bits) of term, phrase and slop-2 batches with the CPU oracle run on the
FULL corpus.  Rendezvous of the NCCL id through a file (no framework)."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from searcharray_b200 import _lib, synth  # noqa: E402
from searcharray_b200.postings import DeviceIndex  # noqa: E402
from searcharray_b200.similarity import compute_idf  # noqa: E402

K1, B = 1.2, 0.75


def main():
    rank, world, key = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    os.environ.setdefault("NCCL_SOCKET_IFNAME", "lo")          # one node: bootstrap over loopback, no IB probing
    os.environ.setdefault("NCCL_IB_DISABLE", "1")
    n_docs, k = 400_000, 10
    L = _lib.lib()
    spec = synth.SynthSpec(n_docs, terms_per_bucket=3, n_phrases=12, n_bigrams=2)
    host, lo, hi = synth.generate_shard(spec, rank, world, n_threads=4)
    avgdl = synth.global_avg_doc_length(spec)
    dev = DeviceIndex(host, device=rank, doc_base=lo)
    h = dev.handle
    uid = (ctypes.c_char * 128)()
    if rank == 0:
        _lib.check(L.sa_comm_unique_id(uid))
        with open(key + ".tmp", "wb") as f:
            f.write(bytes(uid))
        os.replace(key + ".tmp", key)
    else:
        t0 = time.time()
        while not os.path.exists(key):
            assert time.time() - t0 < 300
            time.sleep(0.05)
        with open(key, "rb") as f:
            uid = (ctypes.c_char * 128).from_buffer_copy(f.read(128))
    _lib.check(L.sa_comm_init(h, uid, rank, world))
    df = np.zeros(host.n_terms, dtype=np.uint64)
    tmp = ctypes.c_uint64(0)
    for t in range(host.n_terms):
        _lib.check(L.sa_docfreq(h, t, ctypes.byref(tmp)))
        df[t] = tmp.value
    _lib.check(L.sa_comm_allreduce_sum_u64(h, _lib.p_u64(df), len(df)))

    def run(queries, slop):
        terms = np.asarray([t for q in queries for t in q], dtype=np.uint32)
        starts = np.concatenate(([0], np.cumsum([len(q) for q in queries]))).astype(np.uint32)
        idf = np.asarray([compute_idf(n_docs, df[np.asarray(q)]) for q in queries], dtype=np.float32)
        docs = np.empty((len(queries), k), dtype=np.uint32)
        scores = np.empty((len(queries), k), dtype=np.float32)
        _lib.check(L.sa_score_batch_topk_allgather(h, _lib.p_u32(terms), _lib.p_u32(starts), _lib.p_f32(idf), len(queries),
                                                   slop, float(avgdl), K1, B, k, _lib.p_u32(docs), _lib.p_f32(scores)))
        return docs, scores, idf

    term_q = [[t] for t in range(host.n_terms)]
    phrase_q = [[spec.term_index[t] for t in ph["terms"]] for ph in spec.phrases]
    results = [(term_q, 0) + run(term_q, 0), (phrase_q, 0) + run(phrase_q, 0), (phrase_q[:6], 2) + run(phrase_q[:6], 2)]
    if rank == 0:
        from oracle import ops as oops, search as osearch
        full, _, _ = synth.generate_shard(spec, 0, 1, n_threads=4)
        oidx = osearch.OracleIndex({t: full.term_words(t) for t in range(full.n_terms)}, full.doc_lens,
                                   avg_doc_length=avgdl, corpus_size=n_docs)
        n_checked = 0
        for queries, slop, docs, scores, idf in results:
            for i, q in enumerate(queries):
                assert all(int(oidx.docfreq(t)) == int(df[t]) for t in q)
                dense = oidx.score(q[0] if len(q) == 1 else q, k1=K1, b=B, slop=slop)
                if slop and oops.last_span_undefined:
                    continue
                nz = np.flatnonzero(dense > 0)
                order = nz[np.lexsort((nz, -dense[nz].astype(np.float64)))][:k]
                assert np.array_equal(docs[i][:len(order)], order.astype(np.uint32)), (q, slop, docs[i], order)
                assert np.all(docs[i][len(order):] == 0xFFFFFFFF)
                assert np.array_equal(scores[i][:len(order)].view(np.uint32), dense[order].view(np.uint32)), (q, slop)
                n_checked += 1
        print("NCCL_OK", world, n_checked, flush=True)
    _lib.check(L.sa_comm_barrier(h))
    dev.close()


if __name__ == "__main__":
    main()

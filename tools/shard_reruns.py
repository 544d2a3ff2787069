"""Diagnostic: which single-term queries of one doc-range shard need the exact top-k re-run (candidate overflow)?
    python tools/shard_reruns.py <rank> <world> [bucket]"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from searcharray_b200 import _lib, synth
from searcharray_b200.postings import DeviceIndex

rank, world = int(sys.argv[1]), int(sys.argv[2])
buckets = [int(sys.argv[3])] if len(sys.argv) > 3 else range(6)
spec = synth.SynthSpec(10_000_000)
host, lo, hi = synth.generate_shard(spec, rank, world)
avgdl = synth.global_avg_doc_length(spec)
dev = DeviceIndex(host, device=0, doc_base=lo)
L, h = _lib.lib(), dev.handle
k = 10
for b in buckets:
    names = spec.bucket_terms[b]
    tids = np.asarray([spec.term_index[n] for n in names], dtype=np.uint32)
    bad = []
    for i, t in enumerate(tids):
        terms = np.asarray([t], dtype=np.uint32); starts = np.asarray([0, 1], dtype=np.uint32); idf = np.asarray([2.0], dtype=np.float32)
        docs = np.empty((1, k), dtype=np.uint32); scores = np.empty((1, k), dtype=np.float32); n_over = ctypes.c_uint32(0)
        _lib.check(L.sa_batch_upload(h, _lib.p_u32(terms), _lib.p_u32(starts), _lib.p_f32(idf), 1, 0, float(avgdl), 1.2, 0.75, k))
        _lib.check(L.sa_batch_execute(h))
        _lib.check(L.sa_batch_download(h, _lib.p_u32(docs), _lib.p_f32(scores), ctypes.byref(n_over)))
        if n_over.value:
            bad.append((names[i], int(host.term_lengths[t])))
    print("bucket", b, "reruns", len(bad), bad[:8], flush=True)

"""Times the batched term path per df bucket (which queries are slow?)."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from searcharray_b200 import _lib, synth
from searcharray_b200.postings import DeviceIndex
from searcharray_b200.similarity import compute_idf

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
spec = synth.SynthSpec(n_docs, n_phrase_groups=0)
t0=time.time()
host, lo, hi = synth.generate_shard(spec)
print('gen', time.time()-t0, flush=True)
L = _lib.lib()
dev = DeviceIndex(host, 0, 0)
h = dev.handle
print('upload', time.time()-t0, flush=True)
avgdl = float(np.mean(host.doc_lens))
Q, k = (int(sys.argv[3]) if len(sys.argv) > 3 else 128), 10
only = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else None
for b, p in enumerate(synth.DF_BUCKETS):
    if only is not None and b not in only:
        continue
    tids = np.asarray([spec.term_index[f"b{b}_{j % 8}"] for j in range(Q)], dtype=np.uint32)
    df = host.term_lengths[tids]  # ~ words
    idf = np.asarray([compute_idf(n_docs, np.asarray([max(1, int(x))])) for x in df], dtype=np.float32)
    starts = np.arange(Q + 1, dtype=np.uint32)
    _lib.check(L.sa_batch_upload(h, _lib.p_u32(tids), _lib.p_u32(starts), _lib.p_f32(idf), Q, 0, avgdl, 1.2, 0.75, k))
    for _ in range(3):
        _lib.check(L.sa_batch_execute(h))
    ms = ctypes.c_double(0)
    _lib.check(L.sa_timer_start(h))
    for _ in range(5):
        _lib.check(L.sa_batch_execute(h))
    _lib.check(L.sa_timer_stop(h, ctypes.byref(ms)))
    W = float(np.mean(host.term_lengths[tids]))
    per_q_us = ms.value * 1e3 / (5 * Q)
    alg = 8 * W + 4 * W / 1.3 + 4 * n_docs
    print(f"bucket df/N={p:7.0e}  W={W:10.0f}  {per_q_us:7.2f} us/query  alg {alg/1e6:6.1f} MB  -> {alg/per_q_us/1e6:6.2f} TB/s", flush=True)

#!/usr/bin/env python
"""bench.py -- queries/sec of SearchArray's scoring hot path on B200 (see BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU reference arm

Workload (config.workload): BASELINE.json configs[1] -- 10M-doc synthetic MSMARCO-shaped corpus
(searcharray_b200/synth.py, seeded, generated as postings), single-term BM25.  One "step" = one
pass over a batch of `--queries` stratified single-term queries: every query produces the dense
float32[N] BM25 score vector in HBM and its exact top-k.  For N > 1 the 10M docs are sharded by
contiguous doc-id range (strong scaling), one process per GPU, one ncclAllGather of the per-shard
top-k per batch.

  value : device-resident throughput -- query descriptors already in HBM, CUDA events on the
          library's stream around exactly K x sa_batch_execute (kernels + all-gather), max over ranks.
  e2e   : the same batch through the public C-ABI call with HOST buffers per step
          (sa_score_batch_topk: H2D of the query descriptors, kernels, D2H of the top-k).
  e2e_dense : the literal `.score()` drop-in (sa_score_term), D2H of the dense float32[N] per query.
  roofline  : term_tile_kernel, algorithmic bytes 8*W + 4*df + 4*N per query over the
          kernel's CUDA-event time, against MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline : the oracle port (oracle/, the reference's algorithm in C + numpy, warm tf cache)
          on the host cores, bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K1, B = 1.2, 0.75


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- corpus
def build_corpus(n_docs, rank, world):
    from searcharray_b200 import synth
    spec = synth.SynthSpec(n_docs)
    t0 = time.time()
    host, lo, hi = synth.generate_shard(spec, rank, world)
    # global avg doc length: exact float64 mean over ALL blocks' doc_lens (cheap), as float32
    total = 0.0
    for b in range(synth.N_BLOCKS):
        total += float(np.sum(synth.gen_doc_lens(n_docs, b), dtype=np.float64))
    avgdl = np.float32(total / n_docs)
    log(f"rank {rank}: generated docs [{lo},{hi}) {host.words.nbytes / 1e6:.0f} MB of postings "
        f"in {time.time() - t0:.1f}s, avgdl={avgdl}")
    return spec, host, lo, hi, avgdl


def make_queries(spec, n_queries):
    from searcharray_b200 import synth
    names = synth.stratified_term_queries(spec, n_queries)
    return names, np.asarray([spec.term_index[n] for n in names], dtype=np.uint32)


def idf_of(n_docs, df):
    from searcharray_b200.similarity import compute_idf
    return np.asarray([compute_idf(n_docs, np.asarray([d])) for d in df], dtype=np.float32)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def count_since(self, t_from):
        return sum(1 for t, _ in self.lines if t >= t_from)

    def stop(self, t_from=0.0):
        """Summary of the samples taken at or after t_from (the start of the timed region)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.lines:
            if t < t_from:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------- CPU reference arm
def cpu_reference_runner(host, avgdl, n_docs, k):
    """The reference's CPU path for this workload, restated by the oracle port: warm tf/df caches
    (PosnBitArray caches), as_dense + bm25_score over all N docs, np.argpartition top-k."""
    from oracle import search as osearch
    idx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                              avg_doc_length=avgdl, corpus_size=n_docs, cache=True)

    def one(term_id):
        scores = idx.score(int(term_id), k1=K1, b=B)
        top = np.argpartition(scores, -k)[-k:]            # reference utils/sort.py:24
        return top[np.argsort(-scores[top], kind="stable")]
    return idx, one


def best_thread_count(one, term_ids, cores):
    """The reference's loops release the GIL but every query allocates ~3 dense float32[N]
    temporaries (np.zeros / as_dense / argpartition): on many-core hosts a full-width thread pool
    thrashes the allocator and the memory bus.  Probe a few pool widths and keep the fastest, so
    the baseline is the best the host can do, not a strawman."""
    best, best_qps = 1, 0.0
    for th in sorted({1, 4, 8, 16, 32, 64, cores}):
        if th > cores:
            continue
        # at least one query per thread, or a wide pool is never actually exercised by the probe
        probe = term_ids[:min(len(term_ids), max(24, th))]
        dt = run_cpu_sample(one, probe, th)
        qps = len(probe) / dt
        log(f"cpu probe: {th} threads -> {qps:.1f} qps")
        if qps > best_qps:
            best, best_qps = th, qps
    return best


def cold_cpu_qps(idx, one, term_ids, n=6):
    """the reference's COLD path (`posns.clear_cache()` before each query, as
    test_msmarco.py:362-379 does): tf by popcount and df by unique on every call.  One thread."""
    sample = [int(t) for t in term_ids[:n]]
    t0 = time.perf_counter()
    for t in sample:
        idx._df_cache.clear()
        idx._tf_cache.clear()
        one(t)
    dt = time.perf_counter() - t0
    return {"value": len(sample) / dt, "unit": "queries/s", "cores": 1,
            "sample": f"{len(sample)} queries, tf/df caches cleared before each"}


def run_cpu_sample(one, term_ids, threads):
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    if threads == 1:
        for t in term_ids:
            one(t)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, term_ids))
    return time.perf_counter() - t0


def bench_reference(args, rank, world):
    if rank != 0:
        return                                  # rank 0 alone runs the CPU arm
    spec, host, lo, hi, avgdl = build_corpus(args.docs, 0, 1)
    names, term_ids = make_queries(spec, args.queries)
    cores = os.cpu_count() or 1
    idx, one = cpu_reference_runner(host, avgdl, args.docs, args.k)
    sample = term_ids[:min(len(term_ids), args.ref_sample)]
    for t in np.unique(term_ids):               # warm the tf/df caches like SearchArray.warm()
        idx.docfreq(int(t))
        idx.termfreqs(int(t))
    threads = best_thread_count(one, sample, cores)
    for _ in range(args.warmup):
        run_cpu_sample(one, sample[:max(8, len(sample) // 8)], threads)
    t = 0.0
    for _ in range(args.steps):
        t += run_cpu_sample(one, sample, threads)
    qps = args.steps * len(sample) / t
    cold = cold_cpu_qps(idx, one, sample)
    line = {
        "impl": "reference", "metric": "queries/sec (single-term BM25 + top-k) on 10M-doc synthetic MSMARCO",
        "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, len(sample)),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "host_cores": cores, "kind": "port",
                         "sample": f"{len(sample)} of the {args.queries} stratified term queries per step, "
                                   f"ThreadPool({threads}) = fastest of the probed pool widths, warm tf cache",
                         "cold": cold},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, queries_per_step):
    return {"workload": "10M-doc synthetic MSMARCO, single-term BM25, top-%d (BASELINE configs[1])" % args.k,
            "n_docs": args.docs, "queries_per_step": queries_per_step, "k": args.k,
            "corpus": "searcharray_b200.synth seed 20260924, doc_lens~clip(lognormal(3.9,.45),8,400), "
                      "48 query terms over df/N in {3e-1..1e-4}",
            "sharding": "contiguous doc-id ranges, one process per GPU",
            "cache": "inputs larger than L2: every step streams queries_per_step dense float32[N] vectors"}


# --------------------------------------------------------------------------- our arm
def bench_ours(args, rank, world):
    from searcharray_b200 import _lib
    from searcharray_b200.postings import DeviceIndex
    L = _lib.lib()
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    spec, host, lo, hi, avgdl = build_corpus(args.docs, rank, world)
    names, term_ids = make_queries(spec, args.queries)
    dev = DeviceIndex(host, device=local_rank, doc_base=lo)
    h = dev.handle

    if world > 1:
        # NCCL_DEBUG=VERSION (some images default to it) prints a banner on STDOUT, which would
        # break the one-JSON-line contract
        # ... and WARN still prints the version line.  Unset means silent; whatever level the user
        # asked for goes to a file instead of stdout.
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/sa_b200_nccl_%h_%p.log")
        # Rendezvous for the NCCL unique id without any framework: all ranks of one launch share a
        # node (contract: --nnodes=1) and a parent (the torchrun agent), so rank 0 publishes the id
        # in a file keyed by MASTER_PORT + parent pid and the others poll for it.
        key = f"/tmp/sa_b200_uid_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}.bin"
        uid = (ctypes.c_char * 128)()
        if rank == 0:
            _lib.check(L.sa_comm_unique_id(uid))
            with open(key + ".tmp", "wb") as f:
                f.write(bytes(uid))
            os.replace(key + ".tmp", key)
        else:
            t_wait = time.time()
            while not os.path.exists(key):
                if time.time() - t_wait > 600:
                    raise RuntimeError("timed out waiting for rank 0's NCCL id")
                time.sleep(0.05)
            with open(key, "rb") as f:
                uid = (ctypes.c_char * 128).from_buffer_copy(f.read(128))
        _lib.check(L.sa_comm_init(h, uid, rank, world))
        _lib.check(L.sa_comm_barrier(h))
        if rank == 0:
            os.remove(key)

    def barrier():
        if world > 1:
            _lib.check(L.sa_comm_barrier(h))

    def max_over_ranks(x):
        v = ctypes.c_double(x)
        if world > 1:
            _lib.check(L.sa_comm_allreduce_max(h, ctypes.byref(v)))
        return v.value

    # global document frequencies (idf must use unsharded df)
    df = np.zeros(host.n_terms, dtype=np.uint64)
    tmp = ctypes.c_uint64(0)
    for t in range(host.n_terms):
        _lib.check(L.sa_docfreq(h, t, ctypes.byref(tmp)))
        df[t] = tmp.value
    if world > 1:
        _lib.check(L.sa_comm_allreduce_sum_u64(h, _lib.p_u64(df), len(df)))
    idf = idf_of(args.docs, df[term_ids])
    starts = np.arange(len(term_ids) + 1, dtype=np.uint32)
    Q, k = len(term_ids), args.k
    out_docs = np.empty((Q, k), dtype=np.uint32)
    out_scores = np.empty((Q, k), dtype=np.float32)
    n_over = ctypes.c_uint32(0)

    def upload():
        _lib.check(L.sa_batch_upload(h, _lib.p_u32(term_ids), _lib.p_u32(starts), _lib.p_f32(idf), Q, 0,
                                     float(avgdl), K1, B, k))

    def execute():
        _lib.check(L.sa_batch_execute_allgather(h) if world > 1 else L.sa_batch_execute(h))

    def download():
        if world > 1:
            _lib.check(L.sa_batch_download_allgather(h, _lib.p_u32(out_docs), _lib.p_f32(out_scores),
                                                     ctypes.byref(n_over)))
        else:
            _lib.check(L.sa_batch_download(h, _lib.p_u32(out_docs), _lib.p_f32(out_scores), ctypes.byref(n_over)))
        return n_over.value

    def e2e_step():
        upload()
        execute()
        return download()

    # ---- warm-up (>= 3 full steps); the clock sampler (nvidia-smi -lms) starts here so that it is
    #      already delivering samples when the timed region begins
    clocks = ClockSampler(local_rank)
    clocks.start()
    overflow = 0
    for _ in range(max(args.warmup, 3)):
        overflow += e2e_step()

    # ---- value: device-resident, K x execute between CUDA events on the library stream
    stats = _lib.SaStats()
    upload()
    _lib.check(L.sa_stats_reset(h))
    barrier()
    t_timed = time.time()
    _lib.check(L.sa_timer_start(h))
    for _ in range(args.steps):
        execute()
    ms = ctypes.c_double(0)
    _lib.check(L.sa_timer_stop(h, ctypes.byref(ms)))
    barrier()
    dev_ms = max_over_ranks(ms.value)
    _lib.check(L.sa_stats_get(h, ctypes.byref(stats)))
    launches_value = int(stats.total_launches)
    download()
    value = args.steps * Q / (dev_ms / 1e3)

    # ---- e2e: host buffers in, top-k out, every step
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        overflow += e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = args.steps * Q / e2e_s
    # clocks: samples taken during the two timed regions (device-timed steps + e2e steps).  When those
    # are shorter than a few sampling periods (many GPUs, small shards) the same step is repeated,
    # untimed, for ~0.4 s so that the clocks under this load are still observed.
    clock_note = "sampled during the timed regions"
    if clocks.count_since(t_timed) < 4:
        n_extra = int(min(2000, max(1, 0.4 / max(dev_ms / 1e3 / args.steps, 1e-5))))
        for _ in range(n_extra):
            execute()
        barrier()
        download()
        clock_note = f"timed regions too short to sample: + {n_extra} untimed repeats of the same step"
    clk = clocks.stop(t_timed)
    clk["note"] = clock_note
    h2d = int(term_ids.nbytes + starts.nbytes + idf.nbytes + Q * 24)     # + TermQuery descriptors
    d2h = int(Q * k * 8 + Q * 4)

    # ---- roofline of the dominant kernel (per-launch CUDA events, async)
    W = host.term_lengths[term_ids].astype(np.float64)
    dfl = np.zeros(host.n_terms, dtype=np.float64)
    for t in range(host.n_terms):
        _lib.check(L.sa_docfreq(h, t, ctypes.byref(tmp)))
        dfl[t] = tmp.value
    alg_bytes_step = float(np.sum(8.0 * W + 4.0 * dfl[term_ids] + 4.0 * host.n_docs))
    _lib.check(L.sa_set_profiling(h, 1))
    _lib.check(L.sa_stats_reset(h))
    prof_steps = min(args.steps, 5)
    for _ in range(prof_steps):
        execute()
    _lib.check(L.sa_stats_get(h, ctypes.byref(stats)))
    _lib.check(L.sa_set_profiling(h, 0))
    term_ms = stats.term_kernel_ms / prof_steps
    launches_per_step = stats.term_kernel_launches / prof_steps
    achieved = alg_bytes_step / (term_ms / 1e3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of this
    # very workload (it cannot be measured live); null for any other configuration
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "term_kernel_traffic.json")))
        if tj["config"] == {"n_docs": args.docs, "queries_per_step": Q, "n_gpus": world}:
            traffic = float(tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"])
            traffic_src = tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "term_tile_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_step / launches_per_step,
                "avg_launch_ms": term_ms / launches_per_step, "launches_per_step": launches_per_step,
                "topk_select_ms_per_step": stats.topk_kernel_ms / prof_steps,
                "kernel_share_of_step": term_ms / (dev_ms / args.steps)}

    # ---- phrase workload (BASELINE configs[2]: 4-term phrase, slop 0) as an extra block
    phrase = None
    if args.phrase_queries > 0:
        from searcharray_b200 import synth
        pq_names = synth.phrase_queries(spec, args.phrase_queries)
        p_terms = np.asarray([spec.term_index[t] for ph in pq_names for t in ph], dtype=np.uint32)
        p_starts = np.arange(0, 4 * len(pq_names) + 1, 4, dtype=np.uint32)
        p_idf = np.asarray([float(np.sum(np.log(1 + (args.docs - df[[spec.term_index[t] for t in ph]].astype(np.float64) + 0.5)
                                                / (df[[spec.term_index[t] for t in ph]].astype(np.float64) + 0.5))))
                            for ph in pq_names], dtype=np.float32)
        PQ = len(pq_names)
        p_docs = np.empty((PQ, k), dtype=np.uint32)
        p_scores = np.empty((PQ, k), dtype=np.float32)

        Wp = np.asarray([[host.term_lengths[spec.term_index[t]] for t in ph] for ph in pq_names], dtype=np.float64)

        def phrase_block(slop):
            """One batched pass family of the PQ phrase queries with `slop`: device-timed steps, then the
            same steps end to end (upload + execute + top-k download)."""
            def p_upload():
                _lib.check(L.sa_batch_upload(h, _lib.p_u32(p_terms), _lib.p_u32(p_starts), _lib.p_f32(p_idf), PQ, slop,
                                             float(avgdl), K1, B, k))

            def p_download():
                if world > 1:
                    _lib.check(L.sa_batch_download_allgather(h, _lib.p_u32(p_docs), _lib.p_f32(p_scores), ctypes.byref(n_over)))
                else:
                    _lib.check(L.sa_batch_download(h, _lib.p_u32(p_docs), _lib.p_f32(p_scores), ctypes.byref(n_over)))
                return n_over.value

            p_redo = 0
            for _ in range(3):
                p_upload(); execute(); p_redo += p_download()
            p_upload()
            _lib.check(L.sa_stats_reset(h))
            barrier()
            _lib.check(L.sa_timer_start(h))
            p_steps = max(2, args.steps)
            for _ in range(p_steps):
                execute()
            _lib.check(L.sa_timer_stop(h, ctypes.byref(ms)))
            barrier()
            p_ms = max_over_ranks(ms.value)
            p_download()
            barrier()
            t0 = time.perf_counter()
            for _ in range(p_steps):
                p_upload(); execute(); p_redo += p_download()
            barrier()
            p_e2e_s = max_over_ranks(time.perf_counter() - t0)
            return {"queries_per_step": PQ, "value": p_steps * PQ / (p_ms / 1e3), "unit": "queries/s",
                    "ms_per_step": p_ms / p_steps,
                    "e2e": {"value": p_steps * PQ / p_e2e_s, "unit": "queries/s"},
                    "repairs": int(p_redo),
                    "matches_in_top1": int(np.sum(p_docs[:, 0] != 0xFFFFFFFF))}

        phrase = {"workload": "4-term phrase, slop 0 (BASELINE configs[2]), planted phrases, top-%d" % k}
        phrase.update(phrase_block(0))
        phrase["mean_words_per_query_this_shard"] = float(np.mean(np.sum(Wp, axis=1)))
        phrase["min_list_words_mean"] = float(np.mean(np.min(Wp, axis=1)))
        slop2_batch = None
        if args.slop_queries > 0:
            slop2_batch = {"workload": "4-term phrase, slop 2 (BASELINE configs[3]), same batched top-%d API" % k}
            slop2_batch.update(phrase_block(2))
        # slop = 2 (BASELINE configs[3]) goes through the per-query C-ABI call: span search is
        # latency/branch bound (informational, no roofline expectation)
        if rank == 0 and world == 1 and args.slop_queries > 0:
            from searcharray_b200.postings import _pool
            out = _pool.empty_f32(host.n_docs)
            ns = min(args.slop_queries, PQ)
            _lib.check(L.sa_set_profiling(h, 1))
            _lib.check(L.sa_stats_reset(h))
            matched, dt = 0, 0.0
            for i in range(ns):
                tids = np.ascontiguousarray(p_terms[4 * i:4 * i + 4])
                t0 = time.perf_counter()
                _lib.check(L.sa_score_phrase(h, _lib.p_u32(tids), 4, 2, float(p_idf[i]), float(avgdl), K1, B, 0,
                                             _lib.ALL_BITS, _lib.p_f32(out)))
                dt += time.perf_counter() - t0
                matched += int(np.count_nonzero(out))
            _lib.check(L.sa_stats_get(h, ctypes.byref(stats)))
            _lib.check(L.sa_set_profiling(h, 0))
            phrase["slop2_dense"] = {"workload": "4-term phrase, slop 2, sa_score_phrase per query (the .score() drop-in), "
                                                 "dense float32[N] to the host", "queries": ns,
                                     "e2e": {"value": ns / dt, "unit": "queries/s"},
                                     "kernel_ms_per_query": stats.phrase_kernel_ms / ns,
                                     "mean_matching_docs": matched / ns}
        if slop2_batch is not None:
            phrase["slop2"] = slop2_batch
        # CPU side of the phrase workload: the oracle port, a few queries (each is 50-500 ms)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import search as osearch
            oidx = osearch.OracleIndex({t: host.term_words(t) for t in range(host.n_terms)}, host.doc_lens,
                                       avg_doc_length=avgdl, corpus_size=args.docs, cache=True)
            nsamp = min(8, PQ)
            t0 = time.perf_counter()
            for i in range(nsamp):
                oidx.score([int(x) for x in p_terms[4 * i:4 * i + 4]], k1=K1, b=B)
            dt0 = time.perf_counter() - t0
            t0 = time.perf_counter()
            nslop = min(2, PQ)
            for i in range(nslop):
                oidx.score([int(x) for x in p_terms[4 * i:4 * i + 4]], k1=K1, b=B, slop=2)
            dt2 = time.perf_counter() - t0
            phrase["cpu_baseline"] = {"kind": "port", "cores": 1, "slop0_queries_per_s": nsamp / dt0,
                                      "slop2_queries_per_s": nslop / dt2,
                                      "sample": f"{nsamp} slop-0 and {nslop} slop-2 queries of the step, one thread"}
        upload()          # restore the term batch for the sections below

    # ---- edismax (the shape of BASELINE configs[4], on this run's corpus size): two fields, mixed
    #      2-5 term queries, qf + pf + pf2 + pf3, mm=2, tie=0.3 (reference test_msmarco.py:436-443);
    #      per query: sa_multi_* calls from the host mirror, top-k back (all-gathered over the shards)
    edis = None
    if args.edismax_queries > 0:
        import pandas as pd
        from searcharray_b200 import SearchArray, solr, synth
        from searcharray_b200.shard import ShardComm
        t0 = time.time()
        tspec = synth.SynthSpec(args.docs, field="title")
        thost, _, _ = synth.generate_shard(tspec, rank, world)
        ttotal = 0.0
        for blk in range(synth.N_BLOCKS):
            ttotal += float(np.sum(synth.gen_doc_lens(args.docs, blk, "title"), dtype=np.float64))
        t_avgdl = np.float32(ttotal / args.docs)
        comm = ShardComm(h, rank, world)
        body = SearchArray.from_host_index(host, device=local_rank, doc_base=lo, corpus_size=args.docs,
                                           avg_doc_length=avgdl, global_df=df, comm=comm)
        body._shared["dev"] = dev                      # the body shard is already in HBM
        title = SearchArray.from_host_index(thost, device=local_rank, doc_base=lo, corpus_size=args.docs,
                                            avg_doc_length=t_avgdl, comm=comm)
        tdev = title._device()
        tdf = np.zeros(thost.n_terms, dtype=np.uint64)
        for t in range(thost.n_terms):
            _lib.check(L.sa_docfreq(tdev.handle, t, ctypes.byref(tmp)))
            tdf[t] = tmp.value
        title.global_df = comm.sum_u64(tdf)
        frame = pd.DataFrame({"title": title, "body": body})
        log(f"edismax: title field {thost.words.nbytes / 1e6:.0f} MB of postings, avgdl={t_avgdl}, "
            f"set-up {time.time() - t0:.1f}s")
        eq = synth.edismax_queries(spec, args.edismax_queries)
        ekw = dict(qf=["title^1.0", "body^0.5"], pf=["body"], pf2=["body"], pf3=["body"], mm=2, tie=0.3)
        for qtext in eq[:3]:
            solr.edismax_topk(frame, qtext, k=k, **ekw)
        barrier()
        solr._TIMING = {}
        t0 = time.perf_counter()
        hits = 0
        for qtext in eq:
            d_, s_ = solr.edismax_topk(frame, qtext, k=k, **ekw)
            hits += int(d_[0] != 0xFFFFFFFF)
        barrier()
        e_s = max_over_ranks(time.perf_counter() - t0)
        call_ms = {kk: 1e3 * vv / len(eq) for kk, vv in solr._TIMING.items()}
        solr._TIMING = None
        edis = {"workload": "two-field edismax (title^1.0 body^0.5, pf/pf2/pf3 on body, mm=2, tie=0.3), mixed 2-5 term "
                            "queries, exact float64 top-%d, per-query host-driven sa_multi_* calls" % k,
                "queries": len(eq), "e2e": {"value": len(eq) / e_s, "unit": "queries/s"},
                "ms_per_query": 1e3 * e_s / len(eq), "queries_with_hits": hits,
                "ms_per_query_by_call": call_ms}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import search as osearch, solr as osolr
            ofields = {}
            for name, hidx, adl in (("title", thost, t_avgdl), ("body", host, avgdl)):
                oi = osearch.OracleIndex({t: hidx.term_words(t) for t in range(hidx.n_terms)}, hidx.doc_lens,
                                         avg_doc_length=adl, corpus_size=args.docs, cache=True)
                ofields[name] = osolr.OracleField(oi, hidx.term_dict.term_to_ids)
            nq = min(2, len(eq))
            t0 = time.perf_counter()
            bad = 0
            for qtext in eq[:nq]:
                want = osolr.edismax(ofields, qtext, **ekw)
                d_, s_ = solr.edismax_topk(frame, qtext, k=k, **ekw)
                order = np.lexsort((np.arange(len(want)), -want))[:k]
                order = order[want[order] > 0]
                if not (np.array_equal(d_[:len(order)], order.astype(np.uint32)) and
                        np.allclose(s_[:len(order)], want[order], rtol=1e-5, atol=0)):
                    bad += 1
            dt = time.perf_counter() - t0
            edis["cpu_baseline"] = {"kind": "port", "cores": 1, "value": nq / dt, "unit": "queries/s",
                                    "sample": f"{nq} of the queries, oracle port of solr.py (includes the GPU "
                                              "re-run used for the parity check, negligible)",
                                    "gpu_topk_mismatches": bad}
        upload()          # restore the term batch for the sections below
        del frame, title, tdev

    # ---- e2e_dense: the literal .score() drop-in, dense float32[N] to the host per query
    e2e_dense = None
    if rank == 0 and world == 1:
        from searcharray_b200.postings import _pool
        out = _pool.empty_f32(host.n_docs)
        nd = min(Q, 96)
        for i in range(3):
            _lib.check(L.sa_score_term(h, int(term_ids[i]), float(idf[i]), float(avgdl), K1, B, 0, _lib.ALL_BITS,
                                       _lib.p_f32(out)))
        t0 = time.perf_counter()
        for i in range(nd):
            _lib.check(L.sa_score_term(h, int(term_ids[i]), float(idf[i]), float(avgdl), K1, B, 0, _lib.ALL_BITS,
                                       _lib.p_f32(out)))
        dt = time.perf_counter() - t0
        e2e_dense = {"value": nd / dt, "unit": "queries/s", "d2h_bytes_per_query": int(host.n_docs * 4),
                     "note": "SearchArray.score drop-in: one sa_score_term call per query, pinned result vector"}

    # ---- cpu_baseline (rank 0, N=1): oracle port on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        idx, one = cpu_reference_runner(host, avgdl, args.docs, k)
        for t in np.unique(term_ids):
            idx.docfreq(int(t))
            idx.termfreqs(int(t))
        sample = term_ids[:min(Q, args.ref_sample)]
        threads = best_thread_count(one, sample, cores)
        tt, n = 0.0, 0
        while tt < 10.0 and n < 8:
            tt += run_cpu_sample(one, sample, threads)
            n += 1
        cpu = {"value": n * len(sample) / tt, "unit": "queries/s", "cores": threads, "host_cores": cores,
               "kind": "port",
               "sample": f"{n} x {len(sample)} of the step's queries, ThreadPool({threads}) = fastest of the "
                         "probed pool widths, warm tf cache, score over all N + argpartition top-k"}
        # parity spot-check of the GPU top-k against the oracle on the sample
        bad = 0
        for qi in range(min(16, len(sample))):
            ref = one(int(sample[qi]))
            s_ref = idx.score(int(sample[qi]), k1=K1, b=B)
            order = np.lexsort((np.arange(len(s_ref)), -s_ref.astype(np.float64)))[:k]
            if not np.array_equal(out_docs[qi], order.astype(np.uint32)):
                bad += 1
            del ref
        cpu["gpu_topk_mismatches_in_16"] = bad
        try:
            cpu["cold"] = cold_cpu_qps(idx, one, sample)
        except Exception as e:                      # informational; never fail the run for it
            cpu["cold"] = {"error": repr(e)}

    verify = None
    if rank == 0 and args.verify:
        from oracle import ops as oops, search as osearch
        from searcharray_b200 import synth
        from searcharray_b200.shard import shard_topk_keys, unpack_keys
        full, _, _ = (host, lo, hi) if world == 1 else synth.generate_shard(spec, 0, 1)
        bad = 0
        for qi in range(min(args.verify, Q)):
            t = int(term_ids[qi])
            ids, tfs = osearch.termfreqs_sparse(full.term_words(t))
            sc = tfs.copy()
            oops.bm25_score(sc, full.doc_lens[ids.astype(np.int64)], avgdl, float(idf[qi]), K1, B)
            wd, ws = unpack_keys(shard_topk_keys(ids, sc, k))
            if not (np.array_equal(wd, out_docs[qi]) and np.array_equal(ws.view(np.uint32), out_scores[qi].view(np.uint32))):
                bad += 1
        verify = {"queries_checked": min(args.verify, Q), "mismatches": bad}
        log("verify:", verify)

    if rank == 0:
        line = {
            "metric": "queries/sec (single-term BM25 + top-k) on 10M-doc synthetic MSMARCO",
            "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, Q),
            "clocks": clk,
            "e2e": {"value": e2e, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_value,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e_dense": e2e_dense,
            "phrase": phrase,
            "edismax": edis,
            "topk_overflow_reruns": int(overflow),
            "verify": verify,
        }
        print(json.dumps(line), flush=True)
    dev.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--ref-sample", type=int, default=192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--phrase-queries", type=int, default=256)
    ap.add_argument("--slop-queries", type=int, default=16)
    ap.add_argument("--edismax-queries", type=int, default=48)
    ap.add_argument("--verify", type=int, default=0,
                    help="rank 0 re-generates the FULL corpus and checks this many queries' global top-k "
                         "against the CPU oracle (parity of the sharded / all-gathered path)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        bench_reference(args, rank, world)
    else:
        bench_ours(args, rank, world)


if __name__ == "__main__":
    main()

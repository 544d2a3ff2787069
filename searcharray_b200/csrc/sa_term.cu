// sa_term.cu -- the term-at-a-time BM25 scan (the headline kernel).
//
// Replaces, fused into one launch per query batch:
//   popcount64_reduce   searcharray/roaringish/popcount.pyx:212-237  (tf by doc)
//   as_dense/scatter    searcharray/roaringish/roaringish_ops.pyx:84-98, scatter_assign.h:8-29
//   bm25_score          searcharray/bm25/bm25.pyx:11-41
//
// Design (B200).  The dense float32[N] score vector is cut into tiles of SA_TILE_DOCS docs; one
// CTA owns one (query, tile) and builds the tile in 32 KB of shared memory:
//   1. the slice [lo,hi) of the term's posting words whose docs fall in the tile comes from the
//      term's tile directory (two loads; built at upload for long lists) or, for short lists,
//      from a warp-cooperative 32-ary search;
//   2. the slice is streamed with coalesced 8-byte loads (4 windows in flight per thread).  Words
//      are sorted by doc, so the words of one doc are adjacent: the thread holding the FIRST word
//      of a doc ("head") adds the popcounts of the doc's run (neighbours via warp shuffle over
//      overlapping 32-lane windows), takes the doc's precomputed BM25 length norm (a gather per
//      pass, issued back to back; dense tiles stage the tile's norms with cp.async instead),
//      evaluates tf/(tf+norm)*idf with individually rounded operations (bit-identical to the
//      reference's x86-64 build) and stores the score into the shared tile.  No atomics, no work
//      for docs that do not contain the term;
//   3. the tile is flushed once with 16-byte streaming stores; while it passes through registers
//      every score >= a running, provably valid lower bound of the k-th best score is appended to
//      the query's top-k candidate list (sa_topk.cu), so the dense vector is never re-read.
// The grid is (queries, tiles) -- the query index runs fastest -- so that the CTAs resident on an
// SM at any time belong to MANY queries: dense terms (issue-bound CTAs) and sparse terms
// (store-bound CTAs) overlap, and a tile's norm sectors are shared in L2 by all queries.  With the
// tiles of one query back to back the same kernel was 17 % slower (profiles/README.md).
// HBM traffic: 8*W (words) + the 32 B sectors holding the 4*df norms + 4*N (scores), less whatever
// the queries of one launch share in L2.  History (profiles/): v1 evaluated BM25 for all 16 docs of
// every thread under divergence with shared-memory atomics and was instruction-bound at 21 % of
// the HBM roofline; variants that staged the postings with TMA bulk copies (cp.async.bulk +
// mbarrier), wrote zeros straight to HBM for sparse tiles, or pinned the norm table in L2 measured
// slower and were dropped.
#include <algorithm>
#include <type_traits>

#include "sa_term.cuh"

__device__ __forceinline__ bool payload_keep(u64 w, u64 lo, u64 hi) {
    // reference roaringish_ops.pyx:55: compares the UNSHIFTED masked word
    u64 v = w & SA_MSB_MASK;
    return v >= lo && v <= hi;
}

__device__ __forceinline__ float bm25_from_norm(float tf, float norm, float idf) {
    return __fmul_rn(__fdiv_rn(tf, __fadd_rn(tf, norm)), idf);
}

template <int MODE, bool ALL_DOCS, bool FILTER>
__global__ void __launch_bounds__(SA_TERM_THREADS, 6)
term_tile_kernel(const TermBatchArgs a) {
    __shared__ __align__(16) float s_out[SA_TILE_DOCS];
    __shared__ u32 s_range[2];
    __shared__ u32 s_top[(SA_TERM_THREADS / 32) * 8];
    __shared__ u32 s_ncand, s_tile_max;

    // grid = (queries, tiles): consecutive CTAs work on the SAME tile of different queries, so at any
    // moment an SM holds a mix of dense (issue-bound) and sparse (store-bound) terms, and the tile's
    // norm sectors are shared in L2 by all queries of the launch
    const u32 q = a.query_major ? blockIdx.y : blockIdx.x;
    const u32 tile = a.query_major ? blockIdx.x : blockIdx.y;
    const TermQuery tq = a.queries[q];
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 *__restrict__ words = a.words + tq.word_off;
    const u32 n_words = (u32)tq.n_words;
    const u32 tile_doc0 = tile * SA_TILE_DOCS;                       // local doc index
    const u32 tile_doc0_abs = (u32)a.doc_base + tile_doc0;           // as stored in the words

    // 1. zero the tile; posting slice [lo, hi) of this tile
#pragma unroll
    for (int i = 0; i < SA_TILE_DOCS / SA_TERM_THREADS / 4; i++)
        reinterpret_cast<float4 *>(s_out)[tid + i * SA_TERM_THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
    u32 lo, hi;
    bool quads = false;
    // tf-table path (CTA-uniform): the term has (doc, tf) records and nothing has to look inside the words
    const bool use_recs = !FILTER && !ALL_DOCS && a.recs != nullptr && tq.rec_off != SA_NO_DIR && tq.dir_off != SA_NO_DIR;
    if (use_recs) {
        const u32 *dir = a.rec_dir + tq.dir_off + tile;
        lo = __ldg(dir);
        hi = __ldg(dir + 1);
        // L2 prefetch of the records a LATER tile of this query will read: the grid is (queries, tiles) with the query
        // fastest, so tile + prefetch_tiles is dispatched about one generation of resident CTAs after this one; its
        // record loads -- the second of two dependent DRAM round trips (directory, then records) on a memory system
        // saturated with the dense rows' stores -- then hit L2.  One 128-byte line per thread covers any tile (<= 32 KB).
        const u32 pf_tile = tile + a.prefetch_tiles;
        if (a.prefetch_tiles && (u64)pf_tile * SA_TILE_DOCS < a.n_docs) {
            const u32 plo = __ldg(dir + a.prefetch_tiles), phi = __ldg(dir + a.prefetch_tiles + 1);
            const u32 pi = (plo & ~31u) + tid * 32u;
            if (pi < phi) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.recs + tq.rec_off + pi));
        }
        __syncthreads();
        // Four records per thread only when that still leaves >= k threads, in whole warps, holding a score: the
        // tile bound is the k-th largest of (at most 8 per warp) thread maxima, and with fewer than k of them it
        // degenerates to "keep everything" -- 129 records in 32 threads of ONE warp overflowed the 128 slots (the short
        // last tile of a 2.5M-doc shard).  16 * k records = 4 * k quads = k / 8 full warps of 8 published maxima.
        quads = hi - lo >= max(a.quad_min_recs, 16u * a.topk.k);
    } else if (tq.dir_off != SA_NO_DIR) {                             // CTA-uniform
        const u32 *dir = a.tile_dir + tq.dir_off + tile;
        lo = __ldg(dir);
        hi = __ldg(dir + 1);
        __syncthreads();
    } else {
        if (warp < 2) {
            u64 key = (u64)tile_doc0_abs + (warp ? SA_TILE_DOCS : 0);
            u64 r = warp_lower_bound_shifted(words, 0, n_words, key, SA_KEY_SHIFT);
            if (lane == 0) s_range[warp] = (u32)r;
        }
        __syncthreads();
        lo = s_range[0];
        hi = s_range[1];
    }

    // Dense tiles (SCORE mode): instead of one dependent norm gather per matching doc, the tile's
    // norms are staged in the score tile itself with asynchronous 16-byte copies (cp.async) issued
    // together with the first posting loads -- one coalesced 32 KB read in place of a second DRAM
    // round trip per pass.  A doc's head reads its norm from the tile and stores the score NEGATED:
    // norms are > 0 here and scores >= +0, so the sign bit tells a score (set) from a leftover norm
    // (clear), which the flush turns into 0.
    const bool staged_norm = MODE == TERM_MODE_SCORE && !ALL_DOCS &&
                             (hi - lo) >= (use_recs ? a.staged_norm_min_recs : a.staged_norm_min_words);
    bool norm_ready = !staged_norm;
    if (staged_norm) {
        const float4 *__restrict__ n4 = reinterpret_cast<const float4 *>(a.norm + tile_doc0);
#pragma unroll
        for (int i = 0; i < SA_TILE_DOCS / SA_TERM_THREADS / 4; i++) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(reinterpret_cast<float4 *>(s_out) + tid + i * SA_TERM_THREADS);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(n4 + tid + i * SA_TERM_THREADS) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // 2. stream the slice.  Each warp takes windows of 30 owned words and loads 32 (two look-ahead
    //    lanes), so "is the previous / next word the same doc?" is a register shuffle with no
    //    warp-edge special case.  Words are sorted by doc: the thread holding the FIRST word of a
    //    doc ("head") sums the doc's run (<= 3 words via shuffles, longer runs by look-ahead
    //    loads).  SA_TERM_UNROLL windows are loaded before any is processed, and the heads' norm
    //    gathers are issued back to back before any score is computed (memory-level parallelism).
    u32 my_max = 0;
    const float *__restrict__ norm = a.norm + tile_doc0;
    constexpr u32 OWN = 30;
    constexpr u32 WIN = (SA_TERM_THREADS / 32) * OWN;                  // words per CTA pass (240)
    auto windows = [&](auto unroll_tag, const u32 base) {
        constexpr int UN = decltype(unroll_tag)::value;
        u64 w[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) {
            const u32 i = base + (u * (SA_TERM_THREADS / 32) + warp) * OWN + lane;
            // look-ahead lanes may read into the next tile (another doc) but never past the list
            w[u] = (i < n_words && i < hi + 2) ? __ldg(words + i) : ~0ull;
        }
        u32 pk[UN];                                       // rel << 18 | tf, ~0 = not a head
        float nr[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) {
            const u32 s = base + (u * (SA_TERM_THREADS / 32) + warp) * OWN;   // window start (warp-uniform)
            const u32 i = s + lane;
            const u32 rel = (u32)(w[u] >> SA_KEY_SHIFT) - tile_doc0_abs;  // >= 2^27 for the ~0 filler
            u32 pc = (u32)__popcll(w[u] & SA_LSB_MASK);
            if (FILTER && !payload_keep(w[u], a.min_payload, a.max_payload)) pc = 0;
            const u32 packed = (rel << 5) | pc;                        // pc <= 18
            u32 prev = __shfl_up_sync(0xffffffffu, packed, 1);
            const u32 next = __shfl_down_sync(0xffffffffu, packed, 1);
            const u32 next2 = __shfl_down_sync(0xffffffffu, packed, 2);
            if (lane == 0) {
                prev = ~0u;                                           // s == lo: previous word is another tile's
                if (s > lo && s < hi) prev = ((u32)(__ldg(words + s - 1) >> SA_KEY_SHIFT) - tile_doc0_abs) << 5;
            }
            pk[u] = ~0u;
            nr[u] = 0.0f;
            const bool owned = lane < OWN && i < hi;
            if (owned && (prev >> 5) != rel && rel < SA_TILE_DOCS) {
                u32 tf = pc;
                if ((next >> 5) == rel) {
                    tf += next & 31u;
                    if ((next2 >> 5) == rel) {
                        tf += next2 & 31u;
                        for (u32 j = i + 3; j < n_words; j++) {        // runs of >= 4 words (rare)
                            const u64 w2 = __ldg(words + j);
                            if ((u32)(w2 >> SA_KEY_SHIFT) - tile_doc0_abs != rel) break;
                            if (!(FILTER && !payload_keep(w2, a.min_payload, a.max_payload))) tf += (u32)__popcll(w2 & SA_LSB_MASK);
                        }
                    }
                }
                pk[u] = (rel << 18) | tf;
                if (MODE == TERM_MODE_SCORE && !ALL_DOCS && tf && !staged_norm) nr[u] = __ldg(norm + rel);
            }
        }
        if (!norm_ready) {                                 // CTA-uniform: first pass of a staged tile
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            norm_ready = true;
        }
#pragma unroll
        for (int u = 0; u < UN; u++) {
            if (pk[u] != ~0u) {
                const u32 rel = pk[u] >> 18, tf = pk[u] & 0x3FFFFu;
                float v;
                if (MODE == TERM_MODE_TF || ALL_DOCS) {
                    v = (float)tf;
                } else {
                    v = 0.0f;
                    if (tf) {
                        const float nrm = staged_norm ? s_out[rel] : nr[u];
                        v = bm25_from_norm((float)tf, nrm, tq.idf);
                        my_max = max(my_max, __float_as_uint(v));
                    }
                }
                s_out[rel] = staged_norm ? -v : v;
            }
        }
    };
    if (use_recs) {
        // 2'. tf-table path: one u32 record per matching doc, (doc - tile_doc0) << 19 | tf, in doc order.  Every
        //     lane takes FOUR records with one 16-byte load (record runs start 16-byte aligned; the slice is
        //     widened to whole quads and the strangers masked); no run detection, no shuffles, no popcount.
        const u32 *__restrict__ recs = a.recs + tq.rec_off;
        if (quads) {                       // with a tile bound over thread maxima at most 4 * k docs reach it
            // quads => staged norms (quad_min_recs >= staged_norm_min_recs, launch_term_batch): no norm gathers here.
            // Two quads are loaded before either is processed (two 16-byte loads in flight per thread).
            constexpr int QU = 2;
            for (u32 base = lo & ~3u; base < hi; base += SA_TERM_THREADS * 4 * QU) {  // CTA-uniform trip count
                uint4 r4[QU];
    #pragma unroll
                for (int u = 0; u < QU; u++) {
                    const u32 i = base + (u * SA_TERM_THREADS + tid) * 4;
                    r4[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (i < hi) r4[u] = __ldg(reinterpret_cast<const uint4 *>(recs + i));
                }
                if (!norm_ready) {                             // CTA-uniform: first pass of a staged tile
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    __syncthreads();
                    norm_ready = true;
                }
    #pragma unroll
                for (int u = 0; u < QU; u++) {
                    const u32 i = base + (u * SA_TERM_THREADS + tid) * 4;
                    const u32 rr[4] = {r4[u].x, r4[u].y, r4[u].z, r4[u].w};
    #pragma unroll
                    for (int e = 0; e < 4; e++) {
                        if (i + e < lo || i + e >= hi) continue;
                        const u32 rel = rr[e] >> SA_REC_TF_BITS, tf = rr[e] & SA_REC_TF_MASK;
                        float v = (float)tf;
                        if (MODE == TERM_MODE_SCORE) {
                            v = 0.0f;
                            if (tf) {
                                v = bm25_from_norm((float)tf, staged_norm ? s_out[rel] : 1.0f, tq.idf);
                                my_max = max(my_max, __float_as_uint(v));
                            }
                        }
                        s_out[rel] = staged_norm ? -v : v;
                    }
                }
            }
        } else {
            // sparse tiles: ONE record per thread, so the threads' maxima (the tile bound of step 3) come from as
            // many distinct docs as possible
            for (u32 base = lo; base < hi; base += SA_TERM_THREADS) {                 // CTA-uniform trip count
                const u32 i = base + tid;
                const bool live = i < hi;
                const u32 r = live ? __ldg(recs + i) : 0u;
                float nrm = 1.0f;
                if (MODE == TERM_MODE_SCORE && !staged_norm && live) nrm = __ldg(norm + (r >> SA_REC_TF_BITS));
                if (!norm_ready) {                         // CTA-uniform: first pass of a staged tile
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    __syncthreads();
                    norm_ready = true;
                }
                if (live) {
                    const u32 rel = r >> SA_REC_TF_BITS, tf = r & SA_REC_TF_MASK;
                    float v = (float)tf;
                    if (MODE == TERM_MODE_SCORE) {
                        v = 0.0f;
                        if (tf) {
                            v = bm25_from_norm((float)tf, staged_norm ? s_out[rel] : nrm, tq.idf);
                            my_max = max(my_max, __float_as_uint(v));
                        }
                    }
                    s_out[rel] = staged_norm ? -v : v;
                }
            }
        }
    } else {
    // CTA-uniform schedule: big slices in 4-window passes, the remainder (and small tiles) in
    // single-window passes so sparse tiles do not pay for empty windows.
    u32 base = lo;
    while (base < hi && hi - base > WIN) {
        windows(std::integral_constant<int, SA_TERM_UNROLL>{}, base);
        base += WIN * SA_TERM_UNROLL;
    }
    while (base < hi) {
        windows(std::integral_constant<int, 1>{}, base);
        base += WIN;
    }
    }

    // 3. top-k.  A tile with no more words than candidate slots needs no bound: every positive
    //    score fits.  Otherwise each warp publishes its largest thread maxima and every warp
    //    derives the same tile bound; scores >= bound are this tile's candidates.
    const u32 k = a.topk.k;
    // CTA-uniform (<= k postings: all fit).  Keeping every positive score of tiles with up to `slots` postings instead
    // was measured slower (df/N 1e-2: 9.9 vs 9.2 us/query): candidates cost more than the bound.
    const bool need_bound = k && (hi - lo) > k;
    // threads that can hold a score: one per record / posting word, or one per quad of records on the dense tf-table path
    const u32 M = tile_bound_width(k, quads ? (hi - lo) / 4u : (hi - lo));
    if (need_bound) {
        u32 v = my_max;
        for (u32 r = 0; r < M; r++) {
            u32 m = warp_pop_max(v);
            if (lane == r) s_top[warp * 8 + r] = m;
        }
    }
    if (k && tid == 0) { s_ncand = 0; s_tile_max = 0; }
    __syncthreads();
    float thr_f = 0.0f;
    if (k) {
        u32 thr = 1u;                                   // >= 1: skip zeros (scores are >= +0.0)
        if (need_bound) thr = max(cta_kth_bound(s_top, k, M == 8u), 1u);
        thr_f = __uint_as_float(thr);
    }
    u64 *__restrict__ my_cand = nullptr;
    if (k) my_cand = a.topk.tile_cand + ((u64)q * a.topk.n_tiles + tile) * a.topk.slots;

    // 4. flush the tile: 16-byte streaming stores (the padded buffer makes the tile always in bounds)
    u32 cand_max = 0;
    float4 *__restrict__ out4 = reinterpret_cast<float4 *>(a.out + (u64)q * a.out_stride + tile_doc0);
#pragma unroll
    for (int jj = 0; jj < SA_TILE_DOCS / SA_TERM_THREADS / 4; jj++) {
        const unsigned g = tid + jj * SA_TERM_THREADS;
        float4 v = reinterpret_cast<const float4 *>(s_out)[g];
        if (staged_norm) {                                  // sign set: a (negated) score; clear: a leftover norm
            v.x = __float_as_int(v.x) < 0 ? -v.x : 0.0f;
            v.y = __float_as_int(v.y) < 0 ? -v.y : 0.0f;
            v.z = __float_as_int(v.z) < 0 ? -v.z : 0.0f;
            v.w = __float_as_int(v.w) < 0 ? -v.w : 0.0f;
        }
        if (ALL_DOCS && MODE == TERM_MODE_SCORE) {
            // bm25.pyx:20-25 over EVERY doc (NaN / inf / -0.0 cases of exotic parameters)
            Bm25Params p = a.bm25;
            p.idf = tq.idf;
            const u64 d = (u64)tile_doc0 + (u64)g * 4;
            const float *dl = a.doc_lens + d;
            v.x = (d + 0 < a.n_docs) ? bm25_one(v.x, dl[0], p) : 0.0f;
            v.y = (d + 1 < a.n_docs) ? bm25_one(v.y, dl[1], p) : 0.0f;
            v.z = (d + 2 < a.n_docs) ? bm25_one(v.z, dl[2], p) : 0.0f;
            v.w = (d + 3 < a.n_docs) ? bm25_one(v.w, dl[3], p) : 0.0f;
        }
        __stcs(out4 + g, v);
        if (k) {
            // NaN compares false; negatives are below thr_f > 0
            if ((v.x >= thr_f) | (v.y >= thr_f) | (v.z >= thr_f) | (v.w >= thr_f)) {
                const float vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (vs[e] >= thr_f) {
                        u32 slot = atomicAdd(&s_ncand, 1u);           // shared-memory atomic
                        if (slot < a.topk.slots)
                            my_cand[slot] = ((u64)__float_as_uint(vs[e]) << 32) |
                                            (u64)(0xFFFFFFFFu - (tile_doc0 + g * 4 + e));
                        cand_max = max(cand_max, __float_as_uint(vs[e]));
                    }
                }
            }
        }
    }
    if (k) {
        if (cand_max) atomicMax(&s_tile_max, cand_max);
        __syncthreads();
        if (s_ncand > a.topk.slots && !(ALL_DOCS && MODE == TERM_MODE_SCORE))    // CTA-uniform: ties at the bound
            tile_collect_ties_retry(s_out, staged_norm, __float_as_uint(thr_f), a.topk, my_cand, tile_doc0, s_top,
                                    &s_ncand, &s_tile_max);
        if (tid == 0) {
            const u32 n = s_ncand;
            const u64 t_idx = (u64)q * a.topk.n_tiles + tile;
            a.topk.tile_cnt[t_idx] = min(n, a.topk.slots);
            a.topk.tile_max[t_idx] = s_tile_max;
            if (n > a.topk.slots) a.topk.overflow[q] = 1u;
        }
    }
}

// Top-k candidate collection for an already materialised dense score vector (phrase queries:
// their kernel scatters sparse matches, so the collector runs as a separate tile scan).  Same
// outputs as step 3/4 of term_tile_kernel: per-tile candidate slots, count and maximum.
template <bool SCORE>
__global__ void __launch_bounds__(SA_TERM_THREADS)
dense_topk_tiles_kernel(float *__restrict__ dense, u64 stride, u32 row0, const TopkCtx t,
                        const float *__restrict__ norm, const float *__restrict__ row_idf) {
    __shared__ u32 s_top[(SA_TERM_THREADS / 32) * 8];
    __shared__ u32 s_ncand, s_tile_max;
    const u32 q = blockIdx.y + row0;
    const u32 tile = blockIdx.x;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 tile_doc0 = tile * SA_TILE_DOCS;
    const u32 k = t.k;
    constexpr int NV = SA_TILE_DOCS / SA_TERM_THREADS / 4;
    float4 *__restrict__ src = reinterpret_cast<float4 *>(dense + (u64)q * stride + tile_doc0);
    float4 v[NV];
    u32 my_max = 0;
    const float idf = SCORE ? row_idf[blockIdx.y] : 0.0f;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        v[j] = __ldcs(src + tid + j * SA_TERM_THREADS);
        float vs[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
        if (SCORE) {
            // raw counts -> BM25 (bm25.pyx:20-25; a zero count scores +0.0 for ordinary parameters)
            if ((vs[0] != 0.0f) | (vs[1] != 0.0f) | (vs[2] != 0.0f) | (vs[3] != 0.0f)) {
                const u32 d0 = tile_doc0 + (tid + j * SA_TERM_THREADS) * 4;
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (vs[e] != 0.0f) vs[e] = bm25_from_norm(vs[e], __ldg(norm + d0 + e), idf);
                v[j] = make_float4(vs[0], vs[1], vs[2], vs[3]);
                __stcs(src + tid + j * SA_TERM_THREADS, v[j]);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (vs[e] > 0.0f) my_max = max(my_max, __float_as_uint(vs[e]));
    }
    const u32 M = 8u;
    u32 mv = my_max;
    for (u32 r = 0; r < M; r++) {
        u32 m = warp_pop_max(mv);
        if (lane == r) s_top[warp * 8 + r] = m;
    }
    if (tid == 0) { s_ncand = 0; s_tile_max = 0; }
    __syncthreads();
    const float thr_f = __uint_as_float(max(cta_kth_bound(s_top, k, true), 1u));
    u64 *__restrict__ my_cand = t.tile_cand + ((u64)q * t.n_tiles + tile) * t.slots;
    u32 cand_max = 0;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const float vs[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (vs[e] >= thr_f) {
                u32 slot = atomicAdd(&s_ncand, 1u);
                if (slot < t.slots)
                    my_cand[slot] = ((u64)__float_as_uint(vs[e]) << 32) |
                                    (u64)(0xFFFFFFFFu - (tile_doc0 + (tid + j * SA_TERM_THREADS) * 4 + e));
                cand_max = max(cand_max, __float_as_uint(vs[e]));
            }
        }
    }
    if (cand_max) atomicMax(&s_tile_max, cand_max);
    __syncthreads();
    if (tid == 0) {
        const u32 n = s_ncand;
        const u64 t_idx = (u64)q * t.n_tiles + tile;
        t.tile_cnt[t_idx] = min(n, t.slots);
        t.tile_max[t_idx] = s_tile_max;
        if (n > t.slots) t.overflow[q] = 1u;
    }
}

int launch_dense_topk_tiles(sa_index *ix, float *dense, u64 stride, u32 row0, u32 n_rows, const TopkCtx &t,
                            const float *d_row_idf) {
    if (n_rows == 0 || t.n_tiles == 0) return SA_OK;
    dim3 grid(t.n_tiles, n_rows);
    KernelTimer tm(ix, 1);
    if (d_row_idf) dense_topk_tiles_kernel<true><<<grid, SA_TERM_THREADS, 0, ix->stream>>>(dense, stride, row0, t, ix->d_norm, d_row_idf);
    else dense_topk_tiles_kernel<false><<<grid, SA_TERM_THREADS, 0, ix->stream>>>(dense, stride, row0, t, nullptr, nullptr);
    SA_CUDA(cudaGetLastError());
    tm.stop();
    ix->stats.topk_kernel_launches++;
    ix->stats.total_launches++;
    return SA_OK;
}

// per-doc BM25 length norm, the inner part of bm25.pyx:21-23 with the same rounding sequence
__global__ void norm_kernel(const float *__restrict__ dl, float *__restrict__ norm, u64 n, u64 n_pad,
                            Bm25Params p) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    float v = 0.0f;
    if (i < n) v = __fmul_rn(p.k1, __fadd_rn(p.one_minus_b, __fmul_rn(p.b, __fdiv_rn(dl[i], p.avg_doc_len))));
    norm[i] = v;
}

int sa_ensure_norm(sa_index *ix, float k1, float b, float avg_doc_len) {
    if (ix->norm_valid && ix->norm_k1 == k1 && ix->norm_b == b && ix->norm_avgdl == avg_doc_len) return SA_OK;
    const u64 n_pad = (ix->n_docs + SA_TILE_DOCS - 1) / SA_TILE_DOCS * SA_TILE_DOCS;
    if (!ix->d_norm) {
        SA_CUDA(cudaMalloc(&ix->d_norm, std::max<u64>(n_pad, 1) * sizeof(float)));
        ix->device_bytes += n_pad * sizeof(float);
    }
    Bm25Params p;
    p.idf = 0; p.avg_doc_len = avg_doc_len; p.k1 = k1; p.b = b; p.one_minus_b = 1 - b; p.sparse_ok = 1;
    if (n_pad) {
        norm_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, ix->stream>>>(ix->d_doc_lens, ix->d_norm, ix->n_docs, n_pad, p);
        SA_CUDA(cudaGetLastError());
        ix->stats.total_launches++;
    }
    ix->norm_k1 = k1; ix->norm_b = b; ix->norm_avgdl = avg_doc_len;
    ix->norm_valid = true;
    return SA_OK;
}

int launch_term_batch(sa_index *ix, const TermBatchArgs &a_in, u32 n_queries) {
    if (n_queries == 0 || a_in.n_docs == 0) return SA_OK;
    TermBatchArgs a = a_in;
    {
        // tuning knobs, read per launch (a getenv costs nanoseconds; tools/term_buckets.py sweeps them in one process)
        const char *e;
        a.staged_norm_min_words = (e = getenv("SA_STAGED_NORM_MIN_WORDS")) ? (u32)atol(e) : SA_STAGED_NORM_MIN_WORDS;
        a.staged_norm_min_recs = (e = getenv("SA_STAGED_NORM_MIN_RECS")) ? (u32)atol(e) : SA_STAGED_NORM_MIN_RECS;
        a.quad_min_recs = (e = getenv("SA_TERM_QUAD_MIN_RECS")) ? (u32)atol(e) : SA_TERM_QUAD_MIN_RECS;
        a.prefetch_tiles = (e = getenv("SA_TERM_PREFETCH_TILES")) ? (u32)atol(e) : SA_TERM_PREFETCH_TILES;
        a.quad_min_recs = std::max(a.quad_min_recs, a.staged_norm_min_recs);   // the quad path reads norms from the staged tile only
    }
    a.tile_dir = ix->d_tile_dir;
    a.recs = (a.words == ix->d_words) ? ix->d_recs : nullptr;     // the tf table describes the index's own lists only
    a.rec_dir = ix->d_rec_dir;
    a.norm = ix->d_norm;
    const bool sparse_score = (a.mode == TERM_MODE_SCORE) && a.bm25.sparse_ok;
    if (sparse_score) {
        int rc = sa_ensure_norm(ix, a.bm25.k1, a.bm25.b, a.bm25.avg_doc_len);
        if (rc) return rc;
        a.norm = ix->d_norm;
    }
    const unsigned n_tiles = (unsigned)((a.n_docs + SA_TILE_DOCS - 1) / SA_TILE_DOCS);
    static const bool env_qmajor = getenv("SA_TERM_QUERY_MAJOR") && atoi(getenv("SA_TERM_QUERY_MAJOR")) != 0;
    a.query_major = (env_qmajor || n_tiles > 65535) ? 1 : 0;
    dim3 grid = a.query_major ? dim3(n_tiles, n_queries) : dim3(n_queries, n_tiles);
    dim3 block(SA_TERM_THREADS);
    KernelTimer t(ix, 0);
    if (a.mode == TERM_MODE_TF) {
        if (a.filter) term_tile_kernel<TERM_MODE_TF, false, true><<<grid, block, 0, ix->stream>>>(a);
        else term_tile_kernel<TERM_MODE_TF, false, false><<<grid, block, 0, ix->stream>>>(a);
    } else if (sparse_score) {
        if (a.filter) term_tile_kernel<TERM_MODE_SCORE, false, true><<<grid, block, 0, ix->stream>>>(a);
        else term_tile_kernel<TERM_MODE_SCORE, false, false><<<grid, block, 0, ix->stream>>>(a);
    } else {
        if (a.filter) term_tile_kernel<TERM_MODE_SCORE, true, true><<<grid, block, 0, ix->stream>>>(a);
        else term_tile_kernel<TERM_MODE_SCORE, true, false><<<grid, block, 0, ix->stream>>>(a);
    }
    SA_CUDA(cudaGetLastError());
    t.stop();
    ix->stats.term_kernel_launches++;
    ix->stats.term_kernel_queries += n_queries;
    ix->stats.total_launches++;
    return SA_OK;
}

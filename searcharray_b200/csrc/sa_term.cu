// sa_term.cu -- the term-at-a-time BM25 scan (the headline kernel).
//
// Replaces, fused into one launch per query batch:
//   popcount64_reduce   searcharray/roaringish/popcount.pyx:212-237  (tf by doc)
//   as_dense/scatter    searcharray/roaringish/roaringish_ops.pyx:84-98, scatter_assign.h:8-29
//   bm25_score          searcharray/bm25/bm25.pyx:11-41
//
// Design (B200): the dense float32[N] score vector is cut into tiles of SA_TILE_DOCS docs.
// One CTA owns one (tile, query): it finds the slice of the term's posting words whose doc
// ids fall in the tile (warp-cooperative 32-ary search; words are sorted by doc id), streams
// that slice with coalesced 8-byte loads, accumulates popcounts per doc with shared-memory
// atomics into a 16 KB tile, then converts the tile to BM25 scores (gathering doc_lens only
// where tf > 0 -- the hardware fetches only the touched 32 B sectors) and writes it out
// once, with 16-byte coalesced stores.  HBM traffic = 8*W (words) + <=4*df.. (doc_lens
// sectors) + 4*N (scores), i.e. the algorithmic minimum of SURVEY.md section 8d.
// The epilogue optionally feeds the top-k collector (sa_topk.cu) from registers so the
// dense vector is never re-read.
#include "sa_term.cuh"

__device__ __forceinline__ bool payload_keep(u64 w, u64 lo, u64 hi) {
    // reference roaringish_ops.pyx:55: compares the UNSHIFTED masked word
    u64 v = w & SA_MSB_MASK;
    return v >= lo && v <= hi;
}

// Sort 32 values (one per lane) descending across the warp (bitonic network).
__device__ __forceinline__ u32 warp_sort_desc(u32 v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            u32 o = __shfl_xor_sync(0xffffffffu, v, j);
            bool up = ((lane & k) == 0);          // descending block
            bool lower = ((lane & j) == 0);
            u32 mx = v > o ? v : o, mn = v > o ? o : v;
            v = (up == lower) ? mx : mn;
        }
    }
    return v;
}

template <int MODE, bool ALL_DOCS>
__global__ void __launch_bounds__(SA_TERM_THREADS)
term_tile_kernel(const TermBatchArgs a) {
    __shared__ u32 s_cnt[SA_TILE_DOCS];
    __shared__ u64 s_range[2];
    __shared__ u32 s_warp_bound[SA_TERM_THREADS / 32];

    const u32 q = blockIdx.y;
    const u64 tile = blockIdx.x;
    const TermQuery tq = a.queries[q];
    const u64 tile_doc0 = tile * SA_TILE_DOCS;                // local doc index of the tile start
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 *__restrict__ words = a.words + tq.word_off;

    // 1. zero the tile, and (warps 0/1) find the posting slice [lo, hi) of this tile
#pragma unroll
    for (int i = 0; i < SA_TILE_DOCS / SA_TERM_THREADS / 4; i++)
        reinterpret_cast<uint4 *>(s_cnt)[tid + i * SA_TERM_THREADS] = make_uint4(0, 0, 0, 0);
    if (warp < 2) {
        u64 key = a.doc_base + tile_doc0 + (warp ? SA_TILE_DOCS : 0);
        u64 r = warp_lower_bound_shifted(words, 0, tq.n_words, key, SA_KEY_SHIFT);
        if (lane == 0) s_range[warp] = r;
    }
    __syncthreads();
    const u64 lo = s_range[0], hi = s_range[1];

    // 2. stream the slice: tf[doc] += popcount(payload)
    {
        const u64 base_doc = a.doc_base + tile_doc0;
        u64 i = lo + tid;
        // 4 independent loads in flight per thread
        for (; i + 3 * SA_TERM_THREADS < hi; i += 4 * SA_TERM_THREADS) {
            u64 w0 = ld_stream_u64(words + i);
            u64 w1 = ld_stream_u64(words + i + SA_TERM_THREADS);
            u64 w2 = ld_stream_u64(words + i + 2 * SA_TERM_THREADS);
            u64 w3 = ld_stream_u64(words + i + 3 * SA_TERM_THREADS);
            u64 ws[4] = {w0, w1, w2, w3};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                u64 w = ws[j];
                if (a.filter && !payload_keep(w, a.min_payload, a.max_payload)) continue;
                u64 d = (w >> SA_KEY_SHIFT) - base_doc;
                if (d < SA_TILE_DOCS) atomicAdd(&s_cnt[d], (u32)__popcll(w & SA_LSB_MASK));
            }
        }
        for (; i < hi; i += SA_TERM_THREADS) {
            u64 w = ld_stream_u64(words + i);
            if (a.filter && !payload_keep(w, a.min_payload, a.max_payload)) continue;
            u64 d = (w >> SA_KEY_SHIFT) - base_doc;
            if (d < SA_TILE_DOCS) atomicAdd(&s_cnt[d], (u32)__popcll(w & SA_LSB_MASK));
        }
    }
    __syncthreads();

    // 3. epilogue: tf -> score, one coalesced 16 B store per 4 docs
    Bm25Params p = a.bm25;
    p.idf = tq.idf;
    float *__restrict__ out = a.out + (u64)q * a.out_stride + tile_doc0;
    const float *__restrict__ dls = a.doc_lens + tile_doc0;
    const u64 docs_left = a.n_docs > tile_doc0 ? a.n_docs - tile_doc0 : 0;   // valid docs in tile

    float sc[4][4];
    u32 my_max = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const unsigned g = tid + j * SA_TERM_THREADS;       // float4 group within the tile
        uint4 c = reinterpret_cast<const uint4 *>(s_cnt)[g];
        u32 cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const u64 d = (u64)g * 4 + e;
            float v;
            if (MODE == TERM_MODE_TF) {
                v = (float)cc[e];
            } else if (ALL_DOCS) {
                v = (d < docs_left) ? bm25_one((float)cc[e], dls[d], p) : 0.0f;
            } else {
                v = 0.0f;
                if (cc[e] != 0 && d < docs_left) v = bm25_one((float)cc[e], __ldg(dls + d), p);
            }
            sc[j][e] = v;
            if (d < docs_left) {
                u32 bits = __float_as_uint(v);
                // only positive finite-or-inf scores are top-k candidates (NaN/negative ignored)
                if (v > 0.0f && bits > my_max) my_max = bits;
            }
        }
        // padded buffer: the whole tile is always in bounds
        __stcs(reinterpret_cast<float4 *>(out) + g, make_float4(sc[j][0], sc[j][1], sc[j][2], sc[j][3]));
    }

    // 4. optional: feed the top-k collector from registers
    if (a.topk.k == 0) return;
    const u32 k = a.topk.k;
    // a valid lower bound on the k-th best score: the k-th largest of 32 lane maxima
    // (32 distinct docs).  0 when the warp holds fewer than k positive lanes.
    u32 sorted = warp_sort_desc(my_max);
    u32 wb = __shfl_sync(0xffffffffu, sorted, (k - 1) & 31);
    if (k > 32) wb = 0;
    if (lane == 0) s_warp_bound[warp] = wb;
    __syncthreads();
    u32 cta_bound = 0;
#pragma unroll
    for (int w = 0; w < SA_TERM_THREADS / 32; w++) cta_bound = max(cta_bound, s_warp_bound[w]);
    u32 thr = *((volatile u32 *)(a.topk.thr_bits + q));
    if (cta_bound > thr) {
        if (tid == 0) atomicMax(a.topk.thr_bits + q, cta_bound);
        thr = cta_bound;
    }
    // count my passing values, warp-aggregate one atomicAdd
    u32 npass = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const u64 d = (u64)(tid + j * SA_TERM_THREADS) * 4 + e;
            u32 bits = __float_as_uint(sc[j][e]);
            if (sc[j][e] > 0.0f && bits >= thr && d < docs_left) npass++;
        }
    unsigned any = __ballot_sync(0xffffffffu, npass != 0);
    if (any == 0) return;
    u32 incl = npass;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    u32 total = __shfl_sync(0xffffffffu, incl, 31);
    u32 base = 0;
    if (lane == 0) base = atomicAdd(a.topk.count + q, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    u32 slot = base + incl - npass;
    u64 *cand = a.topk.cand + (u64)q * a.topk.cap;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const u64 d = (u64)(tid + j * SA_TERM_THREADS) * 4 + e;
            u32 bits = __float_as_uint(sc[j][e]);
            if (sc[j][e] > 0.0f && bits >= thr && d < docs_left) {
                if (slot < a.topk.cap)
                    cand[slot] = ((u64)bits << 32) | (u64)(0xFFFFFFFFu - (u32)(tile_doc0 + d));
                slot++;
            }
        }
}

int launch_term_batch(sa_index *ix, const TermBatchArgs &a, u32 n_queries) {
    if (n_queries == 0 || a.n_docs == 0) return SA_OK;
    dim3 grid((unsigned)((a.n_docs + SA_TILE_DOCS - 1) / SA_TILE_DOCS), n_queries);
    dim3 block(SA_TERM_THREADS);
    KernelTimer t(ix, 0);
    if (a.mode == TERM_MODE_TF)
        term_tile_kernel<TERM_MODE_TF, false><<<grid, block, 0, ix->stream>>>(a);
    else if (a.bm25.sparse_ok)
        term_tile_kernel<TERM_MODE_SCORE, false><<<grid, block, 0, ix->stream>>>(a);
    else
        term_tile_kernel<TERM_MODE_SCORE, true><<<grid, block, 0, ix->stream>>>(a);
    SA_CUDA(cudaGetLastError());
    t.stop();
    ix->stats.term_kernel_launches++;
    ix->stats.term_kernel_queries += n_queries;
    ix->stats.total_launches++;
    return SA_OK;
}

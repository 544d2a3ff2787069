// sa_span.cu -- phrase matching with slop > 0 ("span search").
//
// Replaces (reference paths relative to softwaredoug/searcharray):
//   span_search / _intersect_all            searcharray/phrase/spans.py:71-187
//   _span_freqs, _compact_spans, _collect_spans, ActiveSpans   searcharray/roaringish/spans.pyx:70-319
//
// Phase 1 (span_candidates_kernel, one CTA per query) restates _intersect_all as a membership test.
// With A = headers of term 0 and B_k = headers of term k (header = doc|block), the reference keeps,
// for every term, the words whose header lies in  H = L u R u (L - 1) u (R + 1)  where
//   L = AND_k [ (A(x)&B_k(x)) | (B_k(x)&A(x-1)) | (A(x)&B_k(x-1)) ]
//   R = AND_k [ (A(x)&B_k(x)) | (A(x)&B_k(x+1)) | (B_k(x)&A(x+1)) ]
// (the merges / intersects / adjacents of spans.py:79-112 reduce to these presence tests because
// they are applied with the header mask and set semantics).  Every x in H has each term within two
// blocks, so the candidates are enumerated, already sorted and unique, from the shortest list:
// word s (header hs) emits hs-2..hs+2, each only if no earlier word of that list covers it.
// Phase 2 (span_groups_kernel) replays _span_freqs.  The reference walks all terms "up to the next
// doc change" in lock step, i.e. iteration i consumes the i-th DOC GROUP of every term's sliced list
// (normally the same doc; the lists can be misaligned, and then this pairing is what defines the
// result).  Iterations are independent: one warp runs one iteration with the <= 512-entry span table
// in shared memory, lanes sharing the "extend every live span" loop, forks appended in order through
// ballots.  Counts are accumulated per `last_key` like the reference's Counter.
#include <algorithm>

#include "sa_phrase.cuh"
#include "sa_term.cuh"

#define SPAN_CAP 512
#define SPAN_WARPS 4
#define CAND_THREADS 256

struct SpanQuery {
    u32 n_terms;
    u32 slop;
    u32 shortest;                       // index of the shortest list (candidate generator)
    u32 pad;
    u64 off[SA_MAX_PHRASE_TERMS];       // term lists in d_words
    u64 len[SA_MAX_PHRASE_TERMS];
    u64 s_off[SA_MAX_PHRASE_TERMS];     // sliced-list region of term t in the word arena
    u64 g_off[SA_MAX_PHRASE_TERMS];     // group-start region of term t in the u32 arena
    u64 s_cap[SA_MAX_PHRASE_TERMS];
};

struct SpanCounts {                      // written by phase 1, read by phase 2
    u32 n_sliced[SA_MAX_PHRASE_TERMS];
    u32 n_groups[SA_MAX_PHRASE_TERMS];
    u32 overflow;
    u32 undefined;                       // span-table overflows the reference leaves undefined
};

struct SpanArgs {
    const u64 *words;
    const SpanQuery *queries;
    SpanCounts *counts;
    u64 *word_arena;
    u32 *group_arena;
    float *out;                          // [Q][out_stride] pre-zeroed: out[last_key - doc_base] += count
    u64 out_stride;
    u64 n_docs, doc_base;
};

__device__ __forceinline__ u64 lb_hdr(const u64 *__restrict__ a, u64 n, u64 target) {
    u64 lo = 0, hi = n;
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        if ((a[mid] & SA_HDR_MASK) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ u32 block_scan_excl(u32 v, u32 *warp_sums, u32 &total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    u32 base = 0, tot = 0;
    for (int w = 0; w < CAND_THREADS / 32; w++) {
        u32 s = warp_sums[w];
        if (w < (int)warp) base += s;
        tot += s;
    }
    total = tot;
    return base + incl - v;
}

// ---------------------------------------------------------------------------- phase 1
__global__ void __launch_bounds__(CAND_THREADS)
span_candidates_kernel(const SpanArgs a) {
    __shared__ u32 s_warp[CAND_THREADS / 32];
    __shared__ u32 s_count[SA_MAX_PHRASE_TERMS];
    __shared__ u32 s_last_doc[SA_MAX_PHRASE_TERMS], s_groups[SA_MAX_PHRASE_TERMS];
    const u32 q = blockIdx.x;
    const SpanQuery &sq = a.queries[q];
    const u32 n = sq.n_terms;
    const unsigned tid = threadIdx.x;
    if (tid < SA_MAX_PHRASE_TERMS) { s_count[tid] = 0; s_groups[tid] = 0; s_last_doc[tid] = 0xFFFFFFFFu; }
    __syncthreads();
    const u64 *S = a.words + sq.off[sq.shortest];
    const u64 nS = sq.len[sq.shortest];
    constexpr u32 PER = CAND_THREADS / 5;                 // 51 generator words per pass, 5 candidates each

    for (u64 base = 0; base < nS; base += PER) {          // CTA-uniform
        const u32 j = tid / 5;
        const int delta = (int)(tid % 5) - 2;
        const u64 si = base + j;
        bool cand = (tid < PER * 5) && si < nS;
        u64 x = 0;
        if (cand) {
            const u64 hs = S[si] & SA_HDR_MASK;
            if (delta < 0 && hs < (u64)(-delta) * SA_ONE_BLOCK) cand = false;
            else x = hs + (u64)((i64)delta * (i64)SA_ONE_BLOCK);
            // emitted by the FIRST generator word within two blocks of x
            if (cand && si > 0) {
                const u64 hp = S[si - 1] & SA_HDR_MASK;
                if (hp + 2 * SA_ONE_BLOCK >= x) cand = false;
            }
        }
        // presence of every term at x-1, x, x+1 (bit 0: x-1, bit 1: x, bit 2: x+1)
        u32 pres[SA_MAX_PHRASE_TERMS];
        u64 at_x[SA_MAX_PHRASE_TERMS];
        bool keep = false;
        if (cand) {
            const bool has_m1 = x >= SA_ONE_BLOCK;
            for (u32 t = 0; t < n; t++) {
                const u64 *lst = a.words + sq.off[t];
                const u64 len = sq.len[t];
                u64 p = lb_hdr(lst, len, has_m1 ? x - SA_ONE_BLOCK : x);
                u32 m = 0;
                at_x[t] = 0;
                for (int r = 0; r < 3 && p < len; r++) {
                    const u64 w = lst[p];
                    const u64 h = w & SA_HDR_MASK;
                    if (has_m1 && h == x - SA_ONE_BLOCK) { m |= 1u; p++; }
                    else if (h == x) { m |= 2u; at_x[t] = w; p++; }
                    else if (h == x + SA_ONE_BLOCK) { m |= 4u; p++; }
                    else break;
                }
                pres[t] = m;
            }
            // L(y), R(y) for y in {x-1, x, x+1} as far as the presence window allows
            auto A = [&](int o) { return (pres[0] >> (o + 1)) & 1u; };       // o in {-1,0,1}
            bool Lx = true, Rx = true, Lx1 = true, Rxm1 = true;
            for (u32 k = 1; k < n; k++) {
                auto Bk = [&](int o) { return (pres[k] >> (o + 1)) & 1u; };
                Lx &= (A(0) & Bk(0)) | (Bk(0) & A(-1)) | (A(0) & Bk(-1));
                Rx &= (A(0) & Bk(0)) | (A(0) & Bk(1)) | (Bk(0) & A(1));
                Lx1 &= (A(1) & Bk(1)) | (Bk(1) & A(0)) | (A(1) & Bk(0));          // L(x+1)
                Rxm1 &= (A(-1) & Bk(-1)) | (A(-1) & Bk(0)) | (Bk(-1) & A(0));      // R(x-1)
            }
            keep = Lx | Rx | Lx1 | Rxm1;
        }
        // ordered append of the kept words, term by term
        for (u32 t = 0; t < n; t++) {
            const bool put = keep && (pres[t] & 2u);
            u32 total;
            u32 off = block_scan_excl(put ? 1u : 0u, s_warp, total);
            const u32 cnt0 = s_count[t];
            if (put) {
                if (cnt0 + off < sq.s_cap[t]) a.word_arena[sq.s_off[t] + cnt0 + off] = at_x[t];
                else a.counts[q].overflow = 1;
            }
            __syncthreads();
            if (tid == 0) s_count[t] = cnt0 + total;
            __syncthreads();
        }
    }
    if (tid < n) a.counts[q].n_sliced[tid] = min(s_count[tid], (u32)sq.s_cap[tid]);
}

// doc groups of every sliced list (runs after either candidate kernel)
__global__ void __launch_bounds__(CAND_THREADS)
span_groups_build_kernel(const SpanArgs a) {
    __shared__ u32 s_warp[CAND_THREADS / 32];
    __shared__ u32 s_groups;
    const u32 q = blockIdx.x;
    const SpanQuery &sq = a.queries[q];
    const unsigned tid = threadIdx.x;
    for (u32 t = 0; t < sq.n_terms; t++) {
        const u32 cnt = a.counts[q].n_sliced[t];
        const u64 *sl = a.word_arena + sq.s_off[t];
        if (tid == 0) s_groups = 0;
        __syncthreads();
        for (u32 base = 0; base < cnt; base += CAND_THREADS) {
            const u32 i = base + tid;
            bool start = false;
            if (i < cnt) start = (i == 0) || ((sl[i] >> SA_KEY_SHIFT) != (sl[i - 1] >> SA_KEY_SHIFT));
            u32 total;
            u32 off = block_scan_excl(start ? 1u : 0u, s_warp, total);
            const u32 g0 = s_groups;
            if (start) a.group_arena[sq.g_off[t] + g0 + off] = i;
            __syncthreads();
            if (tid == 0) s_groups = g0 + total;
            __syncthreads();
        }
        if (tid == 0) {
            a.group_arena[sq.g_off[t] + s_groups] = cnt;          // sentinel: end of the last group
            a.counts[q].n_groups[t] = s_groups;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------- phase 1, literal (rare) variant
// When every term has a word at header 0 (doc 0, block 0) the reference's `last_lhs_headers - 1
// block` (spans.py:104) underflows on its first element; the merges and the galloping slice that
// follow then run on a list that is no longer sorted, and WHICH candidate words survive depends on
// the exact pointer walk.  That is deterministic, so it is replayed literally -- one thread, the
// reference's own sequence of galloping intersects / adjacents / merges (the restatement follows
// searcharray/roaringish/intersect.pyx:32-190 and merge.pyx:54-134).  Only this corner takes it.
struct U64Buf { u64 *p; u64 n; };

#define DEV_GALLOP(ptr, end, cond)                    \
    do {                                              \
        u64 stride_ = 1;                              \
        while ((ptr) < (end) && (cond)) {             \
            (ptr) += stride_;                         \
            stride_ <<= 1;                            \
        }                                             \
        (ptr) -= (stride_ >> 1);                      \
    } while (0)

__device__ u64 dev_intersect_drop(const u64 *lhs, u64 nl, const u64 *rhs, u64 nr, u64 mask, u64 *li, u64 *ri) {
    const u64 *l = lhs, *r = rhs, *le = lhs + nl, *re = rhs + nr;
    u64 m = 0, last = ~0ull;
    while (l < le && r < re) {
        DEV_GALLOP(l, le, (*l & mask) < (*r & mask));
        DEV_GALLOP(r, re, (*r & mask) < (*l & mask));
        const u64 x = *l & mask, y = *r & mask;
        if (x < y) l++;
        else if (y < x) r++;
        else {
            if ((last & mask) != x) { li[m] = (u64)(l - lhs); if (ri) ri[m] = (u64)(r - rhs); last = *l; m++; }
            l++; r++;
        }
    }
    return m;
}

__device__ u64 dev_adjacent(const u64 *lhs, u64 nl, const u64 *rhs, u64 nr, u64 mask, u64 *li, u64 *ri) {
    const u64 delta = mask & (~mask + 1);
    const u64 *l = lhs, *r = rhs, *le = lhs + nl, *re = rhs + nr;
    u64 m = 0, last = ~0ull;
    while (r < re && (*r & mask) == 0) r++;
    while (l < le && r < re) {
        DEV_GALLOP(l, le, (*l & mask) < ((*r & mask) - delta));
        DEV_GALLOP(r, re, ((*r & mask) - delta) < (*l & mask));
        const u64 x = *l & mask, y = (*r & mask) - delta;
        if (x < y) l++;
        else if (y < x) r++;
        else {
            if ((last & mask) != x) { li[m] = (u64)(l - lhs); ri[m] = (u64)(r - rhs); last = *l; m++; }
            l++; r++;
        }
    }
    return m;
}

__device__ u64 dev_merge(const u64 *lhs, u64 nl, const u64 *rhs, u64 nr, bool drop, u64 *out) {
    u64 i = 0, j = 0, m = 0;
    while (i < nl && j < nr) {
        if (lhs[i] < rhs[j]) out[m++] = lhs[i++];
        else if (rhs[j] < lhs[i]) out[m++] = rhs[j++];
        else { out[m++] = lhs[i]; if (!drop) out[m++] = rhs[j]; i++; j++; }
    }
    while (j < nr) out[m++] = rhs[j++];
    while (i < nl) out[m++] = lhs[i++];
    return m;
}

// all rhs elements whose value occurs in lhs, found the way _gallop_intersect_keep walks
__device__ u64 dev_intersect_keep_rhs(const u64 *lhs, u64 nl, const u64 *rhs, u64 nr, u64 hdr_mask_rhs, u64 *r_out) {
    // lhs: header values (possibly unsorted here!), rhs: words compared by (word & hdr_mask_rhs)
    const u64 *l = lhs, *r = rhs, *le = lhs + nl, *re = rhs + nr;
    u64 m = 0;
    while (l < le && r < re) {
        DEV_GALLOP(l, le, *l < (*r & hdr_mask_rhs));
        DEV_GALLOP(r, re, (*r & hdr_mask_rhs) < *l);
        const u64 x = *l, y = *r & hdr_mask_rhs;
        if (x < y) l++;
        else if (y < x) r++;
        else {
            while (l < le && *l == x) l++;
            while (r < re && (*r & hdr_mask_rhs) == x) { r_out[m++] = *r; r++; }
        }
    }
    return m;
}

__global__ void span_candidates_literal_kernel(const SpanArgs a, u64 *scratch, u64 cap3 /* 3*|A| + 8 */) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const SpanQuery &sq = a.queries[0];
    const u32 n = sq.n_terms;
    const u64 M = SA_HDR_MASK;
    u64 *lh = scratch, *rh = lh + cap3, *tmp = rh + cap3, *tmp2 = tmp + cap3;
    u64 *last_l = tmp2 + cap3, *last_r = last_l + cap3;
    u64 *i0 = last_r + cap3, *i1 = i0 + cap3, *allh = i1 + cap3, *allh2 = allh + 4 * cap3;
    u64 n_ll = 0, n_lr = 0;
    const u64 *curr = a.words + sq.off[0];
    const u64 nc = sq.len[0];
    for (u32 k = 1; k < n; k++) {
        const u64 *nxt = a.words + sq.off[k];
        const u64 nn = sq.len[k];
        u64 m = dev_intersect_drop(curr, nc, nxt, nn, M, i0, nullptr);
        for (u64 j = 0; j < m; j++) tmp[j] = curr[i0[j]] & M;                     // int_headers
        u64 ma = dev_adjacent(curr, nc, nxt, nn, M, i0, i1);                      // curr_to_right, next_to_left
        for (u64 j = 0; j < ma; j++) tmp2[j] = nxt[i1[j]];
        u64 n_lh = dev_merge(tmp, m, tmp2, ma, false, lh);
        for (u64 j = 0; j < ma; j++) tmp2[j] = curr[i0[j]];
        u64 n_rh = dev_merge(tmp, m, tmp2, ma, false, rh);
        u64 mb = dev_adjacent(nxt, nn, curr, nc, M, i0, i1);                      // next_to_right, curr_to_left
        for (u64 j = 0; j < mb; j++) tmp2[j] = curr[i1[j]];
        n_lh = dev_merge(lh, n_lh, tmp2, mb, false, tmp);
        for (u64 j = 0; j < n_lh; j++) lh[j] = tmp[j];
        for (u64 j = 0; j < mb; j++) tmp2[j] = nxt[i0[j]];
        n_rh = dev_merge(rh, n_rh, tmp2, mb, false, tmp);
        for (u64 j = 0; j < n_rh; j++) rh[j] = tmp[j];
        if (k > 1) {
            u64 ml = dev_intersect_drop(last_l, n_ll, lh, n_lh, M, i0, nullptr);
            for (u64 j = 0; j < ml; j++) last_l[j] = last_l[i0[j]];              // ascending: in place is safe
            n_ll = ml;
            u64 mr = dev_intersect_drop(last_r, n_lr, rh, n_rh, M, i0, nullptr);
            for (u64 j = 0; j < mr; j++) last_r[j] = last_r[i0[j]];
            n_lr = mr;
        } else {
            for (u64 j = 0; j < n_lh; j++) last_l[j] = lh[j];
            for (u64 j = 0; j < n_rh; j++) last_r[j] = rh[j];
            n_ll = n_lh;
            n_lr = n_rh;
        }
    }
    for (u64 j = 0; j < n_lr; j++) tmp[j] = last_r[j] + SA_ONE_BLOCK;             // to_rhs
    for (u64 j = 0; j < n_ll; j++) tmp2[j] = last_l[j] - SA_ONE_BLOCK;            // to_lhs (may underflow)
    u64 na = dev_merge(tmp, n_lr, tmp2, n_ll, true, allh);
    na = dev_merge(last_l, n_ll, allh, na, true, allh2);
    na = dev_merge(last_r, n_lr, allh2, na, true, allh);
    for (u64 j = 0; j < na; j++) allh[j] &= M;
    for (u32 t = 0; t < n; t++) {
        u64 m = dev_intersect_keep_rhs(allh, na, a.words + sq.off[t], sq.len[t], M, a.word_arena + sq.s_off[t]);
        a.counts[0].n_sliced[t] = (u32)m;
    }
}

// ---------------------------------------------------------------------------- phase 2
struct WarpSpans {                        // one per warp, in dynamic shared memory
    u64 posns[SPAN_CAP];
    u32 terms[SPAN_CAP];
    int beg[SPAN_CAP];
    int end[SPAN_CAP];
    int cbeg[SPAN_CAP];                   // collected spans (_collect_spans)
    int cend[SPAN_CAP];
};

__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }

// _compact_spans (spans.pyx:141-154): keep width <= max_w and >= 1 term, order preserved
__device__ u32 compact_spans(WarpSpans &S, u32 cursor, int max_w) {
    const unsigned lane = threadIdx.x & 31;
    u32 w = 0;
    for (u32 base = 0; base < cursor; base += 32) {
        const u32 s = base + lane;
        bool keep = false;
        u64 po = 0; u32 te = 0; int be = 0, en = 0;
        if (s < cursor) {
            po = S.posns[s]; te = S.terms[s]; be = S.beg[s]; en = S.end[s];
            keep = (iabs(en - be) <= max_w) && (__popc(te) > 0);
        }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) {
            const u32 d = w + __popc(m & ((1u << lane) - 1));
            S.posns[d] = po; S.terms[d] = te; S.beg[d] = be; S.end[d] = en;
        }
        w += __popc(m);
        __syncwarp();
    }
    return w;
}

// _collect_spans (spans.pyx:157-186): complete, narrow-enough spans after first-come overlap
// replacement; returns how many were collected
__device__ u32 collect_spans(WarpSpans &S, u32 cursor, u32 n_terms, int max_w) {
    const unsigned lane = threadIdx.x & 31;
    u32 ncoll = 0;
    for (u32 s = 0; s < cursor; s++) {
        const u32 te = S.terms[s];
        const u64 po = S.posns[s];
        const int be = S.beg[s], en = S.end[s];
        const bool complete = ((u32)__popc(te) == n_terms) || ((u32)__popcll(po) == n_terms);
        const int nw = iabs(en - be);
        if (!(complete && nw < max_w)) continue;            // warp-uniform
        bool replaced = false;
        for (u32 base = 0; base < ncoll && !replaced; base += 32) {
            const u32 c = base + lane;
            bool hit = false;
            if (c < ncoll) {
                const int cb = S.cbeg[c], ce = S.cend[c];
                hit = (be <= ce && en >= cb) && (nw < iabs(ce - cb));
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                const u32 first = base + (u32)(__ffs(m) - 1);
                if (lane == 0) { S.cbeg[first] = be; S.cend[first] = en; }
                replaced = true;
            }
            __syncwarp();
        }
        if (!replaced) {
            if (lane == 0 && ncoll < SPAN_CAP) { S.cbeg[ncoll] = be; S.cend[ncoll] = en; }
            ncoll++;
            __syncwarp();
        }
    }
    return ncoll;
}

__global__ void __launch_bounds__(SPAN_WARPS * 32)
span_groups_kernel(const SpanArgs a, u32 n_queries) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSpans &S = reinterpret_cast<WarpSpans *>(smem_raw)[warp];
    const u32 warps_total = gridDim.x * SPAN_WARPS;
    const u32 warp_global = blockIdx.x * SPAN_WARPS + warp;

    for (u32 q = 0; q < n_queries; q++) {
        const SpanQuery &sq = a.queries[q];
        const SpanCounts &sc = a.counts[q];
        const u32 n = sq.n_terms;
        const int max_w = (int)(n + sq.slop);
        const u32 iters = sc.n_groups[0];                    // while curr_idx[0] < lengths[1]
        for (u32 it = warp_global; it < iters; it += warps_total) {
            u32 cursor = 0;
            bool full = false;
            u32 last_key = 0;
            u32 mn_pop = 0;                                    // running "min" with the reference's rule
            bool undefined = false;
            for (u32 t = 0; t < n; t++) {
                u32 sum_pop = 0;
                if (it < sc.n_groups[t]) {
                    const u64 *sl = a.word_arena + sq.s_off[t];
                    const u32 *gs = a.group_arena + sq.g_off[t];
                    const u32 w0 = gs[it], w1 = gs[it + 1];
                    const bool has_next_group = (it + 1 < sc.n_groups[t]);
                    const u32 term_bit = 1u << t;
                    bool give_up = false;
                    for (u32 wi = w0; wi < w1 && !give_up; wi++) {
                        const u64 word = sl[wi];
                        last_key = (u32)(word >> SA_KEY_SHIFT);
                        const int payload_base = (int)((word >> SA_LSB_BITS) & 0x3FFFFu) * SA_LSB_BITS;
                        u32 bits = (u32)(word & SA_LSB_MASK);
                        sum_pop += __popc(bits);
                        while (bits) {
                            const int set_idx = __ffs(bits) - 1;
                            bits &= bits - 1;
                            const int posn = set_idx + payload_base;
                            // spans.pyx:107-108 compiled as a 32-bit shift, sign-extended (see oracle)
                            const u64 posn_bit = (u64)(i64)(int)(1u << ((posn % 64) & 31));
                            const u32 fresh = cursor;
                            if (fresh >= SPAN_CAP) {           // reference: out-of-bounds write (undefined)
                                full = true;
                                undefined = true;
                                break;
                            }
                            if (lane == 0) { S.terms[fresh] = term_bit; S.posns[fresh] = posn_bit; S.beg[fresh] = posn; S.end[fresh] = posn; }
                            cursor++;
                            __syncwarp();
                            bool any_fail = false, any_ok = false;
                            for (u32 base = 0; base < fresh; base += 32) {
                                const u32 s = base + lane;
                                bool want_fork = false;
                                u32 te = 0; u64 po = 0; int be = 0, en = 0;
                                if (s < fresh) {
                                    te = S.terms[s]; po = S.posns[s];
                                    const u32 nt_before = __popc(te), np_before = __popcll(po);
                                    if (!(nt_before < n && np_before == n) && !(te & term_bit)) {
                                        const u64 po_new = po | posn_bit;
                                        S.posns[s] = po_new;                       // kept even when cancelled
                                        be = S.beg[s]; en = S.end[s];
                                        const bool cancel = ((u32)__popcll(po_new) == np_before) || (iabs(posn - be) > max_w);
                                        if (!cancel) {
                                            want_fork = true;
                                            te |= term_bit;
                                            S.terms[s] = te;
                                            po = po_new;
                                        }
                                    }
                                }
                                unsigned m = __ballot_sync(0xffffffffu, want_fork);
                                if (want_fork) {
                                    const u32 slot = cursor + __popc(m & ((1u << lane) - 1));
                                    if (slot < SPAN_CAP) {
                                        S.terms[slot] = te; S.posns[slot] = po & ~posn_bit; S.beg[slot] = be; S.end[slot] = en;
                                    }
                                    S.end[s] = posn;
                                }
                                const u32 nf = __popc(m);
                                if (nf) {
                                    if (cursor + nf > SPAN_CAP) any_fail = true;
                                    if (cursor < SPAN_CAP) any_ok = true;
                                    cursor = min(cursor + nf, (u32)SPAN_CAP);
                                }
                                __syncwarp();
                            }
                            if (any_fail) full = true; else if (any_ok) full = false;
                            if (cursor >= SPAN_CAP) break;
                        }
                        if (cursor >= SPAN_CAP) {
                            cursor = compact_spans(S, cursor, max_w);
                            if (cursor >= SPAN_CAP && has_next_group) give_up = true;   // skip to the next doc group
                        }
                    }
                }
                if (mn_pop == 0 || sum_pop < mn_pop) mn_pop = sum_pop;
            }
            u32 add;
            if (full) add = mn_pop;
            else add = collect_spans(S, cursor, n, max_w);
            if (lane == 0) {
                const u64 d = (u64)last_key - a.doc_base;
                if (d < a.n_docs) atomicAdd(a.out + (u64)q * a.out_stride + d, (float)add);
                if (undefined) atomicAdd(&a.counts[q].undefined, 1u);
            }
            __syncwarp();
        }
    }
}

// --------------------------------------------------------------------------------- host
static u64 padded_stride(u64 n_docs) { return (n_docs + SA_TILE_DOCS - 1) / SA_TILE_DOCS * SA_TILE_DOCS; }

// Span search of one query into ix->dense row 0 (raw counts).  Caller holds ix->mu.
int sa_span_run(sa_index *ix, const u64 *d_lists, const u64 *offs, const u64 *lens, uint32_t n_terms, uint32_t slop,
                u32 *n_undefined) {
    const u64 stride = padded_stride(ix->n_docs);
    int rc;
    if ((rc = ix->dense.reserve(stride * sizeof(float)))) return rc;
    SA_CUDA(cudaMemsetAsync(ix->dense.p, 0, stride * sizeof(float), ix->stream));
    SpanQuery sq;
    memset(&sq, 0, sizeof(sq));
    sq.n_terms = n_terms;
    sq.slop = slop;
    u64 shortest_len = ~0ull;
    for (u32 t = 0; t < n_terms; t++) {
        sq.off[t] = offs[t];
        sq.len[t] = lens[t];
        if (sq.len[t] < shortest_len) { shortest_len = sq.len[t]; sq.shortest = t; }
    }
    u64 words_total = 0, groups_total = 0;
    for (u32 t = 0; t < n_terms; t++) {
        sq.s_cap[t] = std::min<u64>(sq.len[t], 5 * shortest_len);
        sq.s_off[t] = words_total;
        sq.g_off[t] = groups_total;
        words_total += sq.s_cap[t] + 2;
        groups_total += sq.s_cap[t] + 2;
    }
    SA_CHECK(groups_total < 0xFFFFFFFFull, "slop query too large");
    if ((rc = ix->phrase_scratch.reserve(words_total * sizeof(u64) + groups_total * sizeof(u32) + 256))) return rc;
    if ((rc = ix->queries.reserve(sizeof(SpanQuery)))) return rc;
    if ((rc = ix->cand_meta.reserve(sizeof(SpanCounts)))) return rc;
    SA_CUDA(cudaMemcpyAsync(ix->queries.p, &sq, sizeof(sq), cudaMemcpyHostToDevice, ix->stream));
    SA_CUDA(cudaMemsetAsync(ix->cand_meta.p, 0, sizeof(SpanCounts), ix->stream));
    SpanArgs a;
    a.words = d_lists;
    a.queries = ix->queries.as<SpanQuery>();
    a.counts = ix->cand_meta.as<SpanCounts>();
    a.word_arena = ix->phrase_scratch.as<u64>();
    a.group_arena = (u32 *)(a.word_arena + words_total);
    a.out = ix->dense.as<float>();
    a.out_stride = stride;
    a.n_docs = ix->n_docs;
    a.doc_base = ix->doc_base;
    {
        KernelTimer t(ix, 2);
        // the reference's header-0 underflow corner (see span_candidates_literal_kernel)
        bool literal = true;
        for (u32 t = 0; t < n_terms; t++) {
            u64 first = 0;
            if (sq.len[t] == 0) { literal = false; break; }
            SA_CUDA(cudaMemcpyAsync(&first, d_lists + sq.off[t], sizeof(u64), cudaMemcpyDeviceToHost, ix->stream));
            SA_CUDA(cudaStreamSynchronize(ix->stream));
            if ((first & SA_HDR_MASK) != 0) { literal = false; break; }
        }
        if (literal) {
            const u64 cap3 = 3 * sq.len[0] + 3 * shortest_len + 16;
            DevBuf lit;
            if ((rc = lit.reserve((9 * cap3 + 8 * cap3) * sizeof(u64)))) return rc;
            span_candidates_literal_kernel<<<1, 1, 0, ix->stream>>>(a, lit.as<u64>(), cap3);
            SA_CUDA(cudaGetLastError());
            SA_CUDA(cudaStreamSynchronize(ix->stream));
            lit.release();
        } else {
            span_candidates_kernel<<<1, CAND_THREADS, 0, ix->stream>>>(a);
            SA_CUDA(cudaGetLastError());
        }
        span_groups_build_kernel<<<1, CAND_THREADS, 0, ix->stream>>>(a);
        SA_CUDA(cudaGetLastError());
        SA_CUDA(cudaFuncSetAttribute(span_groups_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(SPAN_WARPS * sizeof(WarpSpans))));
        const u32 blocks = (u32)std::max<u64>(1, std::min<u64>((u64)ix->num_sms * 2, (5 * shortest_len + SPAN_WARPS - 1) / SPAN_WARPS));
        span_groups_kernel<<<blocks, SPAN_WARPS * 32, SPAN_WARPS * sizeof(WarpSpans), ix->stream>>>(a, 1);
        SA_CUDA(cudaGetLastError());
        t.stop();
        ix->stats.phrase_kernel_launches += 3;
        ix->stats.total_launches += 3;
    }
    SpanCounts h;
    SA_CUDA(cudaMemcpyAsync(&h, ix->cand_meta.p, sizeof(h), cudaMemcpyDeviceToHost, ix->stream));
    SA_CUDA(cudaStreamSynchronize(ix->stream));
    SA_CHECK(!h.overflow, "span candidate arena exhausted (internal sizing error)");
    if (n_undefined) *n_undefined = h.undefined;
    return SA_OK;
}

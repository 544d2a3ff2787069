// sa_phrase.cu -- exact phrase matching (slop == 0) on roaringish posting words.
//
// Replaces (reference paths relative to softwaredoug/searcharray):
//   compute_phrase_freqs + L->R / R->L drivers + _intersect_bigram_matches  phrase/middle_out.py:73-168
//   bigram_freqs, _inner_bigram_freqs, _inner_bigram_same_term, _adj_to_phrase_freq,
//   _adjacent_bigram_freqs, _set_adjbit_at_header                          phrase/bigram_freqs.py:48-307
//   intersect_with_adjacents / intersect / merge / sort_merge_counts /
//   popcount_reduce_at / key_sum_over                                       roaringish/*.pyx
//   PosnBitArray.phrase_freqs dense scatter                                 phrase/middle_out.py:418-446
//
// Design.  Phrase matching never crosses a document, and every term's words are sorted by doc
// id, so the doc-id space is cut into chunks and ONE CTA RUNS THE WHOLE n-TERM PIPELINE FOR ONE
// (query, doc-range chunk): it locates each term's slice for its doc range (warp-cooperative
// 32-ary search), then for every bigram step walks the shorter "driver" list one element per
// thread, binary-searches the other list for the equal header and the adjacent header (the
// reference's galloping intersect has plain set semantics on header-unique lists, SURVEY 8a row
// 10), does the 18-bit shift/AND/popcount, emits the continuation words in order through a block
// scan, and reduces the per-doc counts with a segmented sum.  The running min over steps is a
// search into the previous step's (doc, count) list.  Matches are scattered into the pre-zeroed
// dense vector, optionally through BM25.
//
// The reference's "same term" branch (bigram_freqs.py:139) depends on a GLOBAL property
// (all equal-header pairs identical).  Each launch runs with a speculated flag per step and
// counts (pairs, differing pairs) per step; the host verifies and re-launches on a mis-guess
// (only adversarial inputs ever do).
#include <algorithm>

#include "sa_phrase.cuh"
#include "sa_span.cuh"
#include "sa_term.cuh"
#include "sa_tma.cuh"

int sa_filter_terms(sa_index *ix, const uint32_t *term_ids, uint32_t n_terms, bool use_rows,
                    u64 pay_lo, u64 pay_hi, bool use_payload, std::vector<u64> &offs, std::vector<u64> &lens);
int sa_gather_rows(sa_index *ix, const float *d_dense, float *out_host);

#define PT SA_PHRASE_THREADS
#define PW_SUB_DOCS (SA_TILE_DOCS / (SA_PHRASE_THREADS / 32))   // docs of a tile that one warp of the merge regime owns (1,024)

static u64 docs_per_chunk_of(const sa_index *ix, u32 n_chunks);

struct Elem {
    u64 w0, w1;
    u32 n_emit, cnt, doc;
    bool entry, inner, diff;
};

__device__ __forceinline__ u64 lower_bound_hdr(const u64 *__restrict__ a, u64 n, u64 target) {
    u64 lo = 0, hi = n;
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        if ((a[mid] & SA_HDR_MASK) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One equal-header pair (bigram_freqs.py:104-155 normal branch, :65-101 same-term branch).
__device__ __forceinline__ void inner_calc(u64 l, u64 r, bool same, bool cont_rhs, u32 &cnt, u64 &word) {
    if (!same) {
        u64 ov = (l & SA_LSB_MASK) & ((r & SA_LSB_MASK) >> 1);
        cnt = (u32)__popcll(ov);
        word = cont_rhs ? (((ov << 1) & SA_LSB_MASK) | (r & SA_HDR_MASK)) : (ov | (l & SA_HDR_MASK));
    } else {
        u64 full = l & (r << 1);
        u32 adj = (u32)__popcll(full & SA_LSB_MASK);
        u32 runs = (u32)__popcll((full & (full << 1)) & SA_LSB_MASK);
        cnt = adj - ((runs + 1) >> 1);                        // _adj_to_phrase_freq: - ceil(runs/2)
        word = cont_rhs ? ((((r << 1) & r) & SA_LSB_MASK) | (l & ~SA_LSB_MASK))
                        : ((l & ~SA_LSB_MASK) | ((l & (l >> 1)) & SA_LSB_MASK));   // `>> 1` leak kept
    }
}

// What driver element i contributes.  D = driver list, O = the other list.
template <bool CONT_RHS, bool DRIVER_LHS>
__device__ __forceinline__ Elem compute_elem(const u64 *__restrict__ D, u64 nD, u64 i,
                                             const u64 *__restrict__ O, u64 nO, bool same) {
    Elem e;
    e.w0 = e.w1 = 0;
    e.n_emit = 0;
    e.cnt = 0;
    e.entry = e.inner = e.diff = false;
    const u64 x = D[i];
    const u64 h = x & SA_HDR_MASK;
    e.doc = (u32)(x >> SA_KEY_SHIFT);
    if (DRIVER_LHS) {
        // x is an lhs word: partners are rhs words at header h (inner) and h + 1 block (adjacent)
        u64 pos = lower_bound_hdr(O, nO, h);
        bool inner = pos < nO && (O[pos] & SA_HDR_MASK) == h;
        u64 r = inner ? O[pos] : 0;
        u64 pos2 = pos + (inner ? 1 : 0);
        bool has_adj = pos2 < nO && (O[pos2] & SA_HDR_MASK) == h + SA_ONE_BLOCK;
        u64 r2 = has_adj ? O[pos2] : 0;
        bool am = has_adj && (x & SA_BIT17) && (r2 & 1ull);
        if (inner) {
            u32 c;
            u64 word;
            inner_calc(x, r, same, CONT_RHS, c, word);
            e.cnt += c;
            if (CONT_RHS) {
                // the rhs word at h may also end a cross-word match that started in lhs[i-1]
                if (i > 0) {
                    u64 lp = D[i - 1];
                    if ((lp & SA_HDR_MASK) + SA_ONE_BLOCK == h && (lp & SA_BIT17) && (r & 1ull)) word |= 1ull;
                }
            } else if (am) {
                word |= SA_BIT17;
            }
            e.w0 = word;
            e.n_emit = 1;
            e.inner = true;
            e.diff = (x != r);
        }
        if (am) {
            e.cnt += 1;
            if (CONT_RHS) {
                // adjacent-only continuation, unless lhs[i+1] pairs with that rhs word itself
                bool next_handles = (i + 1 < nD) && ((D[i + 1] & SA_HDR_MASK) == h + SA_ONE_BLOCK);
                if (!next_handles) {
                    u64 w = (r2 & SA_HDR_MASK) | 1ull;
                    if (e.n_emit == 0) e.w0 = w; else e.w1 = w;
                    e.n_emit++;
                }
            } else if (!inner) {
                e.w0 = h | SA_BIT17;
                e.n_emit = 1;
            }
        }
        e.entry = inner || am;
    } else {
        // x is an rhs word: partners are lhs words at header h - 1 block (adjacent) and h (inner)
        const bool can_adj = h >= SA_ONE_BLOCK;
        u64 pos = lower_bound_hdr(O, nO, can_adj ? h - SA_ONE_BLOCK : h);
        bool has_adj = can_adj && pos < nO && (O[pos] & SA_HDR_MASK) == h - SA_ONE_BLOCK;
        u64 lp = has_adj ? O[pos] : 0;
        u64 posl = pos + (has_adj ? 1 : 0);
        bool inner = posl < nO && (O[posl] & SA_HDR_MASK) == h;
        u64 l = inner ? O[posl] : 0;
        bool am = has_adj && (lp & SA_BIT17) && (x & 1ull);
        if (!CONT_RHS && am) {
            // adjacent-only continuation on the lhs side, unless rhs[i-1] pairs with lp itself
            bool prev_handles = (i > 0) && ((D[i - 1] & SA_HDR_MASK) == h - SA_ONE_BLOCK);
            if (!prev_handles) {
                e.w0 = (lp & SA_HDR_MASK) | SA_BIT17;
                e.n_emit = 1;
            }
        }
        if (inner) {
            u32 c;
            u64 word;
            inner_calc(l, x, same, CONT_RHS, c, word);
            e.cnt += c;
            if (CONT_RHS) {
                if (am) word |= 1ull;
            } else if (i + 1 < nD) {
                u64 rn = D[i + 1];
                if ((rn & SA_HDR_MASK) == h + SA_ONE_BLOCK && (l & SA_BIT17) && (rn & 1ull)) word |= SA_BIT17;
            }
            if (e.n_emit == 0) e.w0 = word; else e.w1 = word;
            e.n_emit++;
            e.inner = true;
            e.diff = (l != x);
            e.doc = (u32)(l >> SA_KEY_SHIFT);
        } else if (am) {
            if (CONT_RHS) {
                e.w0 = h | 1ull;
                e.n_emit = 1;
            }
            e.doc = (u32)(lp >> SA_KEY_SHIFT);
        }
        if (am) e.cnt += 1;
        e.entry = inner || am;
    }
    return e;
}

#include "sa_phrase_warp.cuh"

struct StepShared {
    u32 warp_sums[PT / 32];
    u32 edoc[PT];
    u32 ecnt[PT];
    u32 carry_doc, carry_cnt, carry_valid;
    u32 st_inner, st_diff;
    u64 n_cont, n_docs;
};

// exclusive block scan (all PT threads call); returns the exclusive prefix, `total` = block sum
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *warp_sums, u32 &total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                 // protect warp_sums from the previous use
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    u32 base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < PT / 32; w++) {
        u32 s = warp_sums[w];
        if (w < (int)warp) base += s;
        tot += s;
    }
    total = tot;
    return base + incl - v;
}

// One bigram step over this CTA's chunk.  Writes the continuation list (sorted) to cont_out and
// the per-doc counts (doc << 32 | count, sorted by doc, zero counts kept) to docs_out.
template <bool CONT_RHS, bool DRIVER_LHS>
__device__ void bigram_step(const u64 *__restrict__ D, u64 nD, const u64 *__restrict__ O, u64 nO, bool same,
                            u64 *__restrict__ cont_out, u64 *__restrict__ docs_out, StepShared &S) {
    const unsigned tid = threadIdx.x;
    if (tid == 0) {
        S.n_cont = 0;
        S.n_docs = 0;
        S.carry_valid = 0;
        S.st_inner = 0;
        S.st_diff = 0;
    }
    __syncthreads();
    for (u64 t0 = 0; t0 < nD; t0 += PT) {
        const u64 i = t0 + tid;
        Elem e;
        e.n_emit = 0;
        e.entry = e.inner = e.diff = false;
        e.cnt = 0;
        e.doc = 0;
        e.w0 = e.w1 = 0;
        if (i < nD) e = compute_elem<CONT_RHS, DRIVER_LHS>(D, nD, i, O, nO, same);
        // speculation bookkeeping
        unsigned mi = __ballot_sync(0xffffffffu, e.inner), md = __ballot_sync(0xffffffffu, e.diff);
        if ((tid & 31) == 0 && mi) {
            atomicAdd(&S.st_inner, (u32)__popc(mi));
            if (md) atomicAdd(&S.st_diff, (u32)__popc(md));
        }
        // continuation words, in order
        u32 total;
        u32 off = block_excl_scan(e.n_emit, S.warp_sums, total);
        const u64 cbase = S.n_cont;
        if (e.n_emit >= 1) cont_out[cbase + off] = e.w0;
        if (e.n_emit == 2) cont_out[cbase + off + 1] = e.w1;
        // (doc, count) entries of this tile, compacted into shared memory
        u32 etotal;
        u32 eoff = block_excl_scan(e.entry ? 1u : 0u, S.warp_sums, etotal);
        if (e.entry) {
            S.edoc[eoff] = e.doc;
            S.ecnt[eoff] = e.cnt;
        }
        __syncthreads();
        if (tid == 0) S.n_cont = cbase + total;
        if (etotal) {     // block-uniform
            const bool flush = S.carry_valid && S.edoc[0] != S.carry_doc;
            const bool merge = S.carry_valid && S.edoc[0] == S.carry_doc;
            const u32 c_doc = S.carry_doc, c_cnt = S.carry_cnt;
            const u64 dbase = S.n_docs;
            // segmented sum: the head of each run adds up its run
            bool emit = false, is_last = false;
            u32 sum = 0, doc = 0;
            if (tid < etotal) {
                doc = S.edoc[tid];
                bool head = (tid == 0) || (S.edoc[tid - 1] != doc);
                if (head) {
                    u32 k = tid;
                    while (k < etotal && S.edoc[k] == doc) sum += S.ecnt[k++];
                    if (tid == 0 && merge) sum += c_cnt;
                    is_last = (k == etotal);
                    emit = !is_last;
                }
            }
            u32 htotal;
            u32 hoff = block_excl_scan(emit ? 1u : 0u, S.warp_sums, htotal);
            const u64 obase = dbase + (flush ? 1 : 0);
            if (emit) docs_out[obase + hoff] = ((u64)doc << 32) | sum;
            if (tid == 0 && flush) docs_out[dbase] = ((u64)c_doc << 32) | c_cnt;
            __syncthreads();
            if (is_last) {            // exactly one thread: the head of the tile's last run
                S.carry_doc = doc;
                S.carry_cnt = sum;
                S.carry_valid = 1;
                S.n_docs = obase + htotal;
            }
        }
        __syncthreads();
    }
    if (tid == 0 && S.carry_valid) {
        docs_out[S.n_docs] = ((u64)S.carry_doc << 32) | S.carry_cnt;
        S.n_docs += 1;
        S.carry_valid = 0;
    }
    __syncthreads();
}

// cur[i].count = min(cur[i].count, prev[doc].count) (0 if the doc is not in prev)
// == _intersect_bigram_matches (middle_out.py:73-93) on nested / sorted doc lists.
__device__ void and_min(u64 *__restrict__ cur, u64 n_cur, const u64 *__restrict__ prev, u64 n_prev) {
    for (u64 i = threadIdx.x; i < n_cur; i += PT) {
        u64 e = cur[i];
        u64 doc = e >> 32;
        u64 lo = 0, hi = n_prev;
        while (lo < hi) {
            u64 mid = (lo + hi) >> 1;
            if ((prev[mid] >> 32) < doc) lo = mid + 1; else hi = mid;
        }
        u32 c = 0;
        if (lo < n_prev && (prev[lo] >> 32) == doc) c = min((u32)(prev[lo] & 0xFFFFFFFFull), (u32)(e & 0xFFFFFFFFull));
        cur[i] = (doc << 32) | c;
    }
    __syncthreads();
}

struct ChainResult { u64 *docs; u64 n_docs; u64 *cont; u64 n_cont; };

// ---------------------------------------------------------------------------------------------------------------
// The SEARCH regime (|shortest list| << |the others|): one CTA per (query, doc-range chunk) runs the whole chain once
// over its chunk; a step's driver elements binary-search the other list in global memory, which skips most of it.
__global__ void __launch_bounds__(PT, 4)
phrase_kernel(const PhraseArgs a) {
    __shared__ StepShared S;
    __shared__ u64 s_lo[SA_MAX_PHRASE_TERMS], s_n[SA_MAX_PHRASE_TERMS];
    __shared__ u64 s_slab;
    __shared__ int s_ok;
    __shared__ __align__(16) float s_tile[SA_TILE_DOCS];
    __shared__ u32 s_top[(PT / 32) * 8];
    __shared__ u32 s_ncand, s_tile_max;

    // grid = (queries, chunks): neighbouring CTAs belong to different queries (see term_tile_kernel)
    const u32 q = a.qsel ? a.qsel[blockIdx.x] : blockIdx.x;
    const u32 chunk = blockIdx.y;
    const PhraseQuery &pq = a.queries[q];
    const u32 n_terms = pq.n_terms;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 d0 = a.doc_base + (u64)chunk * a.docs_per_chunk;
    const u64 dend = a.doc_base + a.n_docs;
    if (d0 >= dend) return;                    // (grid is sized so this does not happen)
    const u64 d1 = min(d0 + a.docs_per_chunk, dend);
    ChainResult fin;
    fin.docs = nullptr; fin.n_docs = 0; fin.cont = nullptr; fin.n_cont = 0;
    bool run = true;

    // 1. every term's slice for this doc range
    for (u32 t = warp; t < n_terms; t += PT / 32) {
        const u64 *lst = a.words + pq.off[t];
        u64 lo, hi;
        if (pq.dir_plus1[t] && a.tile_dir) {               // chunks are whole tiles: the list's tile directory has the slice
            const u32 *dir = a.tile_dir + (pq.dir_plus1[t] - 1);
            lo = __ldg(dir + (u32)(((u64)chunk * a.docs_per_chunk) / SA_TILE_DOCS));
            hi = __ldg(dir + (u32)((d1 - a.doc_base + SA_TILE_DOCS - 1) / SA_TILE_DOCS));
        } else {
            lo = warp_lower_bound_shifted(lst, 0, pq.len[t], d0, SA_KEY_SHIFT);
            hi = warp_lower_bound_shifted(lst, lo, pq.len[t], d1, SA_KEY_SHIFT);
        }
        if (lane == 0) { s_lo[t] = lo; s_n[t] = hi - lo; }
    }
    __syncthreads();
    u64 cap = 0, widest = 0;
    for (u32 t = 0; t < n_terms; t++) {
        cap = max(cap, s_n[t]);
        if (s_n[t]) widest++;
    }
    // A chunk where fewer than two terms occur has no pairs at any step.  (A chunk that merely
    // misses ONE term must still run its earlier steps: their pairs count towards the global
    // same-term decision of the reference.)
    if (widest < 2) run = false;
    cap += 2;

    // 2. scratch slab: 2 continuation buffers + 4 (doc,count) buffers
    if (tid == 0) {
        s_ok = 1;
        s_slab = 0;
        if (run) {
            unsigned long long need = 6ull * cap;
            unsigned long long at = atomicAdd(a.arena_used, need);
            s_ok = (at + need <= a.arena_cap);
            s_slab = at;
            if (!s_ok) atomicExch(&a.stats[q].overflow, 1u);
        }
    }
    __syncthreads();
    if (!s_ok) run = false;
    u64 *contA = a.arena + s_slab, *contB = contA + cap;
    u64 *docsA = contB + cap, *docsB = docsA + cap, *docsL = docsB + cap, *docsR = docsL + cap;

    auto slice = [&](u32 t) { return a.words + pq.off[t] + s_lo[t]; };

    // Runs one chain over terms [ta, tb).  lr: left-to-right (cont = RHS) else right-to-left.
    auto run_chain = [&](u32 ta, u32 tb, bool lr, u64 *final_docs) -> ChainResult {
        ChainResult res;
        res.docs = final_docs;
        res.n_docs = 0;
        res.cont = contA;
        res.n_cont = 0;
        const u64 *carry = lr ? slice(ta) : slice(tb - 1);
        u64 n_carry = lr ? s_n[ta] : s_n[tb - 1];
        u64 *cont_bufs[2] = {contA, contB};
        u64 *doc_bufs[2] = {docsA, docsB};
        int flip = 0;
        const u64 *prev_docs = nullptr;
        u64 n_prev = 0;
        const u32 n_steps = tb - ta - 1;
        for (u32 s = 0; s < n_steps; s++) {
            const u32 tnew = lr ? (ta + 1 + s) : (tb - 2 - s);     // also the step id
            const bool same = (pq.same_guess >> tnew) & 1u;
            const u64 *other = slice(tnew);
            const u64 n_other = s_n[tnew];
            if (n_carry == 0 || n_other == 0) {   // no pairs from here on in this doc range
                res.n_docs = 0;
                res.n_cont = 0;
                break;
            }
            u64 *cont_out = cont_bufs[flip];
            u64 *docs_out = (s == n_steps - 1) ? final_docs : doc_bufs[flip];
            // the first step may drive from the shorter side; later steps drive from the carry
            const bool drive_carry = (s > 0) || (n_carry <= n_other);
            if (lr) {
                if (drive_carry) bigram_step<true, true>(carry, n_carry, other, n_other, same, cont_out, docs_out, S);
                else bigram_step<true, false>(other, n_other, carry, n_carry, same, cont_out, docs_out, S);
            } else {
                if (drive_carry) bigram_step<false, false>(carry, n_carry, other, n_other, same, cont_out, docs_out, S);
                else bigram_step<false, true>(other, n_other, carry, n_carry, same, cont_out, docs_out, S);
            }
            const u64 n_cont = S.n_cont, n_docs = S.n_docs;
            if (tid == 0) {
                if (S.st_inner) atomicAdd(&a.stats[q].n_inner[tnew], S.st_inner);
                if (S.st_diff) atomicAdd(&a.stats[q].n_diff[tnew], S.st_diff);
                if (n_cont) atomicAdd(&a.stats[q].n_cont, (unsigned long long)n_cont);
            }
            __syncthreads();
            if (prev_docs) and_min(docs_out, n_docs, prev_docs, n_prev);
            prev_docs = docs_out;
            n_prev = n_docs;
            carry = cont_out;
            n_carry = n_cont;
            res.docs = docs_out;
            res.n_docs = n_docs;
            res.cont = cont_out;
            res.n_cont = n_cont;
            flip ^= 1;
            if (n_docs == 0) {       // nothing can survive the remaining steps
                res.n_docs = 0;
                break;
            }
        }
        return res;
    };

    if (run) {
    if (pq.mode == SA_PHRASE_MODE_LR) {
        fin = run_chain(0, n_terms, true, docsL);
    } else if (pq.mode == SA_PHRASE_MODE_RL) {
        fin = run_chain(0, n_terms, false, docsL);
    } else {
        // both chains always run (their pair statistics feed the speculation check)
        ChainResult left = run_chain(0, pq.split, true, docsL);
        fin = run_chain(pq.split, n_terms, false, docsR);
        if (left.n_docs == 0) fin.n_docs = 0;
        and_min(fin.docs, fin.n_docs, left.docs, left.n_docs);
    }

    }
    // optional dump for the per-op parity export (single chunk)
    if (a.dump.cont) {
        for (u64 i = tid; i < fin.n_cont; i += PT) a.dump.cont[i] = fin.cont[i];
        for (u64 i = tid; i < fin.n_docs; i += PT) a.dump.docs[i] = fin.docs[i];
        if (tid == 0) { *a.dump.n_cont = fin.n_cont; *a.dump.n_docs = fin.n_docs; }
    }

    // 3. materialise the dense vector of this doc range tile by tile (phrase_freqs[ids] = counts,
    //    middle_out.py:441): zeros + the matches that fall in the tile, flushed with 16-byte
    //    streaming stores; the same pass collects the tile's top-k candidates.
    float *out = a.out + (u64)q * a.out_stride;
    Bm25Params p = a.bm25;
    p.idf = pq.idf;
    const u32 row = a.topk_row0 + q;
    const u32 tile0 = (u32)(((u64)chunk * a.docs_per_chunk) / SA_TILE_DOCS);
    const u32 tile1 = (u32)((d1 - a.doc_base + SA_TILE_DOCS - 1) / SA_TILE_DOCS);
    // fin.docs is sorted by doc and the tiles ascend: a running cursor replaces a search per tile.
    // Most tiles hold no match at all: they are written as zeros straight from registers (no shared
    // tile, no barrier), so the bulk of the 4*N write runs at fill speed.
    u64 cur = 0;
    u64 next_doc = fin.n_docs ? (fin.docs[0] >> 32) : ~0ull;          // CTA-uniform
    for (u32 tile = tile0; tile < tile1; tile++) {
        const u64 t_abs1 = a.doc_base + (u64)tile * SA_TILE_DOCS + SA_TILE_DOCS;
        if (next_doc >= t_abs1) {
            float4 *__restrict__ out4 = reinterpret_cast<float4 *>(out + (u64)tile * SA_TILE_DOCS);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < SA_TILE_DOCS / PT / 4; i++) __stcs(out4 + tid + i * PT, z);
            if (a.topk.k && tid == 0) {
                const u64 t_idx = (u64)row * a.topk.n_tiles + tile;
                a.topk.tile_cnt[t_idx] = 0;
                a.topk.tile_max[t_idx] = 0;
            }
            continue;
        }
        // first entry at or past the end of this tile: gallop from the cursor, then bisect (uniform)
        const u64 m0 = cur;
        u64 lo = cur + 1, hi = fin.n_docs, st = 1;
        while (lo < hi) {
            const u64 probe = min(lo + st - 1, hi - 1);
            if ((fin.docs[probe] >> 32) < t_abs1) { lo = probe + 1; st <<= 1; }
            else { hi = probe; break; }
        }
        while (lo < hi) {
            const u64 mid = (lo + hi) >> 1;
            if ((fin.docs[mid] >> 32) < t_abs1) lo = mid + 1; else hi = mid;
        }
        const u64 m1 = lo;
        cur = m1;
        next_doc = m1 < fin.n_docs ? (fin.docs[m1] >> 32) : ~0ull;
#pragma unroll
        for (int i = 0; i < SA_TILE_DOCS / PT / 4; i++)
            reinterpret_cast<float4 *>(s_tile)[tid + i * PT] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        u32 my_max = 0, my_match = 0;
        for (u64 i = m0 + tid; i < m1; i += PT) {
            const u64 e = fin.docs[i];
            const u32 c = (u32)(e & 0xFFFFFFFFull);
            if (c == 0) continue;
            const u64 d = (e >> 32) - a.doc_base;
            if (d >= a.n_docs) continue;
            my_match++;
            const float v = a.score ? bm25_one((float)c, __ldg(a.doc_lens + d), p) : (float)c;
            s_tile[d - (u64)tile * SA_TILE_DOCS] = v;
            if (v > 0.0f) my_max = max(my_max, __float_as_uint(v));
        }
        my_match = __reduce_add_sync(0xffffffffu, my_match);
        if (lane == 0 && my_match) atomicAdd(&a.stats[q].n_match, my_match);
        __syncthreads();
        flush_tile_collect(s_tile, out + (u64)tile * SA_TILE_DOCS, a.topk, row, tile, my_max, (u32)(m1 - m0),
                           (u32)min(m1 - m0, (u64)PT), s_top, &s_ncand, &s_tile_max);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// The MERGE regime (balanced lists; chosen per query by the host, sa_phrase_is_staged): persistent CTAs, each claiming
// (query, 16-tile chunk) work items.  A work item is cut into SEGMENTS of whole tiles whose posting slices fit in one
// half of the CTA's staging buffer.  Warp 0 plans a segment from the chunk's tile-directory entries (preloaded into
// shared memory) and its lane 0 arms an mbarrier and issues one TMA bulk copy per term
// (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes); the NEXT segment is planned and issued into
// the other half before the current one is consumed, so the copy engine streams every list from HBM exactly once,
// sequentially, while the SM works (double buffering).  Per tile:
//   1. a doc can only match if it holds EVERY term, so each term's slice sets bits in a doc-presence bitmap
//      (shared-memory atomicOr) and the bitmaps are ANDed.  Conjunctions of common terms are rare (df/N of .3, .1,
//      .03, .01: 9e-6 of the docs), so almost every tile ends here and is written as zeros;
//   2. otherwise each of the eight warps takes 1,024 docs of the tile: it compacts the candidate docs' words of
//      every term in order (ballots) and runs the whole bigram chain on them at warp scope (sa_phrase_warp.cuh);
//   3. the tile is materialised (zeros + BM25-scored matches) and its top-k candidates collected.
// The pair statistics of the same-term speculation cover the candidate docs only -- enough to CONFIRM a guess, not
// to derive the reference's global decision from; any disagreement re-runs the query in the search regime
// (sa_phrase_run_sync / redo_query), which keeps the result exact.
#define PS_MAX_CHUNK_TILES 32

struct SegPlan {                                   // one per staging-buffer half
    u32 te, any;
    u32 n[SA_MAX_PHRASE_TERMS];                    // words of each term in the segment
    u32 dir0[SA_MAX_PHRASE_TERMS];                 // directory entry of the segment's first tile (tile slices by subtraction)
    const u64 *ptr[SA_MAX_PHRASE_TERMS];           // where the segment's slice is read from (staged copy, or global)
};

struct StagedShared {
    u32 dirs[SA_MAX_PHRASE_TERMS][PS_MAX_CHUNK_TILES + 1];   // tile-directory entries of the chunk's tile boundaries
    u32 has_dir[SA_MAX_PHRASE_TERMS];
    u64 c_lo[SA_MAX_PHRASE_TERMS], c_n[SA_MAX_PHRASE_TERMS];  // chunk slices (lists without a directory are searched)
    SegPlan plan[2];
    const u64 *sptr[PT / 32][SA_MAX_PHRASE_TERMS]; // every warp's chain inputs (compacted candidates, or plain sub-slices)
    u32 sn[PT / 32][SA_MAX_PHRASE_TERMS];
    u32 wmatch[2][PT / 32];                        // matches of every warp, alternating between consecutive tiles
    u32 ncand, tile_max, work, ok;
    u32 top[(PT / 32) * 8];
    __align__(16) float tile[SA_TILE_DOCS];        // bitmaps / compaction buffers first, then the dense tile
    __align__(8) u64 bar[2];
};

// warp 0 only: plan the segment that starts at tile `ts` and issue its copies into staging half `h`
__device__ __forceinline__ void staged_plan(const PhraseArgs &a, const PhraseQuery &pq, StagedShared &P, u64 *stage_half,
                                            u32 h, u32 ts, u32 tile0, u32 tile1, u64 dend) {
    const unsigned lane = threadIdx.x & 31;
    const u32 n_terms = pq.n_terms;
    SegPlan &pl = P.plan[h];
    const bool has = lane < n_terms;
    const bool hd = has && P.has_dir[lane];
    const u32 base = hd ? P.dirs[lane][ts - tile0] : 0u;
    const u32 fixed = (has && !hd) ? (u32)min(P.c_n[lane], (u64)0x7FFFFFFFu) : 0u;
    u32 best = ts + 1;
    for (u32 cand = ts + 1; cand <= tile1; cand++) {
        const u32 mine = has ? ((hd ? P.dirs[lane][cand - tile0] - base : fixed) + 4u) : 0u;     // + alignment slack
        const u32 sum = __reduce_add_sync(0xffffffffu, mine);
        if (cand > ts + 1 && sum > a.stage_words) break;
        best = cand;
        if (sum > a.stage_words) break;
    }
    const u32 te = best;
    // every term's slice of the segment
    u64 lo = 0;
    u32 n = 0;
    if (has) {
        if (hd) {
            lo = base;
            n = P.dirs[lane][te - tile0] - base;
        } else {                                   // short list without a directory: search its chunk slice
            const u64 *lst = a.words + pq.off[lane] + P.c_lo[lane];
            const u64 seg_d0 = a.doc_base + (u64)ts * SA_TILE_DOCS, seg_d1 = min(a.doc_base + (u64)te * SA_TILE_DOCS, dend);
            u32 l2 = 0, h2 = (u32)P.c_n[lane];
            while (l2 < h2) { const u32 m = (l2 + h2) >> 1; if ((lst[m] >> SA_KEY_SHIFT) < seg_d0) l2 = m + 1; else h2 = m; }
            u32 l3 = l2, h3 = (u32)P.c_n[lane];
            while (l3 < h3) { const u32 m = (l3 + h3) >> 1; if ((lst[m] >> SA_KEY_SHIFT) < seg_d1) l3 = m + 1; else h3 = m; }
            lo = P.c_lo[lane] + l2;
            n = l3 - l2;
        }
    }
    // staging layout: terms in order while they fit (exclusive scan over the lanes)
    const u64 *lst = has ? a.words + pq.off[lane] : nullptr;
    const u32 wds = (has && n) ? sa_stage_bytes(lst, lo, n) / 8u : 0u;
    u32 incl = wds;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
    const bool st = wds && incl <= a.stage_words;                 // (a term that does not fit is read from global memory)
    const u32 so = incl - wds;
    const u32 total = __reduce_add_sync(0xffffffffu, st ? wds * 8u : 0u);
    if (lane == 0 && total) {
        sa_fence_proxy_async();                                  // the half's previous readers are behind a barrier
        sa_mbar_expect_tx(&P.bar[h], total);
    }
    __syncwarp();
    if (has) {
        pl.n[lane] = n;
        pl.dir0[lane] = base;
        if (st) {
            const u32 head = sa_stage_issue(stage_half + so, lst, lo, n, &P.bar[h]);
            pl.ptr[lane] = stage_half + so + head;
        } else {
            pl.ptr[lane] = lst + lo;
        }
    }
    if (lane == 0) { pl.te = te; pl.any = total ? 1u : 0u; }
    __syncwarp();
}

__device__ void phrase_work_staged(const PhraseArgs &a, const u32 q, const u32 chunk, StagedShared &P, u64 *stage,
                                   u32 (&bar_phase)[2], u64 *cta_slab, const u64 cap) {
    const PhraseQuery &pq = a.queries[q];
    const u32 n_terms = pq.n_terms;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 d0 = a.doc_base + (u64)chunk * a.docs_per_chunk;
    const u64 dend = a.doc_base + a.n_docs;
    if (d0 >= dend) return;
    const u64 d1 = min(d0 + a.docs_per_chunk, dend);
    const u32 tile0 = (u32)(((u64)chunk * a.docs_per_chunk) / SA_TILE_DOCS);
    const u32 tile1 = (u32)((d1 - a.doc_base + SA_TILE_DOCS - 1) / SA_TILE_DOCS);
    const u32 nt = tile1 - tile0;                                 // <= PS_MAX_CHUNK_TILES (host: sa_phrase_staged_chunks)
    float *out = a.out + (u64)q * a.out_stride;
    Bm25Params p = a.bm25;
    p.idf = pq.idf;
    const u32 row = a.topk_row0 + q;

    // ---- the chunk's directory entries -> shared memory; chunk slices of the lists without a directory
    __syncthreads();                                              // the previous work item is done with P
    for (u32 idx = tid; idx < n_terms * (nt + 1); idx += PT) {
        const u32 t = idx / (nt + 1), j = idx % (nt + 1);
        const bool hd = pq.dir_plus1[t] && a.tile_dir;
        P.dirs[t][j] = hd ? __ldg(a.tile_dir + (pq.dir_plus1[t] - 1) + tile0 + j) : 0u;
        if (j == 0) P.has_dir[t] = hd ? 1u : 0u;
    }
    for (u32 t = warp; t < n_terms; t += PT / 32) {
        if (pq.dir_plus1[t] && a.tile_dir) continue;
        const u64 *lst = a.words + pq.off[t];
        const u64 lo = warp_lower_bound_shifted(lst, 0, pq.len[t], d0, SA_KEY_SHIFT);
        const u64 hi = warp_lower_bound_shifted(lst, lo, pq.len[t], d1, SA_KEY_SHIFT);
        if (lane == 0) { P.c_lo[t] = lo; P.c_n[t] = hi - lo; }
    }
    if (tid == 0) P.ok = 1;
    __syncthreads();
    u32 widest = 0;
    for (u32 t = 0; t < n_terms; t++) {
        const u64 n = P.has_dir[t] ? (u64)(P.dirs[t][nt] - P.dirs[t][0]) : P.c_n[t];
        if (n) widest++;
        if (n + 2 > cap && !P.has_dir[t]) { if (tid == 0) { P.ok = 0; atomicExch(&a.stats[q].overflow, 1u); } }
    }
    const bool run = widest >= 2;                                 // fewer than two terms present: no pairs at any step

    if (run && warp == 0) staged_plan(a, pq, P, stage, 0, tile0, tile0, tile1, dend);
    __syncthreads();
    const bool ok = P.ok != 0;

    u32 ts = tile0, seg = 0, tile_no = 0;
    while (ts < tile1) {
        const u32 h = seg & 1u;
        u32 te = tile1;
        if (run) {
            te = P.plan[h].te;
            // prefetch: plan the next segment and issue its copies into the other half before consuming this one
            if (te < tile1 && warp == 0) staged_plan(a, pq, P, stage + (u64)(h ^ 1u) * a.stage_words, h ^ 1u, te, tile0, tile1, dend);
            if (P.plan[h].any) {
                sa_mbar_wait(&P.bar[h], bar_phase[h]);
                bar_phase[h] ^= 1u;
            }
        }
        const SegPlan &pl = P.plan[h];
        for (u32 tile = ts; tile < te; tile++, tile_no++) {
            // ---- every warp takes 1,024 docs of the tile and works on them WITHOUT block barriers: sub-slice bounds
            //      (lane-parallel binary searches), doc-presence bitmaps of the terms (one 32-bit word per lane), their
            //      AND, ordered compaction of the candidate docs' words, the bigram chain, and its 4 KB of the dense tile.
            //      The warp's scratch overlays its own slice of the tile.
            const u64 td0 = a.doc_base + (u64)tile * SA_TILE_DOCS;
            const u64 w_d0 = td0 + (u64)warp * PW_SUB_DOCS, w_d1 = w_d0 + PW_SUB_DOCS;
            float *my_slice = P.tile + warp * PW_SUB_DOCS;
            u32 *wbm = reinterpret_cast<u32 *>(my_slice);                       // [n_terms][32] presence bitmaps
            u32 *wcand = wbm + n_terms * 32;                                    // [32] candidate docs
            u64 *fbw = reinterpret_cast<u64 *>(wcand + 32);                     // compaction area
            const u32 fb_cap = ((PW_SUB_DOCS * 4 - (n_terms + 1) * 128) / 8) / n_terms;
            WarpFin wf;
            wf.docs = nullptr;
            wf.n_docs = 0;
            if (run && ok) {
                // the tile's slice of term `lane`
                const u64 *t_ptr = nullptr;
                u32 t_n = 0;
                if (lane < n_terms) {
                    const u64 *base = pl.ptr[lane];
                    const u32 n = pl.n[lane];
                    u32 lo = 0, hi = n;
                    if (te - ts > 1) {
                        if (P.has_dir[lane]) {
                            lo = P.dirs[lane][tile - tile0] - pl.dir0[lane];
                            hi = P.dirs[lane][tile + 1 - tile0] - pl.dir0[lane];
                        } else {
                            lo = w_lower_bound_doc(base, n, td0);
                            hi = lo + w_lower_bound_doc(base + lo, n - lo, td0 + SA_TILE_DOCS);
                        }
                    }
                    t_ptr = base + lo;
                    t_n = hi - lo;
                }
                // this warp's sub-slice bounds: lane 2t searches the lower, lane 2t+1 the upper doc bound of term t
                u32 bound = 0;
                {
                    const u32 t = lane >> 1;
                    const u64 ptr_bits = __shfl_sync(0xffffffffu, (u64)(uintptr_t)t_ptr, t);
                    const u32 n = __shfl_sync(0xffffffffu, t_n, t);
                    if (t < n_terms) bound = w_lower_bound_doc(reinterpret_cast<const u64 *>((uintptr_t)ptr_bits), n, (lane & 1u) ? w_d1 : w_d0);
                }
                bool all_present = true;
                for (u32 t = 0; t < n_terms; t++) wbm[t * 32 + lane] = 0u;
                __syncwarp();
                for (u32 t = 0; t < n_terms; t++) {
                    const u64 *base = reinterpret_cast<const u64 *>((uintptr_t)__shfl_sync(0xffffffffu, (u64)(uintptr_t)t_ptr, t));
                    const u32 lo = __shfl_sync(0xffffffffu, bound, 2 * t), hi = __shfl_sync(0xffffffffu, bound, 2 * t + 1);
                    all_present = all_present && hi > lo;
                    if (!all_present) break;                                    // warp-uniform
                    u32 *bm = wbm + t * 32;
                    for (u32 i = lo + lane; i < hi; i += 32) {
                        const u32 rel = (u32)((base[i] >> SA_KEY_SHIFT) - w_d0);
                        atomicOr(&bm[rel >> 5], 1u << (rel & 31u));
                    }
                }
                __syncwarp();
                u32 c = 0;
                if (all_present) {
                    c = wbm[lane];
                    for (u32 t = 1; t < n_terms; t++) c &= wbm[t * 32 + lane];
                }
                if (__reduce_add_sync(0xffffffffu, (u32)__popc(c))) {          // warp-uniform: candidates in this sub-range
                    wcand[lane] = c;
                    __syncwarp();
                    for (u32 t = 0; t < n_terms; t++) {
                        const u64 *base = reinterpret_cast<const u64 *>((uintptr_t)__shfl_sync(0xffffffffu, (u64)(uintptr_t)t_ptr, t));
                        const u32 lo = __shfl_sync(0xffffffffu, bound, 2 * t), hi = __shfl_sync(0xffffffffu, bound, 2 * t + 1);
                        u64 *dst = fbw + (u64)t * fb_cap;
                        u32 kept = 0;
                        for (u32 i0 = lo; i0 < hi; i0 += 32) {                 // ordered compaction by ballots
                            const u32 i = i0 + lane;
                            u64 w = 0;
                            bool keep = false;
                            if (i < hi) {
                                w = base[i];
                                const u32 rel = (u32)((w >> SA_KEY_SHIFT) - w_d0);
                                keep = (wcand[rel >> 5] >> (rel & 31u)) & 1u;
                            }
                            const unsigned m = __ballot_sync(0xffffffffu, keep);
                            const u32 at = kept + __popc(m & ((1u << lane) - 1u));
                            if (keep && at < fb_cap) dst[at] = w;
                            kept += __popc(m);
                        }
                        if (lane == 0) {
                            // (a list too long to compact enters the chain whole: its extra docs die at the other terms)
                            P.sptr[warp][t] = kept <= fb_cap ? dst : base + lo;
                            P.sn[warp][t] = kept <= fb_cap ? kept : hi - lo;
                        }
                    }
                    __syncwarp();
                    wf = warp_phrase_chain(pq, P.sptr[warp], P.sn[warp], cta_slab + (u64)warp * 6ull * cap, cap, &a.stats[q]);
                }
            }
            // ---- this warp's 4 KB of the dense tile: zeros + its matches
            __syncwarp();
#pragma unroll
            for (int i = 0; i < PW_SUB_DOCS / 32 / 4; i++)
                reinterpret_cast<float4 *>(my_slice)[lane + i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            u32 my_max = 0, my_match = 0;
            for (u32 i = lane; i < wf.n_docs; i += 32) {
                const u64 e = wf.docs[i];
                const u32 c = (u32)(e & 0xFFFFFFFFull);
                if (c == 0) continue;
                const u64 d = (e >> 32) - a.doc_base;
                if (d >= a.n_docs) continue;
                my_match++;
                const float v = a.score ? bm25_one((float)c, __ldg(a.doc_lens + d), p) : (float)c;
                P.tile[d - (u64)tile * SA_TILE_DOCS] = v;
                if (v > 0.0f) my_max = max(my_max, __float_as_uint(v));
            }
            my_match = __reduce_add_sync(0xffffffffu, my_match);
            if (lane == 0) {
                if (my_match) atomicAdd(&a.stats[q].n_match, my_match);
                P.wmatch[tile_no & 1u][warp] = my_match;
            }
            __syncthreads();                                                  // the tile's only block barrier (besides the flush's own)
            u32 total = 0, holders = 0;
#pragma unroll
            for (int w = 0; w < PT / 32; w++) { const u32 m = P.wmatch[tile_no & 1u][w]; total += m; holders += min(m, 32u); }
            if (total == 0) {
                float4 *__restrict__ out4 = reinterpret_cast<float4 *>(out + (u64)tile * SA_TILE_DOCS);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < SA_TILE_DOCS / PT / 4; i++) __stcs(out4 + tid + i * PT, z);
                if (a.topk.k && tid == 0) {
                    const u64 t_idx = (u64)row * a.topk.n_tiles + tile;
                    a.topk.tile_cnt[t_idx] = 0;
                    a.topk.tile_max[t_idx] = 0;
                }
                continue;                                                      // (the next tile's scratch lives in the warps' own slices)
            }
            flush_tile_collect(P.tile, out + (u64)tile * SA_TILE_DOCS, a.topk, row, tile, my_max, total, holders,
                               P.top, &P.ncand, &P.tile_max);
        }
        __syncthreads();                    // every read of this half of the staging buffer is done: it may be refilled
        ts = te;
        seg++;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The CONJUNCTION regime (balanced lists): one CTA per (query, 8192-doc tile), like the term scan -- small shared
// memory footprint, five CTAs per SM, latency hidden by occupancy.  A doc can only match if it holds EVERY term of the
// phrase, so the CTA streams the tile's slice of every term ONCE with coalesced loads and sets bits in per-term
// doc-presence bitmaps (shared-memory atomicOr), ANDs them, and in the (rare) tiles where candidate docs exist each
// warp takes its 1,024 docs: lane-parallel binary searches for its sub-slice bounds, ordered compaction of the
// candidates' words by ballots, the whole bigram chain at warp scope (sa_phrase_warp.cuh) -- all inside the warp's own
// 4 KB of the tile.  HBM traffic: 8 * sum(W) + 4 * N, each list read once, sequentially: the B_phrase of SURVEY 8d
// without its continuation term.  A sub-range whose candidates do not fit its compaction area flags the query for
// the exact re-run in the search regime (the host routes phrases whose terms co-occur that often there up front).
__global__ void __launch_bounds__(PT, 5)
phrase_tile_kernel(const PhraseArgs a) {
    __shared__ __align__(16) float s_tile[SA_TILE_DOCS];
    __shared__ u32 s_top[(PT / 32) * 8];
    __shared__ u32 s_ncand, s_tile_max;
    __shared__ u32 s_wmatch[PT / 32];
    __shared__ const u64 *s_ptr[PT / 32][SA_MAX_PHRASE_TERMS];
    __shared__ u32 s_n[PT / 32][SA_MAX_PHRASE_TERMS];

    const u32 q = a.qsel ? a.qsel[blockIdx.x] : blockIdx.x;
    const u32 tile = blockIdx.y;
    const PhraseQuery &pq = a.queries[q];
    const u32 n_terms = pq.n_terms;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 td0 = a.doc_base + (u64)tile * SA_TILE_DOCS;
    const u64 td1 = min(td0 + SA_TILE_DOCS, a.doc_base + a.n_docs);
    float *out = a.out + (u64)q * a.out_stride;
    const u32 row = a.topk_row0 + q;

    // the tile's slice of term `lane`: its tile directory, or (short lists) a search
    u64 t_lo = 0;
    u32 t_n = 0;
    for (u32 t = 0; t < n_terms; t++) {
        if (pq.dir_plus1[t] && a.tile_dir) continue;
        const u64 *lst = a.words + pq.off[t];                                  // (all lanes search, lane t keeps the result)
        const u64 lo = warp_lower_bound_shifted(lst, 0, pq.len[t], td0, SA_KEY_SHIFT);
        const u64 hi = warp_lower_bound_shifted(lst, lo, pq.len[t], td1, SA_KEY_SHIFT);
        if (lane == t) { t_lo = lo; t_n = (u32)(hi - lo); }
    }
    if (lane < n_terms && pq.dir_plus1[lane] && a.tile_dir) {
        const u32 *dir = a.tile_dir + (pq.dir_plus1[lane] - 1) + tile;
        t_lo = __ldg(dir);
        t_n = __ldg(dir + 1) - (u32)t_lo;
    }
    const bool all_present = __all_sync(0xffffffffu, lane >= n_terms || t_n > 0);

    u32 c = 0;                                                                  // candidate docs: 32 per thread
    if (all_present) {                                                          // CTA-uniform
        u32 *term_bm = reinterpret_cast<u32 *>(s_tile);                         // [n_terms][256]
        for (u32 i = tid; i < n_terms * (SA_TILE_DOCS / 32); i += PT) term_bm[i] = 0u;
        __syncthreads();
        for (u32 t = 0; t < n_terms; t++) {
            const u64 *lst = a.words + pq.off[t] + __shfl_sync(0xffffffffu, t_lo, t);
            const u32 n = __shfl_sync(0xffffffffu, t_n, t);
            u32 *bm = term_bm + t * (SA_TILE_DOCS / 32);
            u32 i = tid;
            for (; i + 3 * PT < n; i += 4 * PT) {                               // four independent loads in flight per thread
                const u64 w0 = ld_stream_u64(lst + i), w1 = ld_stream_u64(lst + i + PT);
                const u64 w2 = ld_stream_u64(lst + i + 2 * PT), w3 = ld_stream_u64(lst + i + 3 * PT);
                const u32 r0 = (u32)((w0 >> SA_KEY_SHIFT) - td0), r1 = (u32)((w1 >> SA_KEY_SHIFT) - td0);
                const u32 r2 = (u32)((w2 >> SA_KEY_SHIFT) - td0), r3 = (u32)((w3 >> SA_KEY_SHIFT) - td0);
                atomicOr(&bm[r0 >> 5], 1u << (r0 & 31u));
                atomicOr(&bm[r1 >> 5], 1u << (r1 & 31u));
                atomicOr(&bm[r2 >> 5], 1u << (r2 & 31u));
                atomicOr(&bm[r3 >> 5], 1u << (r3 & 31u));
            }
            for (; i < n; i += PT) {
                const u32 rel = (u32)((ld_stream_u64(lst + i) >> SA_KEY_SHIFT) - td0);
                atomicOr(&bm[rel >> 5], 1u << (rel & 31u));
            }
        }
        __syncthreads();
        c = term_bm[tid];
        for (u32 t = 1; t < n_terms; t++) c &= term_bm[t * (SA_TILE_DOCS / 32) + tid];
        __syncthreads();                                                        // the bitmaps are dead: the warps' slices are free
    }

    // ---- candidate docs (rare): every warp on its own 1,024 docs, inside its own 4 KB of the tile
    float *my_slice = s_tile + warp * PW_SUB_DOCS;
    u64 res[2] = {0ull, 0ull};                                                  // this lane's (doc << 32 | count) results
    u32 n_res = 0;
    if (__reduce_add_sync(0xffffffffu, (u32)__popc(c))) {                       // warp-uniform
        u32 *wcand = reinterpret_cast<u32 *>(my_slice);                         // [32]
        // capacities so that candidates + the chain's six buffers fit in the slice
        const u32 fb_cap = (PW_SUB_DOCS * 4 - 128 - 96) / (8 * n_terms + 48);
        u64 *fb = reinterpret_cast<u64 *>(wcand + 32);                          // [n_terms][fb_cap]
        u64 *chain_buf = fb + (u64)n_terms * fb_cap;                            // 6 x (fb_cap + 2)
        const u64 w_d0 = td0 + (u64)warp * PW_SUB_DOCS, w_d1 = w_d0 + PW_SUB_DOCS;
        wcand[lane] = c;
        // sub-slice bounds: lane 2t searches the lower, lane 2t+1 the upper doc bound of term t (global memory, L2-hot)
        u32 bound = 0;
        {
            const u32 t = lane >> 1;
            const u64 lo_t = __shfl_sync(0xffffffffu, t_lo, t);
            const u32 n = __shfl_sync(0xffffffffu, t_n, t);
            if (t < n_terms) bound = w_lower_bound_doc(a.words + pq.off[t] + lo_t, n, (lane & 1u) ? w_d1 : w_d0);
        }
        __syncwarp();
        bool fits = true;
        for (u32 t = 0; t < n_terms; t++) {
            const u64 *base = a.words + pq.off[t] + __shfl_sync(0xffffffffu, t_lo, t);
            const u32 lo = __shfl_sync(0xffffffffu, bound, 2 * t), hi = __shfl_sync(0xffffffffu, bound, 2 * t + 1);
            u64 *dst = fb + (u64)t * fb_cap;
            u32 kept = 0;
            for (u32 i0 = lo; i0 < hi; i0 += 32) {                             // ordered compaction by ballots
                const u32 i = i0 + lane;
                u64 w = 0;
                bool keep = false;
                if (i < hi) {
                    w = base[i];
                    const u32 rel = (u32)((w >> SA_KEY_SHIFT) - w_d0);
                    keep = (wcand[rel >> 5] >> (rel & 31u)) & 1u;
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                const u32 at = kept + __popc(m & ((1u << lane) - 1u));
                if (keep && at < fb_cap) dst[at] = w;
                kept += __popc(m);
            }
            fits = fits && kept <= fb_cap;
            if (lane == 0) { s_ptr[warp][t] = dst; s_n[warp][t] = min(kept, fb_cap); }
        }
        __syncwarp();
        if (fits) {
            const WarpFin wf = warp_phrase_chain(pq, s_ptr[warp], s_n[warp], chain_buf, fb_cap + 2, &a.stats[q]);
            n_res = wf.n_docs;                                                   // <= fb_cap + 2 <= 64: two per lane
            if (lane < n_res) res[0] = wf.docs[lane];
            if (lane + 32 < n_res) res[1] = wf.docs[lane + 32];
        } else if (lane == 0) {
            atomicExch(&a.stats[q].overflow, 2u);                               // dense conjunction: re-run in the search regime
        }
        __syncwarp();
    }
    // ---- this warp's 4 KB of the dense tile: zeros + its matches
#pragma unroll
    for (int i = 0; i < PW_SUB_DOCS / 32 / 4; i++)
        reinterpret_cast<float4 *>(my_slice)[lane + i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    u32 my_max = 0, my_match = 0;
    Bm25Params p = a.bm25;
    p.idf = pq.idf;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        if (lane + 32 * j >= n_res) continue;
        const u32 cnt = (u32)(res[j] & 0xFFFFFFFFull);
        if (cnt == 0) continue;
        const u64 d = (res[j] >> 32) - a.doc_base;
        if (d >= a.n_docs) continue;
        my_match++;
        const float v = a.score ? bm25_one((float)cnt, __ldg(a.doc_lens + d), p) : (float)cnt;
        s_tile[d - (u64)tile * SA_TILE_DOCS] = v;
        if (v > 0.0f) my_max = max(my_max, __float_as_uint(v));
    }
    my_match = __reduce_add_sync(0xffffffffu, my_match);
    if (lane == 0) {
        if (my_match) atomicAdd(&a.stats[q].n_match, my_match);
        s_wmatch[warp] = my_match;
    }
    __syncthreads();
    u32 total = 0, holders = 0;
#pragma unroll
    for (int w = 0; w < PT / 32; w++) { total += s_wmatch[w]; holders += min(s_wmatch[w], 32u); }
    if (total == 0) {
        float4 *__restrict__ out4 = reinterpret_cast<float4 *>(out + (u64)tile * SA_TILE_DOCS);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < SA_TILE_DOCS / PT / 4; i++) __stcs(out4 + tid + i * PT, z);
        if (a.topk.k && tid == 0) {
            const u64 t_idx = (u64)row * a.topk.n_tiles + tile;
            a.topk.tile_cnt[t_idx] = 0;
            a.topk.tile_max[t_idx] = 0;
        }
        return;
    }
    flush_tile_collect(s_tile, out + (u64)tile * SA_TILE_DOCS, a.topk, row, tile, my_max, total, holders, s_top, &s_ncand, &s_tile_max);
}

// Persistent CTAs (grid = resident CTAs of the device): work items (query, chunk) are claimed with an atomic
// counter, each CTA keeps its staging buffer (dynamic shared memory, two halves), its mbarriers and its scratch slab.
__global__ void __launch_bounds__(PT, 2)
phrase_staged_kernel(const PhraseArgs a) {
    __shared__ StagedShared P;
    extern __shared__ __align__(16) u64 s_stage[];
    if (threadIdx.x == 0) {
        sa_mbar_init(&P.bar[0], 1);
        sa_mbar_init(&P.bar[1], 1);
        sa_mbar_fence_init();
    }
    __syncthreads();
    u32 phase[2] = {0u, 0u};
    u64 *slab = a.slabs + (u64)blockIdx.x * (PT / 32) * 6ull * a.slab_cap;      // six buffers for each of the CTA's warps
    const u32 n_work = a.n_sel * a.n_chunks;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) P.work = atomicAdd(a.work_counter, 1u);
        __syncthreads();
        const u32 w = P.work;
        if (w >= n_work) break;
        // consecutive work items belong to different queries (dense and sparse lists interleave on an SM)
        const u32 q = a.qsel[w % a.n_sel], chunk = w / a.n_sel;
        phrase_work_staged(a, q, chunk, P, s_stage, phase, slab, a.slab_cap);
    }
}

int launch_phrase(sa_index *ix, const PhraseArgs &a, u32 n_queries) {
    if (n_queries == 0 || a.n_docs == 0) return SA_OK;
    dim3 grid(n_queries, a.n_chunks);
    KernelTimer t(ix, 2);
    phrase_kernel<<<grid, PT, 0, ix->stream>>>(a);
    SA_CUDA(cudaGetLastError());
    t.stop();
    ix->stats.phrase_kernel_launches++;
    ix->stats.total_launches++;
    return SA_OK;
}

// resident CTAs of phrase_staged_kernel with `stage_words` words of dynamic shared memory
static int staged_grid(sa_index *ix, u32 stage_words, u32 *ctas_out) {
    static bool attr_set = false;
    const size_t dyn = 2 * (size_t)stage_words * sizeof(u64);            // two halves (double buffering)
    if (!attr_set) {
        SA_CUDA(cudaFuncSetAttribute(phrase_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    int per_sm = 0;
    SA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, phrase_staged_kernel, PT, dyn));
    SA_CHECK(per_sm >= 1, "phrase_staged_kernel does not fit on an SM with %u staged words", stage_words);
    *ctas_out = (u32)per_sm * (u32)ix->num_sms;
    return SA_OK;
}

u32 sa_phrase_stage_words() {
    static const long env = getenv("SA_PHRASE_STAGE_WORDS") ? atol(getenv("SA_PHRASE_STAGE_WORDS")) : 0;
    return env > 0 ? (u32)std::min<long>(env, 9 * 1024) : 4608u;        // per half; 2 x 36 KB + 38 KB static: two CTAs per SM
}

// Merge regime or search regime?  Staging reads every list once (8 * sum(W) bytes); the search path costs about
// `ratio` bytes (a dozen 32-byte sectors of dependent probes) per driver element of the first step.
bool sa_phrase_is_staged(const PhraseQuery &pq, u64 n_docs) {
    static const long env = getenv("SA_PHRASE_STAGE_RATIO") ? atol(getenv("SA_PHRASE_STAGE_RATIO")) : -1;
    const u64 ratio = env >= 0 ? (u64)env : 50;
    if (ratio == 0) return false;
    const u32 n = pq.n_terms;
    if (n < 2) return false;
    // expected candidate words per 1,024-doc sub-range (terms taken as independent): they must fit the warp's
    // compaction area with room to spare, or the conjunction regime would keep bouncing the query to the search regime
    double est = (double)PW_SUB_DOCS;
    for (u32 t = 0; t < n; t++) est *= std::min(1.0, (double)pq.len[t] / (double)std::max<u64>(n_docs, 1));
    const double fb_cap = (double)((PW_SUB_DOCS * 4 - 128 - 96) / (8 * n + 48));
    if (est * 1.5 * 3.0 > fb_cap) return false;
    u64 sum = 0, drive = ~0ull;
    for (u32 t = 0; t < n; t++) {
        sum += pq.len[t];
        if (pq.len[t] == 0) return false;
    }
    if (pq.mode == SA_PHRASE_MODE_LR) drive = std::min(pq.len[0], pq.len[1]);
    else if (pq.mode == SA_PHRASE_MODE_RL) drive = std::min(pq.len[n - 1], pq.len[n - 2]);
    else {
        drive = std::min(pq.len[n - 1], pq.len[n - 2]);
        if (pq.split >= 2) drive = std::max(drive, std::min(pq.len[0], pq.len[1]));
    }
    return drive * ratio > sum;
}

// ------------------------------------------------------------------------------- host side
// BM25 over every doc (bm25.pyx:20-25) for parameter sets where tf == 0 does not score +0.0.
__global__ void bm25_dense_kernel(float *__restrict__ tf, const float *__restrict__ dl, u64 n, Bm25Params p) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tf[i] = bm25_one(tf[i], dl[i], p);
}

static u64 padded(u64 n_docs) { return (n_docs + SA_TILE_DOCS - 1) / SA_TILE_DOCS * SA_TILE_DOCS; }

// direction / speculation plan exactly as compute_phrase_freqs picks it (middle_out.py:154-168)
void sa_phrase_plan(PhraseQuery &pq, const u32 *term_ids) {
    const u32 n = pq.n_terms;
    u32 shortest = 0;
    for (u32 i = 1; i < n; i++) if (pq.len[i] < pq.len[shortest]) shortest = i;   // first minimum
    pq.same_guess = 0;
    if (shortest <= 1) {
        pq.mode = SA_PHRASE_MODE_LR;
        if (term_ids[0] == term_ids[1]) pq.same_guess |= 1u << 1;
    } else if (shortest >= n - 2) {
        pq.mode = SA_PHRASE_MODE_RL;
        if (term_ids[n - 2] == term_ids[n - 1]) pq.same_guess |= 1u << (n - 2);
    } else {
        pq.mode = SA_PHRASE_MODE_MID;
        pq.split = shortest;
        if (term_ids[0] == term_ids[1]) pq.same_guess |= 1u << 1;
        if (term_ids[n - 2] == term_ids[n - 1]) pq.same_guess |= 1u << (n - 2);
    }
}

// order in which the steps of a plan run (for verifying the speculation)
static void step_order(const PhraseQuery &pq, std::vector<u32> &order) {
    order.clear();
    const u32 n = pq.n_terms;
    if (pq.mode == SA_PHRASE_MODE_LR) for (u32 s = 1; s < n; s++) order.push_back(s);
    else if (pq.mode == SA_PHRASE_MODE_RL) for (int s = (int)n - 2; s >= 0; s--) order.push_back((u32)s);
    else {
        for (u32 s = 1; s < pq.split; s++) order.push_back(s);
        for (int s = (int)n - 2; s >= (int)pq.split; s--) order.push_back((u32)s);
    }
}

// Runs phrase queries (already planned) into ix->dense; loops until the same-term speculation
// of every query is confirmed.  lists may live in ix->d_words (off = absolute word offsets).
int sa_phrase_run_sync(sa_index *ix, std::vector<PhraseQuery> &pqs, const u64 *d_words,
                              int score, const Bm25Params &p, u32 n_chunks_hint, PhraseDump dump, u64 staged_slab_cap) {
    const u32 Q = (u32)pqs.size();
    const u64 stride = padded(ix->n_docs);
    int rc;
    if ((rc = ix->dense.reserve((size_t)Q * stride * sizeof(float)))) return rc;
    if ((rc = ix->queries.reserve((size_t)Q * sizeof(PhraseQuery)))) return rc;
    if ((rc = ix->cand_meta.reserve((size_t)Q * sizeof(PhraseStats) + 64))) return rc;
    // scratch arena: per query <= 6 * (sum of list lengths + 2 per chunk)
    u32 n_chunks = n_chunks_hint;
    if (n_chunks == 0) {
        u64 want = std::max<u64>(1, (u64)ix->num_sms * 8 / std::max<u32>(Q, 1));
        n_chunks = (u32)std::min<u64>(want, std::max<u64>(1, ix->n_docs / 512));
        n_chunks = std::max<u32>(n_chunks, 1);
    }
    n_chunks = sa_phrase_chunks(ix, n_chunks);
    const u64 docs_per_chunk = docs_per_chunk_of(ix, n_chunks);
    u64 arena_words = 64;
    for (auto &pq : pqs) {
        u64 sum = 0;
        for (u32 t = 0; t < pq.n_terms; t++) sum += pq.len[t];
        arena_words += 6 * (sum + 2ull * n_chunks);
    }
    // merge regime (single query on the index's own lists): persistent CTAs + TMA staging, no bump arena
    bool staged = staged_slab_cap && Q == 1 && d_words == ix->d_words && !dump.cont && sa_phrase_is_staged(pqs[0], ix->n_docs);
    const u64 full_arena_words = arena_words;
    const std::vector<PhraseQuery> pqs_in = pqs;
    if (staged) arena_words = 64;
    if ((rc = ix->phrase_scratch.reserve(arena_words * sizeof(u64) + 64))) return rc;
    unsigned long long *d_used = (unsigned long long *)ix->phrase_scratch.p;
    u64 *d_arena = (u64 *)ix->phrase_scratch.p + 8;
    PhraseStats *d_stats = (PhraseStats *)ix->cand_meta.p;
    std::vector<PhraseStats> h_stats(Q);
    std::vector<u32> order;
    if (staged) {
        if ((rc = ix->misc.reserve(256))) return rc;
        SA_CUDA(cudaMemsetAsync(ix->misc.p, 0, sizeof(u32), ix->stream));      // qsel = {0}
    }

    for (int attempt = 0; attempt < (int)SA_MAX_PHRASE_TERMS + 2; attempt++) {
        SA_CUDA(cudaMemcpyAsync(ix->queries.p, pqs.data(), (size_t)Q * sizeof(PhraseQuery), cudaMemcpyHostToDevice, ix->stream));
        SA_CUDA(cudaMemsetAsync(d_stats, 0, (size_t)Q * sizeof(PhraseStats), ix->stream));
        SA_CUDA(cudaMemsetAsync(d_used, 0, 64, ix->stream));
        if (staged) {
            PhraseSplit sp;
            memset(&sp, 0, sizeof(sp));
            sp.d_staged = (const u32 *)ix->misc.p;
            sp.n_staged = 1;
            sp.staged_chunks = sa_phrase_staged_chunks(ix);
            sp.slab_cap = staged_slab_cap;
            if ((rc = sa_phrase_enqueue(ix, ix->queries.as<PhraseQuery>(), d_stats, 1, ix->dense.as<float>(), stride, 1, d_arena,
                                        d_used, arena_words, score, p, nullptr, 0, &sp))) return rc;
        } else {
        PhraseArgs a;      // (the kernel writes every tile of the dense rows itself: no zero-fill pass)
        memset(&a, 0, sizeof(a));
        a.words = d_words;
        a.tile_dir = (d_words == ix->d_words) ? ix->d_tile_dir : nullptr;
        a.doc_lens = ix->d_doc_lens;
        a.n_docs = ix->n_docs;
        a.doc_base = ix->doc_base;
        a.queries = ix->queries.as<PhraseQuery>();
        a.stats = d_stats;
        a.out = ix->dense.as<float>();
        a.out_stride = stride;
        a.n_chunks = n_chunks;
        a.docs_per_chunk = docs_per_chunk;
        a.arena = d_arena;
        a.arena_used = d_used;
        a.arena_cap = arena_words;
        a.bm25 = p;
        a.score = score;
        a.dump = dump;
        if ((rc = launch_phrase(ix, a, Q))) return rc;
        }
        SA_CUDA(cudaMemcpyAsync(h_stats.data(), d_stats, (size_t)Q * sizeof(PhraseStats), cudaMemcpyDeviceToHost, ix->stream));
        SA_CUDA(cudaStreamSynchronize(ix->stream));
        bool again = false;
        for (u32 q = 0; q < Q; q++) {
            SA_CHECK(h_stats[q].overflow != 1, "phrase scratch arena exhausted (internal sizing error)");
            if (h_stats[q].overflow == 2) { again = true; continue; }      // dense conjunction: the search regime takes it
            step_order(pqs[q], order);
            for (u32 s : order) {
                bool actual = h_stats[q].n_inner[s] > 0 && h_stats[q].n_diff[s] == 0;
                bool guess = (pqs[q].same_guess >> s) & 1u;
                if (actual != guess) {      // later steps ran on a wrong premise: fix this one, redo
                    pqs[q].same_guess ^= 1u << s;
                    again = true;
                    break;
                }
            }
        }
        if (!again) return SA_OK;
        if (staged) {
            // The merge regime counts equal-header pairs in the candidate docs only (docs holding every term): enough to
            // CONFIRM a guess, not to derive the reference's global same-term decision from.  On any disagreement the
            // query starts over in the search regime, whose statistics cover every pair.
            staged = false;
            pqs = pqs_in;
            arena_words = full_arena_words;
            if ((rc = ix->phrase_scratch.reserve(arena_words * sizeof(u64) + 64))) return rc;
            d_used = (unsigned long long *)ix->phrase_scratch.p;
            d_arena = (u64 *)ix->phrase_scratch.p + 8;
        }
    }
    sa_set_error("same-term speculation did not converge");
    return SA_ERR_ARG;
}

// Checks the same-term speculation of one finished query; on the first wrong guess (in step
// order) flips it and returns false: the query must run again.
bool sa_phrase_guess_ok(PhraseQuery &pq, const PhraseStats &st) {
    std::vector<u32> order;
    step_order(pq, order);
    for (u32 s : order) {
        bool actual = st.n_inner[s] > 0 && st.n_diff[s] == 0;
        bool guess = (pq.same_guess >> s) & 1u;
        if (actual != guess) {
            pq.same_guess ^= 1u << s;
            return false;
        }
    }
    return true;
}

u64 sa_phrase_arena_words(const PhraseQuery &pq, u32 n_chunks) {
    u64 sum = 0;
    for (u32 t = 0; t < pq.n_terms; t++) sum += pq.len[t];
    return 6 * (sum + 2ull * n_chunks);
}

// Asynchronous launch of already planned phrase queries living in device memory; dense rows,
// stats and the arena counter must have been zeroed by the caller.
// chunks are whole tiles: returns the chunk count actually used for `wanted` chunks per query
u32 sa_phrase_chunks(const sa_index *ix, u32 wanted) {
    const u64 n_tiles = (ix->n_docs + SA_TILE_DOCS - 1) / SA_TILE_DOCS;
    const u64 tiles_per_chunk = std::max<u64>(1, (n_tiles + std::max<u32>(wanted, 1) - 1) / std::max<u32>(wanted, 1));
    return (u32)((n_tiles + tiles_per_chunk - 1) / tiles_per_chunk);
}

static u64 docs_per_chunk_of(const sa_index *ix, u32 n_chunks) {
    const u64 n_tiles = (ix->n_docs + SA_TILE_DOCS - 1) / SA_TILE_DOCS;
    return ((n_tiles + n_chunks - 1) / n_chunks) * SA_TILE_DOCS;
}

int sa_phrase_enqueue(sa_index *ix, const PhraseQuery *d_pqs, PhraseStats *d_stats, u32 Q,
                      float *dense_rows, u64 stride, u32 n_chunks, u64 *d_arena,
                      unsigned long long *d_arena_used, u64 arena_words, int score, const Bm25Params &p,
                      const TopkCtx *topk, u32 topk_row0, const PhraseSplit *split) {
    PhraseArgs a;
    memset(&a, 0, sizeof(a));
    a.words = ix->d_words;
    a.tile_dir = ix->d_tile_dir;
    a.doc_lens = ix->d_doc_lens;
    a.n_docs = ix->n_docs;
    a.doc_base = ix->doc_base;
    a.queries = d_pqs;
    a.stats = d_stats;
    a.out = dense_rows;
    a.out_stride = stride;
    a.n_chunks = n_chunks;
    a.docs_per_chunk = docs_per_chunk_of(ix, n_chunks);
    a.arena = d_arena;
    a.arena_used = d_arena_used;
    a.arena_cap = arena_words;
    a.bm25 = p;
    a.score = score;
    if (topk) a.topk = *topk;
    a.topk_row0 = topk_row0;
    if (!split || split->n_staged == 0) {
        if (split) { a.qsel = split->d_search; a.n_sel = split->n_search; }
        return launch_phrase(ix, a, split ? split->n_search : Q);
    }
    int rc;
    if (split->n_search) {                      // search regime: one CTA per (query, chunk)
        a.qsel = split->d_search;
        a.n_sel = split->n_search;
        if ((rc = launch_phrase(ix, a, split->n_search))) return rc;
    }
    a.qsel = split->d_staged;
    a.n_sel = split->n_staged;
    static const bool use_tma_pipeline = getenv("SA_PHRASE_TMA_PIPELINE") && atoi(getenv("SA_PHRASE_TMA_PIPELINE")) != 0;
    if (!use_tma_pipeline) {
        // conjunction regime: one CTA per (query, tile), like the term scan
        const unsigned n_tiles = (unsigned)((ix->n_docs + SA_TILE_DOCS - 1) / SA_TILE_DOCS);
        KernelTimer t(ix, 2);
        phrase_tile_kernel<<<dim3(split->n_staged, n_tiles), PT, 0, ix->stream>>>(a);
        SA_CUDA(cudaGetLastError());
        t.stop();
        ix->stats.phrase_kernel_launches++;
        ix->stats.total_launches++;
        return SA_OK;
    }
    // the same regime as persistent CTAs with a double-buffered TMA pipeline (kept for comparison: measured slower,
    // profiles/README.md): SA_PHRASE_TMA_PIPELINE=1
    const u32 stage_words = sa_phrase_stage_words();
    u32 ctas = 0;
    if ((rc = staged_grid(ix, stage_words, &ctas))) return rc;
    if ((rc = ix->phrase_slabs.reserve((size_t)ctas * (PT / 32) * 6 * split->slab_cap * sizeof(u64)))) return rc;
    a.qsel = split->d_staged;
    a.n_sel = split->n_staged;
    a.n_chunks = split->staged_chunks;
    a.docs_per_chunk = docs_per_chunk_of(ix, split->staged_chunks);
    a.stage_words = stage_words;
    a.slabs = ix->phrase_slabs.as<u64>();
    a.slab_cap = split->slab_cap;
    a.work_counter = (u32 *)(d_arena_used + 1);                 // zeroed with the arena counter
    const u64 n_work = (u64)a.n_sel * a.n_chunks;
    KernelTimer t(ix, 2);
    phrase_staged_kernel<<<(unsigned)std::min<u64>(ctas, n_work), PT, 2 * (size_t)stage_words * sizeof(u64), ix->stream>>>(a);
    SA_CUDA(cudaGetLastError());
    t.stop();
    ix->stats.phrase_kernel_launches++;
    ix->stats.total_launches++;
    return SA_OK;
}

// chunks of the merge regime: whole tiles, ~16 tiles each (segments are cut inside the kernel)
u32 sa_phrase_staged_chunks(const sa_index *ix) {
    const u64 n_tiles = (ix->n_docs + SA_TILE_DOCS - 1) / SA_TILE_DOCS;
    return sa_phrase_chunks(ix, (u32)std::max<u64>(1, (n_tiles + 15) / 16));
}

// scratch entries per buffer a persistent CTA needs for this query: no segment slice is longer than the largest
// per-tile slice of its terms (one-tile segments) or the staging capacity (multi-tile segments)
u64 sa_phrase_slab_cap(const sa_index *ix, const u32 *term_ids, u32 n_terms) {
    u64 cap = sa_phrase_stage_words();
    for (u32 i = 0; i < n_terms; i++)
        if (term_ids[i] != SA_NO_TERM && term_ids[i] < ix->n_terms) cap = std::max<u64>(cap, ix->h_max_tile_words[term_ids[i]]);
    return cap + 8;
}

static int phrase_common(sa_index *ix, const uint32_t *term_ids, uint32_t n_terms, uint32_t slop,
                         int score, float idf, float avg_doc_len, float k1, float b,
                         uint64_t min_payload, uint64_t max_payload, float *out_host) {
    SA_CHECK(ix && term_ids && out_host, "NULL argument");
    SA_CHECK(n_terms >= 2, "Must have at least two terms");
    SA_CHECK(n_terms <= SA_MAX_PHRASE_TERMS, "phrases longer than %d terms are not supported", SA_MAX_PHRASE_TERMS);
    std::lock_guard<std::mutex> g(ix->mu);
    SA_CUDA(cudaSetDevice(ix->device));
    if (ix->n_docs == 0) return SA_OK;
    bool missing = false;
    for (u32 i = 0; i < n_terms; i++) {
        SA_CHECK(term_ids[i] == SA_NO_TERM || term_ids[i] < ix->n_terms, "term id %u out of range", term_ids[i]);
        if (term_ids[i] == SA_NO_TERM || ix->h_len[term_ids[i]] == 0) missing = true;
    }
    if (missing || (score && avg_doc_len == 0.0f)) {
        // unknown term inside a phrase -> zeros (postings.py:705-708); with tf == 0 everywhere BM25
        // still runs over all docs in the reference, which only matters for exotic parameters
        memset(out_host, 0, (ix->rows_active ? ix->n_rows : ix->n_docs) * sizeof(float));
        if (!(score && avg_doc_len != 0.0f) || ix->rows_active) return SA_OK;
    }
    Bm25Params p;
    p.idf = idf; p.avg_doc_len = avg_doc_len; p.k1 = k1; p.b = b; p.one_minus_b = 1 - b;
    p.sparse_ok = (ix->doc_lens_nonneg && k1 > 0.0f && std::isfinite(k1) && b >= 0.0f && b < 1.0f &&
                   avg_doc_len > 0.0f && std::isfinite(avg_doc_len) && std::isfinite(idf) && idf >= 0.0f &&
                   !std::signbit(idf)) ? 1 : 0;
    const u64 stride = padded(ix->n_docs);
    int rc;
    bool raw_counts = false;          // dense row holds raw phrase freqs that still need BM25
    const bool use_payload = !(min_payload == 0 && max_payload == SA_ALL_BITS);
    const bool rows = ix->rows_active;
    SA_CHECK(!(rows && score), "score on a sliced array: call termfreqs + bm25 (the Python layer does)");
    // term lists: the index's own, or filtered copies (sliced array / min-max posn), which is what
    // the reference runs on (middle_out.py:427-437: encoder.slice per term, then the same algorithm)
    std::vector<u64> offs(n_terms), lens(n_terms);
    const u64 *d_lists = ix->d_words;
    if (!missing && (rows || use_payload)) {
        if ((rc = sa_filter_terms(ix, term_ids, n_terms, rows, min_payload, max_payload, use_payload, offs, lens))) return rc;
        d_lists = ix->filt.as<u64>();
    } else if (!missing) {
        for (u32 i = 0; i < n_terms; i++) { offs[i] = ix->h_off[term_ids[i]]; lens[i] = ix->h_len[term_ids[i]]; }
    }
    if (!missing && slop > 0) {
        // span search (phrase/spans.py + roaringish/spans.pyx): raw counts, BM25 afterwards
        std::vector<u64> dirs(n_terms, SA_NO_DIR);
        bool literal = true;
        if (d_lists == ix->d_words) {
            for (u32 i = 0; i < n_terms; i++) {
                dirs[i] = ix->h_dir_off[term_ids[i]];
                literal = literal && ix->h_first0[term_ids[i]];
            }
        } else if ((rc = sa_span_is_literal(ix, d_lists, offs.data(), lens.data(), n_terms, &literal))) {
            return rc;
        }
        if ((rc = sa_span_run(ix, d_lists, offs.data(), lens.data(), dirs.data(), n_terms, slop, literal, nullptr))) return rc;
        raw_counts = score != 0;
    } else if (!missing) {
        std::vector<PhraseQuery> pqs(1);
        PhraseQuery &pq = pqs[0];
        memset(&pq, 0, sizeof(pq));
        pq.n_terms = n_terms;
        pq.idf = idf;
        for (u32 i = 0; i < n_terms; i++) {
            pq.off[i] = offs[i];
            pq.len[i] = lens[i];
            if (d_lists == ix->d_words && ix->h_dir_off[term_ids[i]] != SA_NO_DIR) pq.dir_plus1[i] = ix->h_dir_off[term_ids[i]] + 1;
        }
        sa_phrase_plan(pq, term_ids);
        PhraseDump nodump;
        memset(&nodump, 0, sizeof(nodump));
        // raw counts first when BM25 must touch every doc
        if ((rc = sa_phrase_run_sync(ix, pqs, d_lists, score && p.sparse_ok, p, 0, nodump,
                                     d_lists == ix->d_words ? sa_phrase_slab_cap(ix, term_ids, n_terms) : 0))) return rc;
        raw_counts = score && !p.sparse_ok;
    } else {
        if ((rc = ix->dense.reserve(stride * sizeof(float)))) return rc;
        SA_CUDA(cudaMemsetAsync(ix->dense.p, 0, stride * sizeof(float), ix->stream));
        raw_counts = score && !p.sparse_ok;
    }
    if (raw_counts) {     // bm25.pyx:20-25 over every doc (tf == 0 scores +0.0 for ordinary parameters)
        unsigned blocks = (unsigned)((ix->n_docs + 255) / 256);
        bm25_dense_kernel<<<blocks, 256, 0, ix->stream>>>(ix->dense.as<float>(), ix->d_doc_lens, ix->n_docs, p);
        SA_CUDA(cudaGetLastError());
        ix->stats.total_launches++;
    }
    if (rows) return sa_gather_rows(ix, ix->dense.as<float>(), out_host);
    SA_CUDA(cudaMemcpyAsync(out_host, ix->dense.p, ix->n_docs * sizeof(float), cudaMemcpyDeviceToHost, ix->stream));
    SA_CUDA(cudaStreamSynchronize(ix->stream));
    return SA_OK;
}

extern "C" int sa_phrase_freqs(sa_index *ix, const uint32_t *term_ids, uint32_t n_terms, uint32_t slop,
                               uint64_t min_payload, uint64_t max_payload, float *out_host) {
    return phrase_common(ix, term_ids, n_terms, slop, 0, 0.0f, 1.0f, 1.0f, 0.0f, min_payload, max_payload, out_host);
}

extern "C" int sa_score_phrase(sa_index *ix, const uint32_t *term_ids, uint32_t n_terms, uint32_t slop,
                               float idf, float avg_doc_len, float k1, float b,
                               uint64_t min_payload, uint64_t max_payload, float *out_host) {
    return phrase_common(ix, term_ids, n_terms, slop, 1, idf, avg_doc_len, k1, b, min_payload, max_payload, out_host);
}

// ---------------------------------------------------------------- per-op parity exports
extern "C" int sa_op_bigram_freqs(const uint64_t *lhs, uint64_t n_lhs, const uint64_t *rhs, uint64_t n_rhs,
                                  int cont_rhs, int device,
                                  uint64_t *ids_out, float *counts_out, uint64_t *n_ids_out,
                                  uint64_t *next_out, uint64_t *n_next_out) {
    SA_CHECK(ids_out && counts_out && n_ids_out && next_out && n_next_out, "NULL argument");
    *n_ids_out = 0;
    *n_next_out = 0;
    if (n_lhs == 0 || n_rhs == 0) return SA_OK;
    SA_CHECK(lhs && rhs, "NULL argument");
    // a throw-away two-term index over one shard that covers every doc id present
    std::vector<u64> words(lhs, lhs + n_lhs);
    words.insert(words.end(), rhs, rhs + n_rhs);
    u64 max_doc = std::max(lhs[n_lhs - 1], rhs[n_rhs - 1]) >> SA_KEY_SHIFT;
    std::vector<float> dl(max_doc + 1, 1.0f);
    u64 offs[2] = {0, n_lhs}, lens[2] = {n_lhs, n_rhs};
    sa_index *ix = nullptr;
    int rc = sa_index_create(words.data(), words.size(), offs, lens, 2, dl.data(), max_doc + 1, 0, device, &ix);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> g(ix->mu);
        std::vector<PhraseQuery> pqs(1);
        PhraseQuery &pq = pqs[0];
        memset(&pq, 0, sizeof(pq));
        pq.n_terms = 2;
        pq.off[0] = 0; pq.len[0] = n_lhs;
        pq.off[1] = n_lhs; pq.len[1] = n_rhs;
        pq.mode = cont_rhs ? SA_PHRASE_MODE_LR : SA_PHRASE_MODE_RL;
        pq.same_guess = 0;
        const u64 cap = 2 * std::min(n_lhs, n_rhs) + std::max(n_lhs, n_rhs) + 8;
        DevBuf dbuf;
        rc = dbuf.reserve((2 * cap + 2) * sizeof(u64));
        if (!rc) {
            PhraseDump dump;
            dump.cont = dbuf.as<u64>();
            dump.docs = dump.cont + cap;
            dump.n_cont = dump.docs + cap;
            dump.n_docs = dump.n_cont + 1;
            cudaMemsetAsync(dump.n_cont, 0, 2 * sizeof(u64), ix->stream);
            Bm25Params p;
            memset(&p, 0, sizeof(p));
            rc = sa_phrase_run_sync(ix, pqs, ix->d_words, 0, p, 1, dump, 0);
            if (!rc) {
                u64 n[2];
                cudaMemcpy(n, dump.n_cont, 2 * sizeof(u64), cudaMemcpyDeviceToHost);
                std::vector<u64> docs(n[1]);
                cudaMemcpy(next_out, dump.cont, n[0] * sizeof(u64), cudaMemcpyDeviceToHost);
                if (n[1]) cudaMemcpy(docs.data(), dump.docs, n[1] * sizeof(u64), cudaMemcpyDeviceToHost);
                for (u64 i = 0; i < n[1]; i++) {
                    ids_out[i] = docs[i] >> 32;
                    counts_out[i] = (float)(u32)(docs[i] & 0xFFFFFFFFull);
                }
                *n_next_out = n[0];
                *n_ids_out = n[1];
            }
        }
        dbuf.release();
    }
    sa_index_destroy(ix);
    return rc;
}

// popcount64_reduce / as_dense / bm25_score on raw arrays: the term kernel on a one-term index
extern "C" int sa_op_popcount64_reduce(const uint64_t *words, uint64_t n, int device,
                                       uint64_t *keys_out, float *counts_out, uint64_t *n_out) {
    SA_CHECK(keys_out && counts_out && n_out, "NULL argument");
    *n_out = 0;
    if (n == 0) return SA_OK;
    SA_CHECK(words, "NULL argument");
    u64 min_doc = words[0] >> SA_KEY_SHIFT, max_doc = words[n - 1] >> SA_KEY_SHIFT;
    u64 nd = max_doc - min_doc + 1;
    std::vector<float> dl(nd, 1.0f), tf(nd);
    u64 off = 0, len = n;
    sa_index *ix = nullptr;
    int rc = sa_index_create(words, n, &off, &len, 1, dl.data(), nd, min_doc, device, &ix);
    if (rc) return rc;
    rc = sa_termfreqs(ix, 0, 0, SA_ALL_BITS, tf.data());
    sa_index_destroy(ix);
    if (rc) return rc;
    // docs present in the list keep their (possibly zero) count: walk the keys on the host
    u64 m = 0, last = ~0ull;
    for (u64 i = 0; i < n; i++) {
        u64 d = words[i] >> SA_KEY_SHIFT;
        if (d != last) { keys_out[m] = d; counts_out[m] = tf[d - min_doc]; m++; last = d; }
    }
    *n_out = m;
    return SA_OK;
}

extern "C" int sa_op_bm25_score(float *tf_inout, const float *doc_lens, uint64_t n, float avg_doc_len,
                                float idf, float k1, float b, int device) {
    if (n == 0) return SA_OK;
    SA_CHECK(tf_inout && doc_lens, "NULL argument");
    SA_CUDA(cudaSetDevice(device));
    float *d_tf = nullptr, *d_dl = nullptr;
    SA_CUDA(cudaMalloc(&d_tf, n * sizeof(float)));
    if (cudaMalloc(&d_dl, n * sizeof(float)) != cudaSuccess) { cudaFree(d_tf); sa_set_error("cudaMalloc failed"); return SA_ERR_NOMEM; }
    cudaMemcpy(d_tf, tf_inout, n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(d_dl, doc_lens, n * sizeof(float), cudaMemcpyHostToDevice);
    Bm25Params p;
    p.idf = idf; p.avg_doc_len = avg_doc_len; p.k1 = k1; p.b = b; p.one_minus_b = 1 - b; p.sparse_ok = 0;
    bm25_dense_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_tf, d_dl, n, p);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(tf_inout, d_tf, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_tf);
    cudaFree(d_dl);
    if (e != cudaSuccess) { sa_set_error("sa_op_bm25_score: %s", cudaGetErrorString(e)); return SA_ERR_CUDA; }
    return SA_OK;
}


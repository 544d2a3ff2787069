/*
 * sa_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the native (Cython) kernels on SearchArray's scoring
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (searcharray_b200)
 * never does.  Every function cites the reference code it restates
 * (paths relative to /root/reference).  The restatement is pinned against the
 * real reference by tests/golden/make_golden.py -> tests/golden/*.npz.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, like the reference's
 * x86-64 baseline wheels: no FMA contraction in the BM25 expression).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef int64_t i64;

/* ------------------------------------------------------------------ BM25 -- */
/* searcharray/bm25/bm25.pyx:11-25 (_bm25_score): in place, over ALL n docs,
 * operation order exactly as written there. */
void sao_bm25_score(float *tf, const float *doc_lens, float avg_doc_len,
                    float idf, float k1, float b, long n)
{
    const float one_minus_b = 1 - b;
    for (long i = 0; i < n; i++) {
        float t = tf[i];
        float norm = k1 * (one_minus_b + (b * (doc_lens[i] / avg_doc_len)));
        tf[i] = (t / (t + norm)) * idf;
    }
}

/* -------------------------------------------------------------- popcount -- */
/* searcharray/roaringish/popcount.pyx:71-80,120-122 (popcount64) */
void sao_popcount64(const u64 *arr, u64 n, u64 *out)
{
    for (u64 i = 0; i < n; i++)
        out[i] = (u64)__builtin_popcountll(arr[i]);
}

/* searcharray/roaringish/popcount.pyx:212-237,271-278 (popcount64_reduce):
 * group consecutive words by (word >> key_shift); value = sum popcount(word & mask).
 * Returns number of groups (0 for empty input, as the wrapper does). */
u64 sao_popcount64_reduce(const u64 *arr, u64 n, u64 key_shift, u64 value_mask,
                          u64 *keys_out, float *counts_out)
{
    if (n == 0) return 0;
    u64 g = 0;
    u64 cur_key = arr[0] >> key_shift;
    u64 acc = 0;
    for (u64 i = 0; i < n; i++) {
        u64 k = arr[i] >> key_shift;
        if (k != cur_key) {
            keys_out[g] = cur_key;
            counts_out[g] = (float)acc;   /* reference accumulates in f32; exact for < 2^24 */
            g++;
            cur_key = k;
            acc = 0;
        }
        acc += (u64)__builtin_popcountll(arr[i] & value_mask);
    }
    keys_out[g] = cur_key;
    counts_out[g] = (float)acc;
    return g + 1;
}

/* searcharray/roaringish/popcount.pyx:124-165 (popcount_reduce_at): ids are
 * already per-word keys; groups with a zero sum are KEPT. */
u64 sao_popcount_reduce_at(const u64 *ids, const u64 *payload, u64 n,
                           u64 *ids_out, float *counts_out)
{
    if (n == 0) return 0;
    u64 g = 0, cur = ids[0], acc = 0;
    for (u64 i = 0; i < n; i++) {
        if (ids[i] != cur) {
            ids_out[g] = cur; counts_out[g] = (float)acc; g++;
            cur = ids[i]; acc = 0;
        }
        acc += (u64)__builtin_popcountll(payload[i]);
    }
    ids_out[g] = cur; counts_out[g] = (float)acc;
    return g + 1;
}

/* searcharray/roaringish/popcount.pyx:168-204 (key_sum_over) */
u64 sao_key_sum_over(const u64 *ids, const u64 *count, u64 n,
                     u64 *ids_out, float *counts_out)
{
    if (n == 0) return 0;
    u64 g = 0, cur = ids[0], acc = 0;
    for (u64 i = 0; i < n; i++) {
        if (ids[i] != cur) {
            ids_out[g] = cur; counts_out[g] = (float)acc; g++;
            cur = ids[i]; acc = 0;
        }
        acc += count[i];
    }
    ids_out[g] = cur; counts_out[g] = (float)acc;
    return g + 1;
}

/* ---------------------------------------------------------------- dense -- */
/* searcharray/roaringish/roaringish_ops.pyx:84-98 (as_dense) +
 * searcharray/roaringish/scatter_assign.h:8-29 (scatter_naive): out is
 * pre-zeroed by the caller (np.zeros), later duplicates win. */
void sao_scatter(float *out, const u64 *idx, const float *val, u64 n)
{
    for (u64 i = 0; i < n; i++) out[idx[i]] = val[i];
}

/* searcharray/roaringish/roaringish_ops.pyx:46-68 (payload_slice): note the
 * comparison is on (word & msb_mask) UNSHIFTED (SURVEY quirk vi). */
u64 sao_payload_slice(const u64 *arr, u64 n, u64 msb_mask, u64 lo, u64 hi, u64 *out)
{
    u64 m = 0;
    for (u64 i = 0; i < n; i++) {
        u64 v = arr[i] & msb_mask;
        if (v >= lo && v <= hi) out[m++] = arr[i];
    }
    return m;
}

/* --------------------------------------------------------------- unique -- */
/* searcharray/roaringish/unique.pyx:87-104,139-145 (unique, rshift>0 emits the
 * shifted value; rshift==0 emits the value) -- run-length dedup of a sorted array. */
u64 sao_unique(const u64 *arr, u64 n, u64 rshift, u64 *out)
{
    u64 m = 0, i = 0;
    while (i < n) {
        u64 v = arr[i] >> rshift;
        out[m++] = v;
        i++;
        while (i < n && (arr[i] >> rshift) == v) i++;
    }
    return m;
}

/* ------------------------------------------------------------ intersect -- */
/* Exponential ("galloping") advance used by all the intersect variants, restating
 * the inner while-loops of searcharray/roaringish/intersect.pyx:46-55: step 1,2,4..
 * while the probe is still below the target, then back off half of the last stride
 * (which lands on an element known to be below the target, or the start). */
#define GALLOP_BELOW(ptr, end, cond_lt_target)                   \
    do {                                                         \
        u64 stride_ = 1;                                         \
        while ((ptr) < (end) && (cond_lt_target)) {              \
            (ptr) += stride_;                                    \
            stride_ <<= 1;                                       \
        }                                                        \
        (ptr) -= (stride_ >> 1);                                 \
    } while (0)

/* searcharray/roaringish/intersect.pyx:32-74 (_gallop_intersect_drop) + :278-308.
 * Emits index pairs of equal (x & mask); one pair per distinct masked value. */
u64 sao_intersect_drop(const u64 *lhs, u64 nl, const u64 *rhs, u64 nr, u64 mask,
                       u64 *lhs_idx, u64 *rhs_idx)
{
    const u64 *l = lhs, *r = rhs, *le = lhs + nl, *re = rhs + nr;
    u64 m = 0;
    u64 last = (u64)-1;
    while (l < le && r < re) {
        GALLOP_BELOW(l, le, (*l & mask) < (*r & mask));
        GALLOP_BELOW(r, re, (*r & mask) < (*l & mask));
        u64 a = *l & mask, b = *r & mask;
        if (a < b) l++;
        else if (b < a) r++;
        else {
            if ((last & mask) != a) {
                lhs_idx[m] = (u64)(l - lhs);
                rhs_idx[m] = (u64)(r - rhs);
                last = *l;
                m++;
            }
            l++; r++;
        }
    }
    return m;
}

/* searcharray/roaringish/intersect.pyx:77-128 (_gallop_intersect_keep) + :310-320:
 * all lhs indices and all rhs indices whose masked value is in both (dups kept). */
void sao_intersect_keep(const u64 *lhs, u64 nl, const u64 *rhs, u64 nr, u64 mask,
                        u64 *lhs_idx, u64 *rhs_idx, u64 *n_lhs_out, u64 *n_rhs_out)
{
    const u64 *l = lhs, *r = rhs, *le = lhs + nl, *re = rhs + nr;
    u64 ml = 0, mr = 0;
    while (l < le && r < re) {
        GALLOP_BELOW(l, le, (*l & mask) < (*r & mask));
        GALLOP_BELOW(r, re, (*r & mask) < (*l & mask));
        u64 a = *l & mask, b = *r & mask;
        if (a < b) l++;
        else if (b < a) r++;
        else {
            while (l < le && (*l & mask) == a) lhs_idx[ml++] = (u64)(l++ - lhs);
            while (r < re && (*r & mask) == a) rhs_idx[mr++] = (u64)(r++ - rhs);
        }
    }
    *n_lhs_out = ml;
    *n_rhs_out = mr;
}

/* searcharray/roaringish/intersect.pyx:131-190 (_gallop_adjacent) + :323-343:
 * pairs with (lhs & mask) == (rhs & mask) - delta, delta = lowest set bit of mask.
 * rhs entries whose masked value is 0 are skipped first (reference :151-152). */
u64 sao_adjacent(const u64 *lhs, u64 nl, const u64 *rhs, u64 nr, u64 mask,
                 u64 *lhs_idx, u64 *rhs_idx)
{
    const u64 delta = mask & (~mask + 1);
    const u64 *l = lhs, *r = rhs, *le = lhs + nl, *re = rhs + nr;
    u64 m = 0, last = (u64)-1;
    while (r < re && (*r & mask) == 0) r++;
    while (l < le && r < re) {
        GALLOP_BELOW(l, le, (*l & mask) < ((*r & mask) - delta));
        GALLOP_BELOW(r, re, ((*r & mask) - delta) < (*l & mask));
        u64 a = *l & mask, b = (*r & mask) - delta;
        if (a < b) l++;
        else if (b < a) r++;
        else {
            if ((last & mask) != a) {
                lhs_idx[m] = (u64)(l - lhs);
                rhs_idx[m] = (u64)(r - rhs);
                last = *l;
                m++;
            }
            l++; r++;
        }
    }
    return m;
}

/* searcharray/roaringish/intersect.pyx:213-275 (_gallop_int_and_adj_drop) + :346-390:
 * one pass producing both the equal-(x&mask) pairs and the pairs with
 * (lhs&mask)+delta == (rhs&mask).  Returns #equal pairs, *n_adj_out = #adjacent. */
u64 sao_intersect_with_adjacents(const u64 *lhs, u64 nl, const u64 *rhs, u64 nr, u64 mask,
                                 u64 *lhs_idx, u64 *rhs_idx,
                                 u64 *adj_lhs_idx, u64 *adj_rhs_idx, u64 *n_adj_out)
{
    const u64 delta = mask & (~mask + 1);
    const u64 *l = lhs, *r = rhs, *le = lhs + nl, *re = rhs + nr;
    u64 m = 0, ma = 0;
    u64 last = (u64)-1, last_adj = (u64)-1;
    while (l < le && r < re) {
        if ((*l & mask) != (*r & mask)) {
            GALLOP_BELOW(l, le, ((*l & mask) + delta) < (*r & mask));
            GALLOP_BELOW(r, re, (*r & mask) < ((*l & mask) + delta));
        }
        u64 a = *l & mask, b = *r & mask;
        if (a + delta == b) {
            if ((last_adj & mask) != a) {
                adj_lhs_idx[ma] = (u64)(l - lhs);
                adj_rhs_idx[ma] = (u64)(r - rhs);
                last_adj = *l;
                ma++;
            }
            l++;
        } else if (a < b) l++;
        else if (b < a) r++;
        else {
            if ((last & mask) != a) {
                lhs_idx[m] = (u64)(l - lhs);
                rhs_idx[m] = (u64)(r - rhs);
                last = *l;
                m++;
            }
            r++;
        }
    }
    *n_adj_out = ma;
    return m;
}

/* ---------------------------------------------------------------- merge -- */
/* searcharray/roaringish/merge.pyx:54-134,137-158 (merge): sorted two-way merge;
 * equal values are emitted twice unless drop_duplicates. */
u64 sao_merge(const u64 *lhs, u64 nl, const u64 *rhs, u64 nr, int drop, u64 *out)
{
    u64 i = 0, j = 0, m = 0;
    while (i < nl && j < nr) {
        if (lhs[i] < rhs[j]) out[m++] = lhs[i++];
        else if (rhs[j] < lhs[i]) out[m++] = rhs[j++];
        else {
            out[m++] = lhs[i];
            if (!drop) out[m++] = rhs[j];
            i++; j++;
        }
    }
    while (j < nr) out[m++] = rhs[j++];
    while (i < nl) out[m++] = lhs[i++];
    return m;
}

/* searcharray/roaringish/merge.pyx:161-232 (sort_merge_counts): merge two sorted
 * (id, count) lists, adding counts on equal ids. */
u64 sao_sort_merge_counts(const u64 *lids, const float *lcnt, u64 nl,
                          const u64 *rids, const float *rcnt, u64 nr,
                          u64 *ids_out, float *cnt_out)
{
    u64 i = 0, j = 0, m = 0;
    while (i < nl && j < nr) {
        if (lids[i] < rids[j]) { ids_out[m] = lids[i]; cnt_out[m] = lcnt[i]; i++; }
        else if (rids[j] < lids[i]) { ids_out[m] = rids[j]; cnt_out[m] = rcnt[j]; j++; }
        else { ids_out[m] = lids[i]; cnt_out[m] = lcnt[i] + rcnt[j]; i++; j++; }
        m++;
    }
    for (; i < nl; i++, m++) { ids_out[m] = lids[i]; cnt_out[m] = lcnt[i]; }
    for (; j < nr; j++, m++) { ids_out[m] = rids[j]; cnt_out[m] = rcnt[j]; }
    return m;
}

/* ---------------------------------------------------------------- spans -- */
/* searcharray/roaringish/spans.pyx:70-186: the live-span table (capacity 512). */
#define SPAN_CAP 512
typedef struct {
    u64 terms[SPAN_CAP];
    u64 posns[SPAN_CAP];
    i64 beg[SPAN_CAP];
    i64 end[SPAN_CAP];
    u64 cursor;
} span_table;

static i64 iabs64(i64 v) { return v < 0 ? -v : v; }
static i64 span_width(const span_table *s, u64 i) { return iabs64(s->end[i] - s->beg[i]); }

/* spans.pyx:141-154 (_compact_spans): keep spans with width <= max_width and >=1 term */
static void compact_spans(span_table *s, u64 max_width)
{
    u64 w = 0;
    for (u64 i = 0; i < s->cursor; i++) {
        if ((u64)span_width(s, i) > max_width) continue;
        if (__builtin_popcountll(s->terms[i]) > 0) {
            s->terms[w] = s->terms[i]; s->posns[w] = s->posns[i];
            s->beg[w] = s->beg[i]; s->end[w] = s->end[i];
            w++;
        }
    }
    /* the reference builds a fresh zeroed table; slots >= w are never read before
     * being rewritten, so zeroing them is not observable -- do it anyway. */
    for (u64 i = w; i < SPAN_CAP; i++) { s->terms[i] = 0; s->posns[i] = 0; s->beg[i] = 0; s->end[i] = 0; }
    s->cursor = w;
}

/* spans.pyx:157-186 (_collect_spans): number of complete, narrow-enough spans after
 * first-come overlap replacement. */
static u64 collect_spans(const span_table *s, u64 num_terms, u64 max_width)
{
    static __thread span_table coll;
    coll.cursor = 0;
    for (u64 i = 0; i < s->cursor; i++) {
        u64 nt = (u64)__builtin_popcountll(s->terms[i]);
        u64 np = (u64)__builtin_popcountll(s->posns[i]);
        int complete = (nt == num_terms) || (np == num_terms);
        if (!(complete && (u64)span_width(s, i) < max_width)) continue;
        i64 new_w = span_width(s, i);
        int overlaps = 0;
        for (u64 c = 0; c < coll.cursor; c++) {
            if (s->beg[i] <= coll.end[c] && s->end[i] >= coll.beg[c]) {
                i64 cw = iabs64(coll.end[c] - coll.beg[c]);
                if (new_w < cw) {
                    coll.terms[c] = s->terms[i]; coll.posns[c] = s->posns[i];
                    coll.beg[c] = s->beg[i]; coll.end[c] = s->end[i];
                    overlaps = 1;
                    break;
                }
            }
        }
        if (!overlaps) {
            u64 c = coll.cursor;
            /* the reference table has 512 slots and no bound check here; collected
             * spans never exceed live spans (<=512) so this cannot overflow. */
            coll.terms[c] = s->terms[i]; coll.posns[c] = s->posns[i];
            coll.beg[c] = s->beg[i]; coll.end[c] = s->end[i];
            coll.cursor++;
        }
    }
    return coll.cursor;
}

/* searcharray/roaringish/spans.pyx:189-319 (_span_freqs).  posns = all terms' (already
 * header-sliced) words concatenated, lengths[t]..lengths[t+1] the slice of term t.
 * Results are appended to (keys_out, counts_out) in first-touch order, accumulating
 * on a repeated key exactly like the reference's Counter (phrase/spans.py:176,186-187).
 * `posns` must have one readable padding word after the end (the reference reads
 * posns[lengths[t+1]] when a term's slice is exhausted, :199; for the last term that is
 * one past the array -- the caller pads with a 0 word).  *n_undefined counts the times
 * the reference would have written past its 512-slot table (undefined behaviour there;
 * results for those docs are NOT comparable with the reference). */
u64 sao_span_freqs(const u64 *posns, const u64 *lengths, u64 num_terms, u64 slop,
                   u64 key_mask, u64 header_mask, u64 key_bits, u64 lsb_bits,
                   u64 *keys_out, float *counts_out, u64 *n_undefined)
{
    static __thread span_table spans;
    const u64 payload_mask = ~header_mask;
    const u64 payload_msb_mask = header_mask & ~key_mask;
    const u64 max_span_width = num_terms + slop;
    u64 curr_idx[64];
    u64 sum_popcount[64];
    u64 n_out = 0;
    u64 curr_key = 0, last_key = 0, max_key_seen = 0;
    int full = 0;

    memset(&spans, 0, sizeof(spans));
    for (u64 t = 0; t < num_terms; t++) curr_idx[t] = lengths[t];

    while (curr_idx[0] < lengths[1]) {
        for (u64 t = 0; t < num_terms; t++) {
            curr_key = (posns[curr_idx[t]] & key_mask) >> (64 - key_bits);
            sum_popcount[t] = 0;
            while (curr_idx[t] < lengths[t + 1]) {
                last_key = curr_key;
                u64 word = posns[curr_idx[t]];
                u64 payload_base = ((word & payload_msb_mask) >> lsb_bits) * lsb_bits;
                u64 bits = word & payload_mask;
                const u64 term_bit = (u64)1 << t;
                sum_popcount[t] += (u64)__builtin_popcountll(bits);

                while (bits != 0) {
                    i64 set_idx = __builtin_ctzll(bits);
                    bits &= bits - 1;
                    i64 posn = set_idx + (i64)payload_base;
                    /* spans.pyx:107-108 `1 << (curr_posn % 64)` is compiled as a 32-bit int
                     * shift (generated C: `(1 << (__pyx_v_curr_posn % 64))`), i.e. the count
                     * is taken mod 32 and the int result is sign-extended into the u64: a
                     * position with posn%32 == 31 sets bits 31..63.  Parity needs this. */
                    u64 posn_bit = (u64)(i64)(int32_t)((uint32_t)1 << ((posn % 64) & 31));

                    u64 fresh = spans.cursor;
                    if (fresh >= SPAN_CAP) {
                        /* Reference writes slot 512 of a 512-slot table here (undefined
                         * behaviour, reachable only when the table is still full after
                         * compaction in a term's LAST candidate doc).  The oracle defines
                         * it as: stop consuming, fall back to the `full` estimate. */
                        full = 1;
                        if (n_undefined) (*n_undefined)++;
                        break;
                    }
                    spans.terms[fresh] = term_bit;
                    spans.posns[fresh] = posn_bit;
                    spans.beg[fresh] = posn;
                    spans.end[fresh] = posn;
                    spans.cursor++;

                    for (u64 s = 0; s < fresh; s++) {
                        u64 nt_before = (u64)__builtin_popcountll(spans.terms[s]);
                        u64 np_before = (u64)__builtin_popcountll(spans.posns[s]);
                        if (nt_before < num_terms && np_before == num_terms) continue;
                        spans.terms[s] |= term_bit;
                        u64 nt_now = (u64)__builtin_popcountll(spans.terms[s]);
                        if (nt_now > nt_before) {
                            spans.posns[s] |= posn_bit;
                            u64 np_now = (u64)__builtin_popcountll(spans.posns[s]);
                            i64 proposed = iabs64(posn - spans.beg[s]);
                            if (np_before == np_now || (u64)proposed > max_span_width) {
                                spans.terms[s] &= ~term_bit;
                                continue;
                            }
                            if (spans.cursor < SPAN_CAP) {
                                u64 c = spans.cursor;
                                spans.terms[c] = spans.terms[s];
                                spans.posns[c] = spans.posns[s] & ~posn_bit;
                                spans.beg[c] = spans.beg[s];
                                spans.end[c] = spans.end[s];
                                spans.cursor++;
                                full = 0;
                            } else {
                                full = 1;
                            }
                            spans.end[s] = posn;
                        }
                    }
                    if (spans.cursor >= SPAN_CAP) break;
                }
                curr_idx[t]++;
                if (curr_idx[t] < lengths[t + 1])
                    curr_key = (posns[curr_idx[t]] & key_mask) >> (64 - key_bits);
                if (spans.cursor >= SPAN_CAP) {
                    compact_spans(&spans, max_span_width);
                    if (spans.cursor >= SPAN_CAP) {
                        for (u64 i = curr_idx[t]; i < lengths[t + 1]; i++) {
                            curr_key = (posns[i] & key_mask) >> (64 - key_bits);
                            if (curr_key != last_key) { curr_idx[t] = i; break; }
                        }
                    }
                }
                if (curr_key != last_key) break;
            }
        }

        u64 add;
        if (full) {
            u64 mn = 0;
            for (u64 t = 0; t < num_terms; t++)
                if (mn == 0 || sum_popcount[t] < mn) mn = sum_popcount[t];
            add = mn;
        } else {
            add = collect_spans(&spans, num_terms, max_span_width);
        }
        /* Counter semantics: phrase_freqs[last_key] += add (key created even if add==0).
         * Keys normally arrive ascending, so only fall back to a scan when they do not. */
        u64 k = n_out;
        if (n_out > 0 && keys_out[n_out - 1] == last_key) k = n_out - 1;
        else if (n_out > 0 && last_key <= max_key_seen)
            for (k = 0; k < n_out; k++) if (keys_out[k] == last_key) break;
        if (k == n_out) { keys_out[n_out] = last_key; counts_out[n_out] = 0; n_out++; }
        counts_out[k] += (float)add;
        if (last_key > max_key_seen) max_key_seen = last_key;

        memset(&spans, 0, sizeof(spans));
        full = 0;
    }
    return n_out;
}

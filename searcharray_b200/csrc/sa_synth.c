/* sa_synth.c -- seeded synthetic MSMARCO-shaped corpora, generated directly as roaringish postings.
 *
 * Benchmark / test DATA infrastructure (host, plain C + pthreads): not part of the scoring path.
 * tokenising ~5e8 tokens in Python would take hours, so the corpus is
 * produced in the index's upload format (one sorted, header-unique uint64 word list per term;
 * word = doc(28b) | posn/18 (18b) | bitmap of posn%18 (18b), reference roaringish.py:30-35,93-142)
 * and injected the way SearchArray.index injects its build (reference postings.py:293-299).
 *
 * Determinism: every (field, stream, term-or-phrase, block) owns a counter-based random stream
 * (splitmix64 seeded from the key), so the corpus does not depend on the number of threads, and a
 * rank of a 1/2/4/8-GPU run generates exactly ITS doc range of the same global corpus.
 *
 *   doc_lens[d]  = floor(clip(exp(mu + sigma * z), lo, hi))
 *   term t       : each doc holds it with probability p_t; tf = min(Geometric(0.6), 8, doc_len);
 *                  positions uniform in [0, doc_len), duplicates merged
 *   phrase g     : each doc gets the phrase planted with probability plant_p_g at a random start
 *                  (slot s of the phrase -> position start + s, plus, for `gapped` phrases, up to two
 *                  extra one-token gaps in every second plant: matches only with slop)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define MAX_TF 8
#define MAX_PHRASE 8

typedef struct {
    uint64_t s;
} rng_t;

/* ---- a tiny dynamic-schedule parallel-for over pthreads (no OpenMP runtime dependency) ---- */
typedef void (*task_fn)(void *ctx, uint64_t i);
typedef struct {
    task_fn fn;
    void *ctx;
    uint64_t n;
    uint64_t next;       /* atomic */
} pf_t;

static void *pf_worker(void *arg) {
    pf_t *P = (pf_t *)arg;
    for (;;) {
        const uint64_t i = __atomic_fetch_add(&P->next, 1, __ATOMIC_RELAXED);
        if (i >= P->n) break;
        P->fn(P->ctx, i);
    }
    return NULL;
}

static void parallel_for(uint64_t n, task_fn fn, void *ctx, int n_threads) {
    pf_t P = {fn, ctx, n, 0};
    if (n_threads > (int)n) n_threads = (int)n;
    if (n_threads <= 1) { pf_worker(&P); return; }
    pthread_t *th = (pthread_t *)malloc(n_threads * sizeof(pthread_t));
    int started = 0;
    for (int t = 0; t < n_threads - 1; t++)
        if (pthread_create(&th[started], NULL, pf_worker, &P) == 0) started++;
    pf_worker(&P);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
}

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline rng_t rng_make(uint64_t seed, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    rng_t r;
    uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ull);
    h = mix64(h ^ (a + 0x9E3779B97F4A7C15ull));
    h = mix64(h ^ (b + 0x3C6EF372FE94F82Aull));
    h = mix64(h ^ (c + 0xDAA66D2C7DDF743Full));
    h = mix64(h ^ (d + 0x78DDE6E5FD29F054ull));
    r.s = h;
    return r;
}

static inline uint64_t rng_next(rng_t *r) {
    r->s += 0x9E3779B97F4A7C15ull;
    return mix64(r->s);
}

/* uniform in (0, 1] (never 0: safe under log) */
static inline double rng_u(rng_t *r) { return ((double)(rng_next(r) >> 11) + 1.0) * (1.0 / 9007199254740992.0); }

typedef struct {
    uint64_t *docs;      /* absolute doc id per plant */
    uint32_t *pos;       /* [n][n_terms] positions per slot */
    uint64_t n;
} plant_t;

typedef struct {
    uint64_t *words;
    uint64_t n, cap;
} task_out_t;

typedef struct sa_synth {
    uint64_t seed, field_key, n_docs;
    uint32_t n_blocks, blk_lo, blk_hi;
    uint32_t n_terms, n_phrases;
    float *doc_lens;         /* this shard's docs */
    uint64_t doc_lo, doc_hi;
    task_out_t *out;         /* [n_terms][n_shard_blocks] */
    uint64_t *term_len;      /* [n_terms] */
} sa_synth;

static inline uint64_t block_bound(uint64_t n_docs, uint32_t n_blocks, uint32_t b) {
    return (uint64_t)(((__uint128_t)n_docs * b) / n_blocks);
}

static void gen_doc_lens_block(uint64_t seed, uint64_t field_key, uint64_t n_docs, uint32_t n_blocks, uint32_t b,
                               double mu, double sigma, float lo, float hi, float *out) {
    const uint64_t d0 = block_bound(n_docs, n_blocks, b), d1 = block_bound(n_docs, n_blocks, b + 1);
    rng_t r = rng_make(seed, field_key, 0, b, 0);
    for (uint64_t d = d0; d < d1; d++) {
        /* Box-Muller, one normal per doc (the sine half is dropped: simpler and stream-stable) */
        const double u1 = rng_u(&r), u2 = rng_u(&r);
        const double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
        double v = exp(mu + sigma * z);
        if (v < lo) v = lo;
        if (v > hi) v = hi;
        out[d - d0] = (float)floor(v);
    }
}

static int cmp_u64(const void *a, const void *b) {
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

static inline void out_push(task_out_t *o, uint64_t w) {
    if (o->n == o->cap) {
        o->cap = o->cap ? o->cap + (o->cap >> 1) + 64 : 1024;
        o->words = (uint64_t *)realloc(o->words, o->cap * sizeof(uint64_t));
    }
    o->words[o->n++] = w;
}

/* positions of one doc (unsorted, may repeat) -> words */
static inline void emit_doc(task_out_t *o, uint64_t doc, uint32_t *pos, int n) {
    for (int i = 1; i < n; i++) {            /* insertion sort, n is tiny */
        uint32_t v = pos[i];
        int j = i - 1;
        while (j >= 0 && pos[j] > v) { pos[j + 1] = pos[j]; j--; }
        pos[j + 1] = v;
    }
    uint64_t cur_hdr = ~0ull, bits = 0;
    for (int i = 0; i < n; i++) {
        const uint64_t hdr = (doc << 36) | ((uint64_t)(pos[i] / 18) << 18);
        if (hdr != cur_hdr) {
            if (cur_hdr != ~0ull) out_push(o, cur_hdr | bits);
            cur_hdr = hdr;
            bits = 0;
        }
        bits |= 1ull << (pos[i] % 18);
    }
    if (cur_hdr != ~0ull) out_push(o, cur_hdr | bits);
}

typedef struct {
    sa_synth *S;
    double len_mu, len_sigma;
    float len_lo, len_hi;
    const double *term_p;
    const uint32_t *ph_start, *ph_terms, *ph_gapped;
    const double *ph_plant_p;
    plant_t *plants;                 /* [n_phrases][nb] */
    const uint32_t *mem_start, *mem_g, *mem_s, *order;
    uint32_t nb;
} ctx_t;

static void task_doc_lens(void *vc, uint64_t i) {
    ctx_t *C = (ctx_t *)vc;
    sa_synth *S = C->S;
    const uint32_t b = S->blk_lo + (uint32_t)i;
    gen_doc_lens_block(S->seed, S->field_key, S->n_docs, S->n_blocks, b, C->len_mu, C->len_sigma, C->len_lo, C->len_hi,
                       S->doc_lens + (block_bound(S->n_docs, S->n_blocks, b) - S->doc_lo));
}

/* planted phrase occurrences of one (phrase, block) */
static void task_plants(void *vc, uint64_t i) {
    ctx_t *C = (ctx_t *)vc;
    sa_synth *S = C->S;
    const uint32_t g = (uint32_t)(i / C->nb), bi = (uint32_t)(i % C->nb), b = S->blk_lo + bi;
    const uint32_t nt = C->ph_start[g + 1] - C->ph_start[g];
    const double p = C->ph_plant_p[g];
    plant_t *P = &C->plants[(size_t)g * C->nb + bi];
    if (!(p > 0.0) || nt == 0 || nt > MAX_PHRASE) return;
    const uint64_t d0 = block_bound(S->n_docs, S->n_blocks, b), d1 = block_bound(S->n_docs, S->n_blocks, b + 1);
    rng_t r = rng_make(S->seed, S->field_key, 2, g, b);
    const double lq = log(1.0 - (p < 0.999999 ? p : 0.999999));
    uint64_t cap = (uint64_t)((double)(d1 - d0) * p * 1.2) + 64, n = 0;
    P->docs = (uint64_t *)malloc(cap * sizeof(uint64_t));
    P->pos = (uint32_t *)malloc(cap * nt * sizeof(uint32_t));
    uint64_t d = d0;
    for (;;) {
        const double gskip = floor(log(rng_u(&r)) / lq);
        if (gskip >= (double)(d1 - d)) break;
        d += (uint64_t)gskip;
        const double us = rng_u(&r), ug = rng_u(&r);
        uint32_t gap[MAX_PHRASE];
        uint32_t extra = 0;
        memset(gap, 0, sizeof(gap));
        if (C->ph_gapped[g] && nt >= 2 && ug < 0.5) {
            /* one or two one-token gaps somewhere between the slots (a slop <= 2 match only) */
            const uint64_t x = rng_next(&r);
            gap[1 + (x % (nt - 1))] += 1;
            extra = 1;
            if ((x >> 32) & 1) { gap[1 + ((x >> 8) % (nt - 1))] += 1; extra = 2; }
        }
        const uint32_t dl = (uint32_t)S->doc_lens[d - S->doc_lo];
        const uint32_t span = nt + extra;
        if (dl >= span) {
            if (n == cap) {
                cap += (cap >> 1) + 64;
                P->docs = (uint64_t *)realloc(P->docs, cap * sizeof(uint64_t));
                P->pos = (uint32_t *)realloc(P->pos, cap * nt * sizeof(uint32_t));
            }
            uint32_t start = (uint32_t)(us * (double)(dl - span + 1));
            if (start > dl - span) start = dl - span;
            uint32_t at = start;
            for (uint32_t s = 0; s < nt; s++) {
                at += gap[s];
                P->pos[n * nt + s] = at + s;
            }
            P->docs[n++] = d;
        }
        d++;
        if (d >= d1) break;
    }
    P->n = n;
}

/* postings of one (term, block) */
static void task_postings(void *vc, uint64_t i) {
    ctx_t *C = (ctx_t *)vc;
    sa_synth *S = C->S;
    const uint32_t nb = C->nb;
    const uint32_t t = C->order[i / nb], bi = (uint32_t)(i % nb), b = S->blk_lo + bi;
    const double p = C->term_p[t];
    task_out_t *o = &S->out[(size_t)t * nb + bi];
    const uint64_t d0 = block_bound(S->n_docs, S->n_blocks, b), d1 = block_bound(S->n_docs, S->n_blocks, b + 1);
    /* planted keys (doc << 18 | pos) of every (phrase, slot) this term fills, sorted */
    uint64_t n_pl = 0;
    for (uint32_t m = C->mem_start[t]; m < C->mem_start[t + 1]; m++) n_pl += C->plants[(size_t)C->mem_g[m] * nb + bi].n;
    uint64_t *pl = NULL;
    if (n_pl) {
        pl = (uint64_t *)malloc(n_pl * sizeof(uint64_t));
        uint64_t at = 0;
        for (uint32_t m = C->mem_start[t]; m < C->mem_start[t + 1]; m++) {
            const uint32_t g = C->mem_g[m];
            const plant_t *P = &C->plants[(size_t)g * nb + bi];
            const uint32_t nt = C->ph_start[g + 1] - C->ph_start[g];
            for (uint64_t k = 0; k < P->n; k++) pl[at++] = (P->docs[k] << 18) | P->pos[k * nt + C->mem_s[m]];
        }
        qsort(pl, n_pl, sizeof(uint64_t), cmp_u64);
    }
    o->cap = (uint64_t)((double)(d1 - d0) * p * 1.8) + n_pl + 256;
    o->words = (uint64_t *)malloc(o->cap * sizeof(uint64_t));
    o->n = 0;
    rng_t r = rng_make(S->seed, S->field_key, 1, t, b);
    uint64_t pi = 0;
    uint32_t pos[MAX_TF + 64];
    uint64_t d = d0;
    const int have_p = p > 0.0;
    const double lq = have_p ? log(1.0 - (p < 0.999999 ? p : 0.999999)) : -1.0;
    const double l04 = log(0.4);
    for (;;) {
        uint64_t dr = d1;                                  /* next random doc */
        if (have_p && d < d1) {
            const double g = floor(log(rng_u(&r)) / lq);
            dr = g >= (double)(d1 - d) ? d1 : d + (uint64_t)g;
        }
        /* planted docs before the next random doc */
        while (pi < n_pl && (pl[pi] >> 18) < dr) {
            const uint64_t doc = pl[pi] >> 18;
            int n = 0;
            while (pi < n_pl && (pl[pi] >> 18) == doc) { if (n < MAX_TF + 64) pos[n++] = (uint32_t)(pl[pi] & 0x3FFFF); pi++; }
            emit_doc(o, doc, pos, n);
        }
        if (dr >= d1) break;
        const uint32_t dl = (uint32_t)S->doc_lens[dr - S->doc_lo];
        uint32_t tf = 1 + (uint32_t)floor(log(rng_u(&r)) / l04);
        if (tf > MAX_TF) tf = MAX_TF;
        if (tf > dl) tf = dl;
        int n = 0;
        for (uint32_t k = 0; k < tf; k++) {
            uint32_t q = (uint32_t)(rng_u(&r) * (double)dl);
            if (q >= dl) q = dl - 1;
            pos[n++] = q;
        }
        while (pi < n_pl && (pl[pi] >> 18) == dr) { if (n < MAX_TF + 64) pos[n++] = (uint32_t)(pl[pi] & 0x3FFFF); pi++; }
        if (n) emit_doc(o, dr, pos, n);
        d = dr + 1;
    }
    free(pl);
}

sa_synth *sa_synth_run(uint64_t seed, uint64_t field_key, uint64_t n_docs, uint32_t n_blocks,
                       uint32_t blk_lo, uint32_t blk_hi,
                       double len_mu, double len_sigma, float len_lo, float len_hi,
                       uint32_t n_terms, const double *term_p,
                       uint32_t n_phrases, const uint32_t *ph_start, const uint32_t *ph_terms,
                       const double *ph_plant_p, const uint32_t *ph_gapped, int n_threads) {
    sa_synth *S = (sa_synth *)calloc(1, sizeof(sa_synth));
    S->seed = seed; S->field_key = field_key; S->n_docs = n_docs;
    S->n_blocks = n_blocks; S->blk_lo = blk_lo; S->blk_hi = blk_hi;
    S->n_terms = n_terms; S->n_phrases = n_phrases;
    S->doc_lo = block_bound(n_docs, n_blocks, blk_lo);
    S->doc_hi = block_bound(n_docs, n_blocks, blk_hi);
    const uint32_t nb = blk_hi - blk_lo;
    S->doc_lens = (float *)malloc((S->doc_hi - S->doc_lo + 1) * sizeof(float));
    S->out = (task_out_t *)calloc((size_t)n_terms * nb + 1, sizeof(task_out_t));
    S->term_len = (uint64_t *)calloc(n_terms + 1, sizeof(uint64_t));
    ctx_t C;
    memset(&C, 0, sizeof(C));
    C.S = S; C.len_mu = len_mu; C.len_sigma = len_sigma; C.len_lo = len_lo; C.len_hi = len_hi;
    C.term_p = term_p; C.ph_start = ph_start; C.ph_terms = ph_terms; C.ph_gapped = ph_gapped; C.ph_plant_p = ph_plant_p;
    C.nb = nb;
    C.plants = (plant_t *)calloc((size_t)n_phrases * nb + 1, sizeof(plant_t));
    if (n_threads < 1) n_threads = 1;

    parallel_for(nb, task_doc_lens, &C, n_threads);                        /* phase 0 */
    parallel_for((uint64_t)n_phrases * nb, task_plants, &C, n_threads);   /* phase A */

    /* term -> (phrase, slot) memberships */
    uint32_t *mem_start = (uint32_t *)calloc(n_terms + 2, sizeof(uint32_t));
    const uint32_t n_slots_total = n_phrases ? ph_start[n_phrases] : 0;
    for (uint32_t i = 0; i < n_slots_total; i++) mem_start[ph_terms[i] + 1]++;
    for (uint32_t t = 0; t < n_terms; t++) mem_start[t + 1] += mem_start[t];
    uint32_t *mem_g = (uint32_t *)malloc((n_slots_total + 1) * sizeof(uint32_t));
    uint32_t *mem_s = (uint32_t *)malloc((n_slots_total + 1) * sizeof(uint32_t));
    {
        uint32_t *fill = (uint32_t *)calloc(n_terms + 1, sizeof(uint32_t));
        for (uint32_t g = 0; g < n_phrases; g++)
            for (uint32_t i = ph_start[g]; i < ph_start[g + 1]; i++) {
                const uint32_t t = ph_terms[i], at = mem_start[t] + fill[t]++;
                mem_g[at] = g;
                mem_s[at] = i - ph_start[g];
            }
        free(fill);
    }
    /* heavy terms first for balance */
    uint32_t *order = (uint32_t *)malloc((n_terms + 1) * sizeof(uint32_t));
    for (uint32_t t = 0; t < n_terms; t++) order[t] = t;
    for (uint32_t i = 1; i < n_terms; i++) {            /* insertion sort by p desc (n_terms ~ 1e3) */
        uint32_t v = order[i];
        int j = (int)i - 1;
        while (j >= 0 && term_p[order[j]] < term_p[v]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = v;
    }
    C.mem_start = mem_start; C.mem_g = mem_g; C.mem_s = mem_s; C.order = order;
    parallel_for((uint64_t)n_terms * nb, task_postings, &C, n_threads);   /* phase B */

    for (uint32_t t = 0; t < n_terms; t++)
        for (uint32_t bi = 0; bi < nb; bi++) S->term_len[t] += S->out[(size_t)t * nb + bi].n;
    for (size_t i = 0; i < (size_t)n_phrases * nb; i++) { free(C.plants[i].docs); free(C.plants[i].pos); }
    free(C.plants);
    free(mem_start); free(mem_g); free(mem_s); free(order);
    return S;
}

void sa_synth_dims(const sa_synth *S, uint64_t *doc_lo, uint64_t *doc_hi, uint64_t *term_len_out) {
    *doc_lo = S->doc_lo;
    *doc_hi = S->doc_hi;
    memcpy(term_len_out, S->term_len, S->n_terms * sizeof(uint64_t));
}

typedef struct { const sa_synth *S; uint64_t *offs; uint64_t *words_out; } copy_ctx_t;
static void task_copy(void *vc, uint64_t i) {
    copy_ctx_t *K = (copy_ctx_t *)vc;
    if (K->S->out[i].n) memcpy(K->words_out + K->offs[i], K->S->out[i].words, K->S->out[i].n * sizeof(uint64_t));
}

/* words_out: all terms' lists concatenated in term-id order (each list = its blocks in doc order) */
void sa_synth_copy(const sa_synth *S, uint64_t *words_out, float *doc_lens_out, int n_threads) {
    const uint32_t nb = S->blk_hi - S->blk_lo;
    const size_t n = (size_t)S->n_terms * nb;
    uint64_t *offs = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    uint64_t at = 0;
    for (size_t i = 0; i < n; i++) { offs[i] = at; at += S->out[i].n; }
    copy_ctx_t K = {S, offs, words_out};
    parallel_for(n, task_copy, &K, n_threads < 1 ? 1 : (n_threads > 16 ? 16 : n_threads));
    free(offs);
    memcpy(doc_lens_out, S->doc_lens, (S->doc_hi - S->doc_lo) * sizeof(float));
}

void sa_synth_free(sa_synth *S) {
    if (!S) return;
    const uint32_t nb = S->blk_hi - S->blk_lo;
    for (size_t i = 0; i < (size_t)S->n_terms * nb; i++) free(S->out[i].words);
    free(S->out);
    free(S->term_len);
    free(S->doc_lens);
    free(S);
}

typedef struct { uint64_t seed, field_key, n_docs; uint32_t n_blocks; double mu, sigma; float lo, hi; double *part; } sum_ctx_t;
static void task_sum(void *vc, uint64_t b) {
    sum_ctx_t *K = (sum_ctx_t *)vc;
    const uint64_t d0 = block_bound(K->n_docs, K->n_blocks, (uint32_t)b), d1 = block_bound(K->n_docs, K->n_blocks, (uint32_t)b + 1);
    float *tmp = (float *)malloc((d1 - d0 + 1) * sizeof(float));
    gen_doc_lens_block(K->seed, K->field_key, K->n_docs, K->n_blocks, (uint32_t)b, K->mu, K->sigma, K->lo, K->hi, tmp);
    double s = 0.0;
    for (uint64_t i = 0; i < d1 - d0; i++) s += tmp[i];
    K->part[b] = s;
    free(tmp);
}

/* exact float64 sum of ALL blocks' doc lengths (global average doc length, identical on every rank) */
double sa_synth_doc_len_sum(uint64_t seed, uint64_t field_key, uint64_t n_docs, uint32_t n_blocks,
                            double len_mu, double len_sigma, float len_lo, float len_hi, int n_threads) {
    double *part = (double *)calloc(n_blocks, sizeof(double));
    sum_ctx_t K = {seed, field_key, n_docs, n_blocks, len_mu, len_sigma, len_lo, len_hi, part};
    parallel_for(n_blocks, task_sum, &K, n_threads < 1 ? 1 : n_threads);
    double total = 0.0;
    for (uint32_t b = 0; b < n_blocks; b++) total += part[b];
    free(part);
    return total;
}

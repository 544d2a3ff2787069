// sa_phrase.cuh -- declarations for the phrase (slop == 0) path.
#pragma once
#include "sa_common.cuh"
#include "sa_term.cuh"

#define SA_PHRASE_THREADS 256
#define SA_PHRASE_MODE_LR 0      // left-to-right chain  (reference middle_out.py:96-122)
#define SA_PHRASE_MODE_RL 1      // right-to-left chain  (reference middle_out.py:125-151)
#define SA_PHRASE_MODE_MID 2     // middle-out: LR on [0,split), RL on [split,n), then and/min (:163-168)

// One phrase query against one shard.
struct PhraseQuery {
    u32 n_terms;
    u32 mode;
    u32 split;            // SA_PHRASE_MODE_MID only
    u32 same_guess;       // bit s set => step s is speculated to take the "same term" branch
    u64 off[SA_MAX_PHRASE_TERMS];   // word offset of each term's list (relative to `words`)
    u64 len[SA_MAX_PHRASE_TERMS];
    u64 dir_plus1[SA_MAX_PHRASE_TERMS];   // 1 + offset of the list's tile directory in d_tile_dir; 0 = none
    float idf;
    u32 pad;
};

struct PhraseStats {      // per query, accumulated over all chunks (global atomics)
    u32 n_inner[SA_MAX_PHRASE_TERMS];   // equal-header pairs seen at step s
    u32 n_diff[SA_MAX_PHRASE_TERMS];    // ... of which lhs word != rhs word
    u32 overflow;                        // 1: scratch arena exhausted; 2: conjunction regime gave up (dense candidates) -> search regime
    u32 n_match;                         // docs with a non-zero phrase count (M of SURVEY 8d's B_phrase)
    unsigned long long n_cont;           // continuation words written over all steps (sum of C_s)
};

// Optional dump of one CTA's final lists (per-op parity export; needs n_chunks == 1).
struct PhraseDump {
    u64 *cont;        // continuation words of the last step
    u64 *n_cont;
    u64 *docs;        // doc << 32 | count of the last step (BEFORE and/min with earlier steps)
    u64 *n_docs;
};

struct PhraseArgs {
    const u64 *words;           // lists live at words + off
    const float *doc_lens;
    u64 n_docs, doc_base;
    const PhraseQuery *queries;
    PhraseStats *stats;
    float *out;                 // [Q][out_stride], pre-zeroed
    u64 out_stride;
    u32 n_chunks;               // doc-range chunks per query (grid.x)
    u64 docs_per_chunk;
    u64 *arena;                 // scratch bump arena (u64 words)
    unsigned long long *arena_used;
    u64 arena_cap;
    Bm25Params bm25;
    int score;                  // 0: write phrase freqs, 1: BM25 (sparse)
    PhraseDump dump;
    // Every CTA writes the dense tiles of its doc range itself (zeros + matches) -- no separate
    // zero-fill pass -- and, when topk.k != 0, collects their top-k candidates on the way.
    // docs_per_chunk is a multiple of SA_TILE_DOCS.  topk.overflow / tile arrays are indexed by
    // row = topk_row0 + query.
    TopkCtx topk;
    u32 topk_row0;
    const u32 *tile_dir;        // word tile directories (sa_index::d_tile_dir) or NULL (filtered lists)
    const u32 *qsel;            // launch only these queries (indices into `queries`); NULL = all, in order
    u32 n_sel;
    // merge regime (phrase_staged_kernel, persistent CTAs)
    u32 stage_words;            // capacity of a CTA's staging buffer (dynamic shared memory), in words
    u64 *slabs;                 // per-CTA scratch: 6 * slab_cap words each
    u64 slab_cap;
    u32 *work_counter;
};

// Which queries of a batch run in which regime (device index lists into the PhraseQuery array)
struct PhraseSplit {
    const u32 *d_search;
    u32 n_search;
    const u32 *d_staged;
    u32 n_staged;
    u32 staged_chunks;
    u64 slab_cap;
};

int launch_phrase(sa_index *ix, const PhraseArgs &a, u32 n_queries);
void sa_phrase_plan(PhraseQuery &pq, const u32 *term_ids);
bool sa_phrase_guess_ok(PhraseQuery &pq, const PhraseStats &st);
u64 sa_phrase_arena_words(const PhraseQuery &pq, u32 n_chunks);
int sa_phrase_enqueue(sa_index *ix, const PhraseQuery *d_pqs, PhraseStats *d_stats, u32 Q,
                      float *dense_rows, u64 stride, u32 n_chunks, u64 *d_arena,
                      unsigned long long *d_arena_used, u64 arena_words, int score, const Bm25Params &p,
                      const TopkCtx *topk, u32 topk_row0, const PhraseSplit *split);
u32 sa_phrase_staged_chunks(const sa_index *ix);
u64 sa_phrase_slab_cap(const sa_index *ix, const u32 *term_ids, u32 n_terms);
u32 sa_phrase_chunks(const sa_index *ix, u32 wanted);
bool sa_phrase_is_staged(const PhraseQuery &pq, u64 n_docs);
u32 sa_phrase_stage_words();
int sa_phrase_run_sync(sa_index *ix, std::vector<PhraseQuery> &pqs, const u64 *d_words,
                       int score, const Bm25Params &p, u32 n_chunks_hint, PhraseDump dump, u64 staged_slab_cap);

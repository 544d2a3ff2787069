// sa_term.cuh -- kernel-side declarations shared by the term-path translation units.
#pragma once
#include "sa_common.cuh"

#define SA_TILE_DOCS 8192          // docs per CTA tile (32 KB of float32 scores)
#define SA_TERM_UNROLL 4           // 30-word windows loaded per warp before processing
#define SA_TERM_THREADS 256
#define SA_TOPK_MAX 32             // warp-level threshold estimation handles k <= 32

enum TermMode { TERM_MODE_TF = 0, TERM_MODE_SCORE = 1 };

// Per-query top-k collection state in HBM (see sa_topk.cu).  No global atomics: every
// (query, tile) CTA owns `slots` candidate slots.
struct TopkCtx {
    u32 *tile_cnt;     // [Q][n_tiles] candidates written by the tile's CTA (<= slots)
    u32 *tile_max;     // [Q][n_tiles] score bits of the tile's best candidate (0 = none)
    u64 *tile_cand;    // [Q][n_tiles][slots] key = score_bits << 32 | (0xFFFFFFFF - local_doc)
    u32 *overflow;     // [Q] set when some tile had more than `slots` candidates
    u32 n_tiles;
    u32 slots;
    u32 k;             // 0 => no top-k collection
};

struct TermBatchArgs {
    const u64 *words;
    const float *doc_lens;
    const float *norm;          // per-doc BM25 length norm (padded to a tile multiple), SCORE mode
    const u32 *tile_dir;        // tile directories (see sa_index::d_tile_dir)
    u64 n_docs;
    u64 doc_base;
    const TermQuery *queries;   // [Q]
    float *out;                 // [Q][out_stride]
    u64 out_stride;             // multiple of SA_TILE_DOCS
    Bm25Params bm25;            // idf field unused (per query)
    u64 min_payload, max_payload;
    int filter;                 // apply the payload_slice filter
    int mode;
    TopkCtx topk;
};

int launch_term_batch(sa_index *ix, const TermBatchArgs &a, u32 n_queries);
int sa_ensure_norm(sa_index *ix, float k1, float b, float avg_doc_len);
int launch_topk_select(sa_index *ix, const TopkCtx &t, u32 n_queries, u64 doc_base, u64 *d_out_keys,
                       const u32 *d_out_index);
u32 sa_topk_slots(u32 k);
int launch_dense_topk_tiles(sa_index *ix, const float *dense, u64 stride, u32 row0, u32 n_rows, const TopkCtx &t);
int launch_topk_merge(sa_index *ix, const u64 *d_in, u32 world, u32 n_queries, u32 k, u64 *d_out);
// batch plumbing shared by sa_index.cu / sa_comm.cu (callers hold ix->mu)
int sa_batch_upload_locked(sa_index *ix, const uint32_t *terms, const uint32_t *term_starts,
                           const float *idf, uint32_t n_queries, uint32_t slop,
                           float avg_doc_len, float k1, float b, uint32_t k);
int sa_batch_execute_locked(sa_index *ix);
int sa_batch_fix_overflow_locked(sa_index *ix, u32 *n_redone);
void sa_batch_dims(sa_index *ix, u32 *nq, u32 *k);
void sa_unpack_keys(const u64 *keys, u64 n, uint32_t *out_docs, float *out_scores);

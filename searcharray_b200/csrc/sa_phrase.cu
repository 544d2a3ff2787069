// sa_phrase.cu -- phrase / slop path (under construction) and per-op test exports.
#include "sa_term.cuh"

#define SA_TODO(name)                                \
    do {                                             \
        sa_set_error(name ": not implemented yet");  \
        return SA_ERR_ARG;                           \
    } while (0)

extern "C" int sa_index_set_rows(sa_index *, const uint64_t *, uint64_t) { SA_TODO("sa_index_set_rows"); }
extern "C" int sa_docfreq_rows(sa_index *, uint32_t, uint64_t *) { SA_TODO("sa_docfreq_rows"); }
extern "C" int sa_phrase_freqs(sa_index *, const uint32_t *, uint32_t, uint32_t, uint64_t, uint64_t, float *) { SA_TODO("sa_phrase_freqs"); }
extern "C" int sa_score_phrase(sa_index *, const uint32_t *, uint32_t, uint32_t, float, float, float, float,
                               uint64_t, uint64_t, float *) { SA_TODO("sa_score_phrase"); }
extern "C" int sa_op_popcount64_reduce(const uint64_t *, uint64_t, int, uint64_t *, float *, uint64_t *) { SA_TODO("sa_op_popcount64_reduce"); }
extern "C" int sa_op_bm25_score(float *, const float *, uint64_t, float, float, float, float, int) { SA_TODO("sa_op_bm25_score"); }
extern "C" int sa_op_bigram_freqs(const uint64_t *, uint64_t, const uint64_t *, uint64_t, int, int,
                                  uint64_t *, float *, uint64_t *, uint64_t *, uint64_t *) { SA_TODO("sa_op_bigram_freqs"); }

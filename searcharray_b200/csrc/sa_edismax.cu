// sa_edismax.cu -- multi-field query combination on the device.
//
// Replaces the numpy part of the reference's edismax (searcharray/solr.py:117-355): the per-(term,
// field) BM25 vectors, the per-field phrase vectors and the combined score vector stay in HBM; the
// host only parses the query, computes idf (numpy, like the reference) and reads back the result
// (dense, or just the top-k).
//
//   sa_multi_qf          solr.py:117-178.  One fused term-kernel launch per field writes the field's
//                        BM25 rows; edismax_combine_kernel folds them per doc with the reference's
//                        exact arithmetic: float32 `score * boost`, float64 running sum / maximum,
//                        term = max + (sum - max) * tie, mm on "terms scoring > 0", sum over terms
//                        in order (term-centric); all-float32 per-field sums (field-centric).
//   sa_multi_filter      the phrase phases run on arrays SLICED to qf > 0 (solr.py:326-330): the
//                        posting lists of the query terms are filtered by the match mask
//                        (sa_filter.cu) and their filtered doc frequencies returned (quirk iii).
//   sa_multi_phrases     every pf / pf2 / pf3 phrase of one field in one phrase-kernel launch on the
//                        filtered lists (BM25 applied in the kernel).
//   sa_multi_add_phase   solr.py:335-353: float32 sum of the phase's boosted vectors in list order,
//                        added to qf where qf != 0.
//   sa_multi_topk        exact top-k of the float64 vector: the high 32 bits of a positive double
//                        order like the double itself, so the float32 top-k machinery (sa_topk.cu)
//                        finds the k-th largest high word; every doc at or above it is then sorted
//                        exactly (score desc, doc asc).
#include <algorithm>
#include <cmath>
#include <functional>

#include "sa_phrase.cuh"
#include "sa_term.cuh"

int sa_filter_terms_mask(sa_index *ix, const uint32_t *term_ids, uint32_t n_terms, const unsigned char *d_mask,
                         u64 pay_lo, u64 pay_hi, bool use_payload, std::vector<u64> &offs, std::vector<u64> &lens,
                         std::vector<u64> *df_out);

#define ED_MAX_FIELDS 8
#define ED_MAX_ROWS 64
#define ED_TOPK_CAP 2048

struct sa_multi {
    std::vector<sa_index *> fields;
    int device = 0;
    u64 n_docs = 0, doc_base = 0, stride = 0;
    cudaStream_t stream = nullptr;
    double *d_qf = nullptr;              // [stride] combined scores (float32 values widened in field-centric mode)
    unsigned char *d_mask = nullptr;     // [stride] qf > 0 after the qf phase
    float *d_proxy = nullptr;            // [stride] high words of qf (top-k)
    unsigned long long *d_count = nullptr;
    u64 *d_pairs = nullptr;              // top-k candidates: score bits, doc
    bool f32_mode = false, has_qf = false;
    std::vector<std::vector<u64>> filt_offs, filt_lens;   // per field: last sa_multi_filter
    std::vector<u32> phrase_rows;        // per field: rows produced by the last sa_multi_phrases
    std::vector<u64> filt_bound;         // per field: words reserved for filtered lists (0 = not computed yet)
    DevBuf cand, meta, keys;
    std::mutex mu;
};

// All kernels of one multi call run on the multi's stream, including the ones the per-field
// helpers launch on `ix->stream`: the field streams are swapped for the duration of the call.
struct FieldGuard {
    sa_index *ix;
    cudaStream_t saved;
    std::unique_lock<std::mutex> lk;
    FieldGuard(sa_index *ix_, cudaStream_t s) : ix(ix_), saved(ix_->stream), lk(ix_->mu) {
        cudaStreamSynchronize(saved);
        ix->stream = s;
    }
    ~FieldGuard() {
        cudaStreamSynchronize(ix->stream);
        ix->stream = saved;
    }
};

struct CombineArgs {
    const float *rows[ED_MAX_FIELDS];    // field f: [n_terms[f]][stride]
    u32 n_terms[ED_MAX_FIELDS];
    float boost[ED_MAX_FIELDS];
    u32 has_boost[ED_MAX_FIELDS];
    u32 mm[ED_MAX_FIELDS];
    u32 n_fields;
    double tie;
    u64 n_docs, stride;
    double *qf;
    unsigned char *mask;
    unsigned long long *count;
};

// term-centric (solr.py:117-147)
__global__ void __launch_bounds__(256)
edismax_combine_terms_kernel(const CombineArgs a) {
    const u64 d = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (d < a.stride) {
        double total = 0.0;
        u32 matched = 0;
        if (d < a.n_docs) {
            const u32 T = a.n_terms[0];
            for (u32 t = 0; t < T; t++) {
                double run_sum = 0.0, run_max = 0.0;
                for (u32 f = 0; f < a.n_fields; f++) {
                    float s = a.rows[f][(u64)t * a.stride + d];
                    if (a.has_boost[f]) s = __fmul_rn(s, a.boost[f]);
                    run_sum = __dadd_rn(run_sum, (double)s);
                    run_max = fmax(run_max, (double)s);               // np.maximum (no NaNs on this path)
                }
                const double term = __dadd_rn(run_max, __dmul_rn(__dadd_rn(run_sum, -run_max), a.tie));
                if (term > 0.0) matched++;
                total = t == 0 ? term : __dadd_rn(total, term);
            }
            if (matched < a.mm[0]) total = 0.0;
        }
        a.qf[d] = total;
        hit = total > 0.0;
        a.mask[d] = hit ? 1 : 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(a.count, (unsigned long long)__popc(m));
}

// field-centric (solr.py:150-178): everything float32
__global__ void __launch_bounds__(256)
edismax_combine_fields_kernel(const CombineArgs a) {
    const u64 d = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (d < a.stride) {
        float result = 0.0f;
        if (d < a.n_docs) {
            float summed = 0.0f, best = 0.0f;
            for (u32 f = 0; f < a.n_fields; f++) {
                float tot = 0.0f;
                u32 matched = 0;
                for (u32 t = 0; t < a.n_terms[f]; t++) {
                    const float s = a.rows[f][(u64)t * a.stride + d];
                    if (s > 0.0f) matched++;
                    tot = t == 0 ? s : __fadd_rn(tot, s);
                }
                if (matched < a.mm[f]) tot = 0.0f;
                if (a.has_boost[f]) tot = __fmul_rn(tot, a.boost[f]);
                summed = f == 0 ? tot : __fadd_rn(summed, tot);
                best = f == 0 ? tot : fmaxf(best, tot);
            }
            // qf + (summed - qf) * tie with a Python-float tie: numpy keeps float32
            result = __fadd_rn(best, __fmul_rn(__fadd_rn(summed, -best), (float)a.tie));
        }
        a.qf[d] = (double)result;
        hit = result > 0.0f;
        a.mask[d] = hit ? 1 : 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(a.count, (unsigned long long)__popc(m));
}

struct PhaseArgs {
    const float *rows[ED_MAX_ROWS];
    float boost[ED_MAX_ROWS];
    u32 has_boost[ED_MAX_ROWS];
    u32 n;
    u64 n_docs;
    double *qf;
    int f32_mode;
};

// qf[where qf != 0] += float32 sum of the phase's vectors, in order (solr.py:335-353)
__global__ void __launch_bounds__(256)
edismax_add_phase_kernel(const PhaseArgs a) {
    const u64 d = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= a.n_docs) return;
    const double cur = a.qf[d];
    if (cur == 0.0) return;
    float acc = 0.0f;
    for (u32 i = 0; i < a.n; i++) {
        float s = a.rows[i][d];
        if (a.has_boost[i]) s = __fmul_rn(s, a.boost[i]);
        acc = i == 0 ? s : __fadd_rn(acc, s);
    }
    if (a.f32_mode) a.qf[d] = (double)__fadd_rn((float)cur, acc);
    else a.qf[d] = __dadd_rn(cur, (double)acc);
}

__global__ void __launch_bounds__(256)
edismax_proxy_kernel(const double *__restrict__ qf, float *__restrict__ proxy, u64 stride) {
    const u64 d = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= stride) return;
    const double v = qf[d];
    proxy[d] = v > 0.0 ? __uint_as_float((u32)((u64)__double_as_longlong(v) >> 32)) : 0.0f;
}

// every doc whose high word is >= the k-th best high word (keys[k-1] of the proxy top-k; 0 = fewer than k matches)
__global__ void __launch_bounds__(256)
edismax_gather_kernel(const double *__restrict__ qf, u64 n_docs, const u64 *__restrict__ proxy_keys, u32 k,
                      u64 *__restrict__ pairs, unsigned long long *__restrict__ count, u32 cap) {
    const u64 d = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const double v = qf[d];
    if (!(v > 0.0)) return;
    const u32 kth_hi = (u32)(proxy_keys[k - 1] >> 32);
    const u64 bits = (u64)__double_as_longlong(v);
    if ((u32)(bits >> 32) < kth_hi) return;
    const unsigned long long slot = atomicAdd(count, 1ull);
    if (slot < cap) { pairs[2 * slot] = bits; pairs[2 * slot + 1] = d; }
}

// one CTA: exact order of <= ED_TOPK_CAP candidates by (score desc, doc asc), first k out
__global__ void __launch_bounds__(1024)
edismax_sort_kernel(const u64 *__restrict__ pairs, const unsigned long long *__restrict__ count, u32 k, u64 doc_base,
                    double *__restrict__ out_scores, u32 *__restrict__ out_docs) {
    __shared__ u64 s_key[ED_TOPK_CAP];
    __shared__ u32 s_doc[ED_TOPK_CAP];
    const u32 n = (u32)min((unsigned long long)ED_TOPK_CAP, *count);
    for (u32 i = threadIdx.x; i < ED_TOPK_CAP; i += blockDim.x) {
        s_key[i] = i < n ? pairs[2 * i] : 0ull;
        s_doc[i] = i < n ? (u32)pairs[2 * i + 1] : 0xFFFFFFFFu;
    }
    __syncthreads();
    for (u32 size = 2; size <= ED_TOPK_CAP; size <<= 1) {
        for (u32 strideI = size >> 1; strideI > 0; strideI >>= 1) {
            for (u32 i = threadIdx.x; i < ED_TOPK_CAP / 2; i += blockDim.x) {
                const u32 lo = 2 * i - (i & (strideI - 1));
                const u32 hi = lo + strideI;
                const bool desc_block = ((lo & size) == 0);
                // "a before b": larger score first, then smaller doc
                const bool a_first = s_key[lo] > s_key[hi] || (s_key[lo] == s_key[hi] && s_doc[lo] < s_doc[hi]);
                if (a_first != desc_block) {
                    const u64 tk = s_key[lo]; s_key[lo] = s_key[hi]; s_key[hi] = tk;
                    const u32 td = s_doc[lo]; s_doc[lo] = s_doc[hi]; s_doc[hi] = td;
                }
            }
            __syncthreads();
        }
    }
    for (u32 i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = i < n && s_key[i] != 0ull;
        out_scores[i] = ok ? __longlong_as_double((long long)s_key[i]) : 0.0;
        out_docs[i] = ok ? (u32)(s_doc[i] + doc_base) : SA_NO_DOC;
    }
}

// ---- exact fallback of the top-k when more than ED_TOPK_CAP docs share the leading 32 bits of the k-th score
// (BM25 scores are a function of (tf, doc length): on a field of uniform short docs thousands of docs tie exactly).
// hist[b] = docs with score > 0 whose bits match `prefix` above `shift + 8` and whose next byte is b
__global__ void __launch_bounds__(256)
edismax_hist_kernel(const double *__restrict__ qf, u64 n_docs, u64 prefix, int shift, unsigned long long *__restrict__ hist) {
    __shared__ unsigned int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    for (u64 d = (u64)blockIdx.x * blockDim.x + threadIdx.x; d < n_docs; d += (u64)gridDim.x * blockDim.x) {
        const double v = qf[d];
        if (!(v > 0.0)) continue;
        const u64 bits = (u64)__double_as_longlong(v);
        if (shift < 56 && (bits >> (shift + 8)) != (prefix >> (shift + 8))) continue;
        atomicAdd(&s_h[(bits >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_h[threadIdx.x]);
}

// docs scoring strictly more than `kth_bits` (fewer than k of them) -> pairs; ties per 1024-doc block -> tie_cnt
__global__ void __launch_bounds__(256)
edismax_above_kernel(const double *__restrict__ qf, u64 n_docs, u64 kth_bits, u64 *__restrict__ pairs,
                     unsigned long long *__restrict__ count, u32 cap, u32 *__restrict__ tie_cnt) {
    __shared__ unsigned int s_t;
    if (threadIdx.x == 0) s_t = 0;
    __syncthreads();
    const u64 d0 = (u64)blockIdx.x * 1024;
    u32 mine = 0;
    for (u32 i = threadIdx.x; i < 1024; i += 256) {
        const u64 d = d0 + i;
        if (d >= n_docs) break;
        const double v = qf[d];
        if (!(v > 0.0)) continue;
        const u64 bits = (u64)__double_as_longlong(v);
        if (bits > kth_bits) {
            const unsigned long long slot = atomicAdd(count, 1ull);
            if (slot < cap) { pairs[2 * slot] = bits; pairs[2 * slot + 1] = d; }
        } else if (bits == kth_bits) {
            mine++;
        }
    }
    if (mine) atomicAdd(&s_t, mine);
    __syncthreads();
    if (threadIdx.x == 0) tie_cnt[blockIdx.x] = s_t;
}

// ------------------------------------------------------------------------------ host
extern "C" int sa_multi_create(sa_index *const *fields, uint32_t n_fields, sa_multi **out) {
    SA_CHECK(fields && out && n_fields >= 1 && n_fields <= ED_MAX_FIELDS, "1..%d fields", ED_MAX_FIELDS);
    for (u32 f = 0; f < n_fields; f++) {
        SA_CHECK(fields[f], "field %u is NULL", f);
        SA_CHECK(fields[f]->device == fields[0]->device && fields[f]->n_docs == fields[0]->n_docs &&
                 fields[f]->doc_base == fields[0]->doc_base, "fields must share device, doc range and size");
    }
    sa_multi *m = new sa_multi();
    m->fields.assign(fields, fields + n_fields);
    m->device = fields[0]->device;
    m->n_docs = fields[0]->n_docs;
    m->doc_base = fields[0]->doc_base;
    m->stride = (m->n_docs + SA_TILE_DOCS - 1) / SA_TILE_DOCS * SA_TILE_DOCS;
    m->filt_offs.resize(n_fields);
    m->filt_lens.resize(n_fields);
    m->phrase_rows.assign(n_fields, 0);
    m->filt_bound.assign(n_fields, 0);
    cudaSetDevice(m->device);
    const u64 s = std::max<u64>(m->stride, SA_TILE_DOCS);
    bool ok = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&m->d_qf, s * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&m->d_mask, s) == cudaSuccess &&
              cudaMalloc(&m->d_proxy, s * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&m->d_count, 64) == cudaSuccess &&
              cudaMalloc(&m->d_pairs, 2ull * ED_TOPK_CAP * sizeof(u64)) == cudaSuccess;
    if (!ok) {
        sa_set_error("sa_multi_create: %s", cudaGetErrorString(cudaGetLastError()));
        sa_multi_destroy(m);
        return SA_ERR_NOMEM;
    }
    *out = m;
    return SA_OK;
}

extern "C" int sa_multi_destroy(sa_multi *m) {
    if (!m) return SA_OK;
    cudaSetDevice(m->device);
    if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
    cudaFree(m->d_qf);
    cudaFree(m->d_mask);
    cudaFree(m->d_proxy);
    cudaFree(m->d_count);
    cudaFree(m->d_pairs);
    m->cand.release();
    m->meta.release();
    m->keys.release();
    delete m;
    return SA_OK;
}

extern "C" int sa_multi_qf(sa_multi *m, int field_centric, const uint32_t *n_terms, const uint32_t *term_ids,
                           const float *idf, const float *boost, const uint32_t *has_boost,
                           const float *avg_doc_len, const float *k1, const float *b, const uint32_t *mm,
                           double tie, uint64_t *n_matches) {
    SA_CHECK(m && n_terms && boost && has_boost && avg_doc_len && k1 && b && mm && n_matches, "NULL argument");
    std::lock_guard<std::mutex> g(m->mu);
    SA_CUDA(cudaSetDevice(m->device));
    const u32 F = (u32)m->fields.size();
    CombineArgs a;
    memset(&a, 0, sizeof(a));
    u32 at = 0;
    for (u32 f = 0; f < F; f++) {
        const u32 T = n_terms[f];
        SA_CHECK(T <= SA_MAX_PHRASE_TERMS, "too many query terms");
        SA_CHECK(field_centric || T == n_terms[0], "term-centric needs the same number of terms per field");
        SA_CHECK(T == 0 || (term_ids && idf), "NULL argument");
        sa_index *ix = m->fields[f];
        FieldGuard fg(ix, m->stream);
        int rc;
        if ((rc = ix->dense.reserve(std::max<u64>(T, 1) * m->stride * sizeof(float)))) return rc;
        a.rows[f] = ix->dense.as<float>();
        a.n_terms[f] = T;
        a.boost[f] = boost[f];
        a.has_boost[f] = has_boost[f];
        a.mm[f] = mm[f];
        if (T && (avg_doc_len[f] == 0.0f || m->n_docs == 0)) {         // similarity.py:31-32: zeros
            SA_CUDA(cudaMemsetAsync(ix->dense.p, 0, (size_t)T * m->stride * sizeof(float), m->stream));
        } else if (T) {
            std::vector<TermQuery> tqs(T);
            Bm25Params p = sa_make_bm25(ix, 1.0f, avg_doc_len[f], k1[f], b[f]);
            for (u32 t = 0; t < T; t++) {
                const u32 id = term_ids[at + t];
                SA_CHECK(id == SA_NO_TERM || id < ix->n_terms, "term id %u out of range", id);
                tqs[t] = sa_make_term_query(ix, id, idf[at + t]);
                if (!sa_make_bm25(ix, idf[at + t], avg_doc_len[f], k1[f], b[f]).sparse_ok) p.sparse_ok = 0;
            }
            if ((rc = ix->queries.reserve(T * sizeof(TermQuery)))) return rc;
            SA_CUDA(cudaMemcpyAsync(ix->queries.p, tqs.data(), T * sizeof(TermQuery), cudaMemcpyHostToDevice, m->stream));
            TopkCtx none;
            memset(&none, 0, sizeof(none));
            TermBatchArgs ta;
            memset(&ta, 0, sizeof(ta));
            ta.words = ix->d_words;
            ta.doc_lens = ix->d_doc_lens;
            ta.n_docs = ix->n_docs;
            ta.doc_base = ix->doc_base;
            ta.queries = ix->queries.as<TermQuery>();
            ta.out = ix->dense.as<float>();
            ta.out_stride = m->stride;
            ta.bm25 = p;
            ta.min_payload = 0;
            ta.max_payload = SA_ALL_BITS;
            ta.mode = TERM_MODE_SCORE;
            ta.topk = none;
            if ((rc = launch_term_batch(ix, ta, T))) return rc;
            SA_CUDA(cudaStreamSynchronize(m->stream));                  // tqs leaves scope
        }
        at += T;
    }
    a.n_fields = F;
    a.tie = tie;
    a.n_docs = m->n_docs;
    a.stride = m->stride;
    a.qf = m->d_qf;
    a.mask = m->d_mask;
    a.count = m->d_count;
    SA_CUDA(cudaMemsetAsync(m->d_count, 0, sizeof(unsigned long long), m->stream));
    const unsigned blocks = (unsigned)((std::max<u64>(m->stride, 1) + 255) / 256);
    if (field_centric) edismax_combine_fields_kernel<<<blocks, 256, 0, m->stream>>>(a);
    else edismax_combine_terms_kernel<<<blocks, 256, 0, m->stream>>>(a);
    SA_CUDA(cudaGetLastError());
    unsigned long long cnt = 0;
    SA_CUDA(cudaMemcpyAsync(&cnt, m->d_count, sizeof(cnt), cudaMemcpyDeviceToHost, m->stream));
    SA_CUDA(cudaStreamSynchronize(m->stream));
    *n_matches = cnt;
    m->f32_mode = field_centric != 0;
    m->has_qf = true;
    for (u32 f = 0; f < F; f++) { m->filt_offs[f].clear(); m->filt_lens[f].clear(); m->phrase_rows[f] = 0; }
    return SA_OK;
}

extern "C" int sa_multi_filter(sa_multi *m, uint32_t field, const uint32_t *term_ids, uint32_t n_terms,
                               uint64_t *df_out) {
    SA_CHECK(m && term_ids && df_out && field < m->fields.size() && n_terms >= 1, "bad argument");
    SA_CHECK(m->has_qf, "sa_multi_qf has not run");
    std::lock_guard<std::mutex> g(m->mu);
    SA_CUDA(cudaSetDevice(m->device));
    sa_index *ix = m->fields[field];
    for (u32 t = 0; t < n_terms; t++)
        SA_CHECK(term_ids[t] == SA_NO_TERM || term_ids[t] < ix->n_terms, "term id %u out of range", term_ids[t]);
    FieldGuard fg(ix, m->stream);
    std::vector<u64> dfs;
    if (m->filt_bound[field] == 0) {
        // one allocation for any query: room for filtered copies of the longest lists a query could name
        std::vector<u64> lens(ix->h_len);
        const size_t top = std::min<size_t>(lens.size(), SA_MAX_PHRASE_TERMS);
        std::partial_sort(lens.begin(), lens.begin() + top, lens.end(), std::greater<u64>());
        u64 bound = 64;
        for (size_t i = 0; i < top; i++) bound += lens[i] + 2;
        m->filt_bound[field] = bound;
    }
    {
        int rc0 = ix->filt.reserve(m->filt_bound[field] * sizeof(u64));
        if (rc0) return rc0;
    }
    int rc = sa_filter_terms_mask(ix, term_ids, n_terms, m->d_mask, 0, SA_ALL_BITS, false,
                                  m->filt_offs[field], m->filt_lens[field], &dfs);
    if (rc) return rc;
    for (u32 t = 0; t < n_terms; t++) df_out[t] = dfs[t];
    return SA_OK;
}

extern "C" int sa_multi_phrases(sa_multi *m, uint32_t field, uint32_t n_phrases, const uint32_t *phrase_starts,
                                const uint32_t *term_slots, const uint32_t *term_ids, const float *idf,
                                float avg_doc_len, float k1, float b) {
    SA_CHECK(m && field < m->fields.size() && n_phrases >= 1 && n_phrases <= ED_MAX_ROWS, "bad argument");
    SA_CHECK(phrase_starts && term_slots && term_ids && idf, "NULL argument");
    std::lock_guard<std::mutex> g(m->mu);
    SA_CUDA(cudaSetDevice(m->device));
    sa_index *ix = m->fields[field];
    const std::vector<u64> &offs = m->filt_offs[field], &lens = m->filt_lens[field];
    SA_CHECK(!offs.empty(), "sa_multi_filter has not run for this field");
    FieldGuard fg(ix, m->stream);
    int rc;
    m->phrase_rows[field] = n_phrases;
    if (avg_doc_len == 0.0f || m->n_docs == 0) {
        if ((rc = ix->dense.reserve((size_t)n_phrases * m->stride * sizeof(float)))) return rc;
        SA_CUDA(cudaMemsetAsync(ix->dense.p, 0, (size_t)n_phrases * m->stride * sizeof(float), m->stream));
        return SA_OK;
    }
    std::vector<PhraseQuery> pqs(n_phrases);
    Bm25Params p = sa_make_bm25(ix, 1.0f, avg_doc_len, k1, b);
    for (u32 i = 0; i < n_phrases; i++) {
        PhraseQuery &pq = pqs[i];
        memset(&pq, 0, sizeof(pq));
        const u32 s0 = phrase_starts[i], nt = phrase_starts[i + 1] - s0;
        SA_CHECK(nt >= 2 && nt <= SA_MAX_PHRASE_TERMS, "phrase %u: 2..%d terms", i, SA_MAX_PHRASE_TERMS);
        pq.n_terms = nt;
        pq.idf = idf[i];
        SA_CHECK(sa_make_bm25(ix, idf[i], avg_doc_len, k1, b).sparse_ok,
                 "edismax phrase phases need ordinary BM25 parameters (k1 > 0, 0 <= b < 1, finite idf >= 0)");
        bool missing = false;
        for (u32 j = 0; j < nt; j++) {
            const u32 slot = term_slots[s0 + j];
            SA_CHECK(slot < offs.size(), "term slot out of range");
            if (term_ids[s0 + j] == SA_NO_TERM || lens[slot] == 0) missing = true;
            pq.off[j] = offs[slot];
            pq.len[j] = lens[slot];
        }
        if (missing) for (u32 j = 0; j < nt; j++) pq.len[j] = 0;       // unknown term -> zeros (postings.py:705-708)
        sa_phrase_plan(pq, term_ids + s0);
    }
    PhraseDump nodump;
    memset(&nodump, 0, sizeof(nodump));
    return sa_phrase_run_sync(ix, pqs, ix->filt.as<u64>(), 1, p, 0, nodump, 0);
}

extern "C" int sa_multi_add_phase(sa_multi *m, uint32_t n_entries, const uint32_t *entry_field,
                                  const uint32_t *entry_row, const float *entry_boost, const uint32_t *entry_has_boost) {
    SA_CHECK(m && m->has_qf, "sa_multi_qf has not run");
    if (n_entries == 0) return SA_OK;
    SA_CHECK(entry_field && entry_row && entry_boost && entry_has_boost && n_entries <= ED_MAX_ROWS, "bad argument");
    std::lock_guard<std::mutex> g(m->mu);
    SA_CUDA(cudaSetDevice(m->device));
    PhaseArgs a;
    memset(&a, 0, sizeof(a));
    for (u32 i = 0; i < n_entries; i++) {
        SA_CHECK(entry_field[i] < m->fields.size() && entry_row[i] < m->phrase_rows[entry_field[i]], "entry %u out of range", i);
        a.rows[i] = m->fields[entry_field[i]]->dense.as<float>() + (u64)entry_row[i] * m->stride;
        a.boost[i] = entry_boost[i];
        a.has_boost[i] = entry_has_boost[i];
    }
    a.n = n_entries;
    a.n_docs = m->n_docs;
    a.qf = m->d_qf;
    a.f32_mode = m->f32_mode ? 1 : 0;
    if (m->n_docs) {
        edismax_add_phase_kernel<<<(unsigned)((m->n_docs + 255) / 256), 256, 0, m->stream>>>(a);
        SA_CUDA(cudaGetLastError());
    }
    SA_CUDA(cudaStreamSynchronize(m->stream));
    return SA_OK;
}

extern "C" int sa_multi_download(sa_multi *m, void *out, int as_float32) {
    SA_CHECK(m && out && m->has_qf, "nothing to download");
    std::lock_guard<std::mutex> g(m->mu);
    SA_CUDA(cudaSetDevice(m->device));
    if (m->n_docs == 0) return SA_OK;
    if (!as_float32) {
        SA_CUDA(cudaMemcpyAsync(out, m->d_qf, m->n_docs * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        SA_CUDA(cudaStreamSynchronize(m->stream));
        return SA_OK;
    }
    std::vector<double> tmp(m->n_docs);
    SA_CUDA(cudaMemcpyAsync(tmp.data(), m->d_qf, m->n_docs * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    SA_CUDA(cudaStreamSynchronize(m->stream));
    float *o = (float *)out;
    for (u64 i = 0; i < m->n_docs; i++) o[i] = (float)tmp[i];           // exact: the values are float32
    return SA_OK;
}

extern "C" int sa_multi_is_float32(sa_multi *m, int *out) {
    SA_CHECK(m && out, "NULL argument");
    *out = m->f32_mode ? 1 : 0;
    return SA_OK;
}

extern "C" int sa_multi_topk(sa_multi *m, uint32_t k, uint32_t *out_docs, double *out_scores) {
    SA_CHECK(m && out_docs && out_scores && m->has_qf, "bad argument");
    SA_CHECK(k >= 1 && k <= SA_TOPK_MAX, "k must be in [1, %d]", SA_TOPK_MAX);
    std::lock_guard<std::mutex> g(m->mu);
    SA_CUDA(cudaSetDevice(m->device));
    for (u32 i = 0; i < k; i++) { out_docs[i] = SA_NO_DOC; out_scores[i] = 0.0; }
    if (m->n_docs == 0) return SA_OK;
    sa_index *ix = m->fields[0];
    FieldGuard fg(ix, m->stream);
    const u32 T = (u32)(m->stride / SA_TILE_DOCS);
    int rc;
    unsigned blocks = (unsigned)((m->stride + 255) / 256);
    edismax_proxy_kernel<<<blocks, 256, 0, m->stream>>>(m->d_qf, m->d_proxy, m->stride);
    SA_CUDA(cudaGetLastError());
    u32 slots = sa_topk_slots(k);
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = m->cand.reserve((size_t)T * ((size_t)slots * sizeof(u64) + 2 * sizeof(u32)) + 64))) return rc;
        if ((rc = m->meta.reserve(256))) return rc;
        if ((rc = m->keys.reserve((size_t)k * (sizeof(u64) + sizeof(double) + sizeof(u32)) + 64))) return rc;
        SA_CUDA(cudaMemsetAsync(m->meta.p, 0, 256, m->stream));
        TopkCtx t;
        t.tile_cand = m->cand.as<u64>();
        t.tile_cnt = (u32 *)(t.tile_cand + (u64)T * slots);
        t.tile_max = t.tile_cnt + T;
        t.overflow = m->meta.as<u32>();
        t.n_tiles = T;
        t.slots = slots;
        t.k = k;
        if ((rc = launch_dense_topk_tiles(ix, m->d_proxy, m->stride, 0, 1, t, nullptr))) return rc;
        if ((rc = launch_topk_select(ix, t, 1, 0, m->keys.as<u64>(), nullptr))) return rc;
        u32 ovf = 0;
        SA_CUDA(cudaMemcpyAsync(&ovf, m->meta.p, sizeof(u32), cudaMemcpyDeviceToHost, m->stream));
        SA_CUDA(cudaStreamSynchronize(m->stream));
        if (!ovf) break;
        slots = SA_TILE_DOCS;                                           // cannot overflow
    }
    SA_CUDA(cudaMemsetAsync(m->d_count, 0, sizeof(unsigned long long), m->stream));
    edismax_gather_kernel<<<(unsigned)((m->n_docs + 255) / 256), 256, 0, m->stream>>>(m->d_qf, m->n_docs, m->keys.as<u64>(), k,
                                                                                   m->d_pairs, m->d_count, ED_TOPK_CAP);
    SA_CUDA(cudaGetLastError());
    double *d_scores = (double *)(m->keys.as<u64>() + k);
    u32 *d_docs = (u32 *)(d_scores + k);
    edismax_sort_kernel<<<1, 1024, 0, m->stream>>>(m->d_pairs, m->d_count, k, m->doc_base, d_scores, d_docs);
    SA_CUDA(cudaGetLastError());
    unsigned long long cnt = 0;
    SA_CUDA(cudaMemcpyAsync(&cnt, m->d_count, sizeof(cnt), cudaMemcpyDeviceToHost, m->stream));
    SA_CUDA(cudaMemcpyAsync(out_scores, d_scores, k * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    SA_CUDA(cudaMemcpyAsync(out_docs, d_docs, k * sizeof(u32), cudaMemcpyDeviceToHost, m->stream));
    SA_CUDA(cudaStreamSynchronize(m->stream));
    if (cnt <= ED_TOPK_CAP) return SA_OK;

    // ---- more than ED_TOPK_CAP docs share the k-th score's leading bits: exact selection on the full 64 bits
    // 1. the k-th largest score (with multiplicity): 8-pass byte-wise radix select over all docs
    unsigned long long *d_hist = (unsigned long long *)m->d_pairs;             // 256 counters (the pairs buffer is free)
    unsigned long long h_hist[256];
    u64 prefix = 0, need = k;
    const unsigned hb = (unsigned)std::min<u64>(1024, (m->n_docs + 255) / 256);
    for (int shift = 56; shift >= 0; shift -= 8) {
        SA_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(h_hist), m->stream));
        edismax_hist_kernel<<<hb, 256, 0, m->stream>>>(m->d_qf, m->n_docs, prefix, shift, d_hist);
        SA_CUDA(cudaGetLastError());
        SA_CUDA(cudaMemcpyAsync(h_hist, d_hist, sizeof(h_hist), cudaMemcpyDeviceToHost, m->stream));
        SA_CUDA(cudaStreamSynchronize(m->stream));
        int b = 255;
        for (; b > 0; b--) {
            if (h_hist[b] >= need) break;
            need -= h_hist[b];
        }
        prefix |= (u64)b << shift;
    }
    const u64 kth_bits = prefix;                     // `need` docs with exactly this score belong to the top k
    // 2. docs above it (fewer than k) + ties per 1024-doc block
    const u32 n_blocks = (u32)((m->n_docs + 1023) / 1024);
    if ((rc = m->cand.reserve((size_t)n_blocks * sizeof(u32) + 64))) return rc;
    std::vector<u32> tie_cnt(n_blocks);
    std::vector<u64> above(2 * (size_t)k);
    SA_CUDA(cudaMemsetAsync(m->d_count, 0, sizeof(unsigned long long), m->stream));
    edismax_above_kernel<<<n_blocks, 256, 0, m->stream>>>(m->d_qf, m->n_docs, kth_bits, m->d_pairs, m->d_count, k, m->cand.as<u32>());
    SA_CUDA(cudaGetLastError());
    SA_CUDA(cudaMemcpyAsync(&cnt, m->d_count, sizeof(cnt), cudaMemcpyDeviceToHost, m->stream));
    SA_CUDA(cudaMemcpyAsync(tie_cnt.data(), m->cand.p, (size_t)n_blocks * sizeof(u32), cudaMemcpyDeviceToHost, m->stream));
    SA_CUDA(cudaStreamSynchronize(m->stream));
    SA_CHECK(cnt < k, "top-k fallback: inconsistent selection");
    if (cnt) SA_CUDA(cudaMemcpy(above.data(), m->d_pairs, 2 * (size_t)cnt * sizeof(u64), cudaMemcpyDeviceToHost));
    std::vector<std::pair<u64, u64>> best;           // (score bits, doc)
    for (u64 i = 0; i < cnt; i++) best.push_back({above[2 * i], above[2 * i + 1]});
    std::sort(best.begin(), best.end(), [](const std::pair<u64, u64> &x, const std::pair<u64, u64> &y) {
        return x.first > y.first || (x.first == y.first && x.second < y.second);
    });
    // 3. the `need` tied docs with the smallest ids: walk the blocks in order, read only the blocks that hold them
    u64 want_ties = std::min<u64>(need, (u64)k - cnt);
    std::vector<double> blk(1024);
    for (u32 bI = 0; bI < n_blocks && want_ties; bI++) {
        if (!tie_cnt[bI]) continue;
        const u64 d0 = (u64)bI * 1024, nb = std::min<u64>(1024, m->n_docs - d0);
        SA_CUDA(cudaMemcpy(blk.data(), m->d_qf + d0, nb * sizeof(double), cudaMemcpyDeviceToHost));
        for (u64 i = 0; i < nb && want_ties; i++) {
            u64 bits;
            memcpy(&bits, &blk[i], 8);
            if (blk[i] > 0.0 && bits == kth_bits) { best.push_back({bits, d0 + i}); want_ties--; }
        }
    }
    for (u32 i = 0; i < k; i++) {
        if (i < best.size()) {
            memcpy(&out_scores[i], &best[i].first, 8);
            out_docs[i] = (u32)(best[i].second + m->doc_base);
        } else {
            out_scores[i] = 0.0;
            out_docs[i] = SA_NO_DOC;
        }
    }
    return SA_OK;
}

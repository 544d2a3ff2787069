// sa_filter.cu -- sliced arrays and position filters.
//
// Replaces RoaringishEncoder.slice + payload_slice + FilteredPosns (reference
// searcharray/roaringish/roaringish.py:245-282, roaringish_ops.pyx:46-68, phrase/middle_out.py:291-317):
// the reference materialises, per term, the sub-list of posting words whose doc is among the slice's
// rows and whose block passes the min/max_posn test, then runs the ordinary algorithms on those
// lists.  Same here: an order-preserving compaction kernel writes the filtered lists into scratch,
// the term / phrase / span kernels run on them unchanged, and a gather kernel picks the slice's rows
// out of the dense result (`phrase_freqs[self.term_mat.rows]`, postings.py:702-704).
#include <algorithm>

#include "sa_phrase.cuh"
#include "sa_term.cuh"

struct FilterJob { u64 src_off, src_len, dst_off, chunk_off; };   // chunk_off: first entry in the chunk-count table

#define FILT_THREADS 256
#define FILT_ITEMS 8                                  // words per thread per chunk
#define FILT_CHUNK (FILT_THREADS * FILT_ITEMS)

__device__ __forceinline__ bool filter_keep(u64 w, const unsigned char *__restrict__ row_mask, u64 doc_base, u64 n_docs,
                                            u64 pay_lo, u64 pay_hi, int use_payload) {
    bool keep = true;
    if (row_mask) {
        const u64 d = (w >> SA_KEY_SHIFT) - doc_base;
        keep = d < n_docs && row_mask[d];
    }
    if (keep && use_payload) {
        const u64 v = w & SA_MSB_MASK;                  // UNSHIFTED compare, reference roaringish_ops.pyx:55
        keep = v >= pay_lo && v <= pay_hi;
    }
    return keep;
}

// Order-preserving compaction in three steps: per-chunk keep counts, a scan of the chunk counts of
// every list, then the write pass (the keep test is cheap, so it is simply evaluated twice).
template <bool WRITE>
__global__ void __launch_bounds__(FILT_THREADS)
filter_lists_kernel(const u64 *__restrict__ words, const FilterJob *__restrict__ jobs, u64 *__restrict__ dst,
                    u32 *__restrict__ chunk_counts, const unsigned char *__restrict__ row_mask, u64 doc_base, u64 n_docs,
                    u64 pay_lo, u64 pay_hi, int use_payload) {
    __shared__ u32 s_warp[FILT_THREADS / 32];
    const FilterJob job = jobs[blockIdx.y];
    const u64 base = (u64)blockIdx.x * FILT_CHUNK;
    if (base >= job.src_len) return;
    const u64 *__restrict__ src = words + job.src_off;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // thread t owns words base + t*ITEMS .. +ITEMS-1 (contiguous: the output order is the thread order)
    u64 w[FILT_ITEMS];
    u32 keep_bits = 0;
#pragma unroll
    for (int j = 0; j < FILT_ITEMS; j++) {
        const u64 i = base + (u64)tid * FILT_ITEMS + j;
        w[j] = 0;
        if (i < job.src_len) {
            w[j] = src[i];
            if (filter_keep(w[j], row_mask, doc_base, n_docs, pay_lo, pay_hi, use_payload)) keep_bits |= 1u << j;
        }
    }
    const u32 cnt = __popc(keep_bits);
    u32 incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    u32 off = incl - cnt, tot = 0;
#pragma unroll
    for (int wI = 0; wI < FILT_THREADS / 32; wI++) { u32 c = s_warp[wI]; if (wI < (int)warp) off += c; tot += c; }
    if (!WRITE) {
        if (tid == 0) chunk_counts[job.chunk_off + blockIdx.x] = tot;
        return;
    }
    u64 *__restrict__ out = dst + job.dst_off + chunk_counts[job.chunk_off + blockIdx.x] + off;   // scanned: exclusive offset
#pragma unroll
    for (int j = 0; j < FILT_ITEMS; j++)
        if ((keep_bits >> j) & 1u) *out++ = w[j];
}

// one CTA per list: exclusive scan of its chunk counts (in place); totals[job] = kept words
__global__ void __launch_bounds__(FILT_THREADS)
filter_scan_kernel(const FilterJob *__restrict__ jobs, u32 *__restrict__ chunk_counts, u32 *__restrict__ totals) {
    __shared__ u32 s_warp[FILT_THREADS / 32];
    const FilterJob job = jobs[blockIdx.x];
    const u32 n_chunks = (u32)((job.src_len + FILT_CHUNK - 1) / FILT_CHUNK);
    u32 *cc = chunk_counts + job.chunk_off;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32 carry = 0;
    for (u32 base = 0; base < n_chunks; base += FILT_THREADS) {
        const u32 c = base + tid;
        const u32 v = c < n_chunks ? cc[c] : 0u;
        u32 incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        u32 off = incl - v, tot = 0;
#pragma unroll
        for (int wI = 0; wI < FILT_THREADS / 32; wI++) { u32 x = s_warp[wI]; if (wI < (int)warp) off += x; tot += x; }
        if (c < n_chunks) cc[c] = carry + off;
        carry += tot;
    }
    if (tid == 0) totals[blockIdx.x] = carry;
}

__global__ void gather_rows_kernel(const float *__restrict__ dense, const u64 *__restrict__ rows, u64 n_rows,
                                   float *__restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows) out[i] = dense[rows[i]];
}

__global__ void count_docs_kernel(const u64 *__restrict__ list, u64 n, u32 *__restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool head = (i == 0) || ((list[i] >> SA_KEY_SHIFT) != (list[i - 1] >> SA_KEY_SHIFT));
    unsigned m = __ballot_sync(__activemask(), head);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (u32)__popc(m));
}

// Filters the given terms' lists into ix->filt; fills offs/lens (relative to ix->filt) per term.
// d_mask: per-doc keep bytes (NULL = no doc filter).  When d_df_out is given, the number of distinct
// docs of every filtered list is also returned (PosnBitArray.docfreq on FilteredPosns).
int sa_filter_terms_mask(sa_index *ix, const uint32_t *term_ids, uint32_t n_terms, const unsigned char *d_mask,
                         u64 pay_lo, u64 pay_hi, bool use_payload, std::vector<u64> &offs, std::vector<u64> &lens,
                         std::vector<u64> *df_out) {
    std::vector<FilterJob> jobs(n_terms);
    u64 total = 0, chunks_total = 0, max_chunks = 1;
    for (u32 t = 0; t < n_terms; t++) {
        const bool known = term_ids[t] != SA_NO_TERM;
        jobs[t].src_off = known ? ix->h_off[term_ids[t]] : 0;
        jobs[t].src_len = known ? ix->h_len[term_ids[t]] : 0;
        jobs[t].dst_off = total;
        jobs[t].chunk_off = chunks_total;
        const u64 nc = (jobs[t].src_len + FILT_CHUNK - 1) / FILT_CHUNK;
        chunks_total += nc;
        max_chunks = std::max(max_chunks, nc);
        total += jobs[t].src_len + 2;
    }
    int rc;
    if ((rc = ix->filt.reserve((total + 4) * sizeof(u64)))) return rc;
    const size_t jobs_bytes = ((size_t)n_terms * sizeof(FilterJob) + 255) / 256 * 256;
    if ((rc = ix->misc.reserve(jobs_bytes + (chunks_total + 2 * (size_t)n_terms + 8) * sizeof(u32)))) return rc;
    FilterJob *d_jobs = ix->misc.as<FilterJob>();
    u32 *d_chunks = (u32 *)((char *)ix->misc.p + jobs_bytes);
    u32 *d_totals = d_chunks + chunks_total;
    u32 *d_df = d_totals + n_terms;
    SA_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), n_terms * sizeof(FilterJob), cudaMemcpyHostToDevice, ix->stream));
    dim3 grid((unsigned)max_chunks, n_terms);
    filter_lists_kernel<false><<<grid, FILT_THREADS, 0, ix->stream>>>(ix->d_words, d_jobs, ix->filt.as<u64>(), d_chunks, d_mask,
                                                                     ix->doc_base, ix->n_docs, pay_lo, pay_hi, use_payload ? 1 : 0);
    SA_CUDA(cudaGetLastError());
    filter_scan_kernel<<<n_terms, FILT_THREADS, 0, ix->stream>>>(d_jobs, d_chunks, d_totals);
    SA_CUDA(cudaGetLastError());
    filter_lists_kernel<true><<<grid, FILT_THREADS, 0, ix->stream>>>(ix->d_words, d_jobs, ix->filt.as<u64>(), d_chunks, d_mask,
                                                                    ix->doc_base, ix->n_docs, pay_lo, pay_hi, use_payload ? 1 : 0);
    SA_CUDA(cudaGetLastError());
    ix->stats.total_launches += 3;
    std::vector<u32> h_counts(n_terms);
    SA_CUDA(cudaMemcpyAsync(h_counts.data(), d_totals, n_terms * sizeof(u32), cudaMemcpyDeviceToHost, ix->stream));
    SA_CUDA(cudaStreamSynchronize(ix->stream));
    offs.resize(n_terms);
    lens.resize(n_terms);
    for (u32 t = 0; t < n_terms; t++) { offs[t] = jobs[t].dst_off; lens[t] = h_counts[t]; }
    if (df_out) {
        df_out->assign(n_terms, 0);
        SA_CUDA(cudaMemsetAsync(d_df, 0, n_terms * sizeof(u32), ix->stream));
        for (u32 t = 0; t < n_terms; t++) {
            if (lens[t] == 0) continue;
            count_docs_kernel<<<(unsigned)((lens[t] + 255) / 256), 256, 0, ix->stream>>>(ix->filt.as<u64>() + offs[t], lens[t], d_df + t);
            SA_CUDA(cudaGetLastError());
            ix->stats.total_launches++;
        }
        SA_CUDA(cudaMemcpyAsync(h_counts.data(), d_df, n_terms * sizeof(u32), cudaMemcpyDeviceToHost, ix->stream));
        SA_CUDA(cudaStreamSynchronize(ix->stream));
        for (u32 t = 0; t < n_terms; t++) (*df_out)[t] = h_counts[t];
    }
    return SA_OK;
}

int sa_filter_terms(sa_index *ix, const uint32_t *term_ids, uint32_t n_terms, bool use_rows,
                    u64 pay_lo, u64 pay_hi, bool use_payload, std::vector<u64> &offs, std::vector<u64> &lens) {
    return sa_filter_terms_mask(ix, term_ids, n_terms, use_rows ? ix->d_row_mask : nullptr, pay_lo, pay_hi, use_payload,
                                offs, lens, nullptr);
}

int sa_gather_rows(sa_index *ix, const float *d_dense, float *out_host) {
    int rc;
    if ((rc = ix->gather.reserve(ix->n_rows * sizeof(float) + 64))) return rc;
    gather_rows_kernel<<<(unsigned)((ix->n_rows + 255) / 256), 256, 0, ix->stream>>>(d_dense, ix->d_rows, ix->n_rows,
                                                                                   ix->gather.as<float>());
    SA_CUDA(cudaGetLastError());
    ix->stats.total_launches++;
    SA_CUDA(cudaMemcpyAsync(out_host, ix->gather.p, ix->n_rows * sizeof(float), cudaMemcpyDeviceToHost, ix->stream));
    SA_CUDA(cudaStreamSynchronize(ix->stream));
    return SA_OK;
}

extern "C" int sa_index_set_rows(sa_index *ix, const uint64_t *rows, uint64_t n_rows) {
    SA_CHECK(ix, "index is NULL");
    std::lock_guard<std::mutex> g(ix->mu);
    SA_CUDA(cudaSetDevice(ix->device));
    if (rows == nullptr) {          // clear the filter
        SA_CHECK(n_rows == 0, "rows is NULL");
        ix->n_rows = 0;
        ix->rows_active = false;
        return SA_OK;
    }
    std::vector<unsigned char> mask(ix->n_docs, 0);
    for (u64 i = 0; i < n_rows; i++) {
        SA_CHECK(rows[i] < ix->n_docs, "row %llu out of range", (unsigned long long)rows[i]);
        mask[rows[i]] = 1;
    }
    cudaFree(ix->d_rows);
    ix->d_rows = nullptr;
    SA_CUDA(cudaMalloc(&ix->d_rows, std::max<u64>(n_rows, 1) * sizeof(u64)));
    if (!ix->d_row_mask) SA_CUDA(cudaMalloc(&ix->d_row_mask, std::max<u64>(ix->n_docs, 1)));
    if (n_rows) SA_CUDA(cudaMemcpyAsync(ix->d_rows, rows, n_rows * sizeof(u64), cudaMemcpyHostToDevice, ix->stream));
    if (ix->n_docs) SA_CUDA(cudaMemcpyAsync(ix->d_row_mask, mask.data(), ix->n_docs, cudaMemcpyHostToDevice, ix->stream));
    SA_CUDA(cudaStreamSynchronize(ix->stream));
    ix->n_rows = n_rows;
    ix->rows_active = true;
    return SA_OK;
}

// PosnBitArray.docfreq on FilteredPosns (quirk iii: df of the filtered postings)
extern "C" int sa_docfreq_rows(sa_index *ix, uint32_t term_id, uint64_t *df_out) {
    SA_CHECK(ix && df_out, "NULL argument");
    if (term_id == SA_NO_TERM) { *df_out = 0; return SA_OK; }
    SA_CHECK(term_id < ix->n_terms, "term id %u out of range", term_id);
    std::lock_guard<std::mutex> g(ix->mu);
    SA_CUDA(cudaSetDevice(ix->device));
    if (!ix->rows_active) { *df_out = ix->h_df[term_id]; return SA_OK; }
    std::vector<u64> offs, lens, dfs;
    int rc = sa_filter_terms_mask(ix, &term_id, 1, ix->d_row_mask, 0, SA_ALL_BITS, false, offs, lens, &dfs);
    if (rc) return rc;
    *df_out = dfs[0];
    return SA_OK;
}

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

struct FilterJob { u64 src_off, src_len, dst_off; };

__global__ void __launch_bounds__(256)
filter_lists_kernel(const u64 *__restrict__ words, const FilterJob *__restrict__ jobs, u64 *__restrict__ dst,
                    u32 *__restrict__ counts, const unsigned char *__restrict__ row_mask, u64 doc_base, u64 n_docs,
                    u64 pay_lo, u64 pay_hi, int use_payload) {
    __shared__ u32 s_warp[8];
    __shared__ u32 s_base;
    const FilterJob job = jobs[blockIdx.x];
    const u64 *src = words + job.src_off;
    u64 *out = dst + job.dst_off;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (u64 base = 0; base < job.src_len; base += 256) {
        const u64 i = base + tid;
        bool keep = false;
        u64 w = 0;
        if (i < job.src_len) {
            w = src[i];
            keep = true;
            if (row_mask) {
                const u64 d = (w >> SA_KEY_SHIFT) - doc_base;
                keep = d < n_docs && row_mask[d];
            }
            if (keep && use_payload) {
                const u64 v = w & SA_MSB_MASK;          // UNSHIFTED compare, reference roaringish_ops.pyx:55
                keep = v >= pay_lo && v <= pay_hi;
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        u32 off = 0, tot = 0;
        for (int wI = 0; wI < 8; wI++) { u32 c = s_warp[wI]; if (wI < (int)warp) off += c; tot += c; }
        const u32 b0 = s_base;
        if (keep) out[b0 + off + __popc(m & ((1u << lane) - 1))] = w;
        __syncthreads();
        if (tid == 0) s_base = b0 + tot;
        __syncthreads();
    }
    if (tid == 0) counts[blockIdx.x] = s_base;
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
int sa_filter_terms(sa_index *ix, const uint32_t *term_ids, uint32_t n_terms, bool use_rows,
                    u64 pay_lo, u64 pay_hi, bool use_payload, std::vector<u64> &offs, std::vector<u64> &lens) {
    std::vector<FilterJob> jobs(n_terms);
    u64 total = 0;
    for (u32 t = 0; t < n_terms; t++) {
        jobs[t].src_off = ix->h_off[term_ids[t]];
        jobs[t].src_len = ix->h_len[term_ids[t]];
        jobs[t].dst_off = total;
        total += jobs[t].src_len + 2;
    }
    int rc;
    if ((rc = ix->filt.reserve((total + 4) * sizeof(u64)))) return rc;
    if ((rc = ix->misc.reserve(n_terms * (sizeof(FilterJob) + sizeof(u32)) + 64))) return rc;
    FilterJob *d_jobs = ix->misc.as<FilterJob>();
    u32 *d_counts = (u32 *)(d_jobs + n_terms);
    SA_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), n_terms * sizeof(FilterJob), cudaMemcpyHostToDevice, ix->stream));
    filter_lists_kernel<<<n_terms, 256, 0, ix->stream>>>(ix->d_words, d_jobs, ix->filt.as<u64>(), d_counts,
                                                         use_rows ? ix->d_row_mask : nullptr, ix->doc_base, ix->n_docs,
                                                         pay_lo, pay_hi, use_payload ? 1 : 0);
    SA_CUDA(cudaGetLastError());
    ix->stats.total_launches++;
    std::vector<u32> h_counts(n_terms);
    SA_CUDA(cudaMemcpyAsync(h_counts.data(), d_counts, n_terms * sizeof(u32), cudaMemcpyDeviceToHost, ix->stream));
    SA_CUDA(cudaStreamSynchronize(ix->stream));
    offs.resize(n_terms);
    lens.resize(n_terms);
    for (u32 t = 0; t < n_terms; t++) { offs[t] = jobs[t].dst_off; lens[t] = h_counts[t]; }
    return SA_OK;
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
    std::vector<u64> offs, lens;
    int rc = sa_filter_terms(ix, &term_id, 1, true, 0, SA_ALL_BITS, false, offs, lens);
    if (rc) return rc;
    *df_out = 0;
    if (lens[0] == 0) return SA_OK;
    if ((rc = ix->misc.reserve(256))) return rc;
    SA_CUDA(cudaMemsetAsync(ix->misc.p, 0, sizeof(u32), ix->stream));
    count_docs_kernel<<<(unsigned)((lens[0] + 255) / 256), 256, 0, ix->stream>>>(ix->filt.as<u64>() + offs[0], lens[0], ix->misc.as<u32>());
    SA_CUDA(cudaGetLastError());
    ix->stats.total_launches++;
    u32 df = 0;
    SA_CUDA(cudaMemcpyAsync(&df, ix->misc.p, sizeof(u32), cudaMemcpyDeviceToHost, ix->stream));
    SA_CUDA(cudaStreamSynchronize(ix->stream));
    *df_out = df;
    return SA_OK;
}

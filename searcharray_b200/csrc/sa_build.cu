// sa_build.cu -- index build on the device.
//
// Replaces the numpy half of the reference's index build AFTER tokenisation (which stays on the host: it is a
// Python loop over strings, searcharray/indexing.py:64-98):
//   _invert_docs_terms / _lex_sort     searcharray/indexing.py:101-115   stable sort of the (term, doc, posn) triples by term
//   RoaringishEncoder.encode           searcharray/roaringish/roaringish.py:93-142   header = doc << 36 | (posn // 18) << 18,
//                                      one bit per posn % 18, np.bitwise_or.reduceat over equal headers
//   PosnBitArrayFromFlatBuilder.build  searcharray/phrase/middle_out.py (term boundaries -> ArrayDict slices)
//
// Triples arrive in document order (docs ascending, positions ascending inside a doc), exactly as _gather_tokens
// emits them.  The only sort needed is the STABLE sort by term id: a least-significant-digit radix sort of
// (term id, original index) pairs -- cub::DeviceRadixSort, NVIDIA's library sort shipped with the CUDA toolkit, used
// as a plain library primitive the way a BLAS call would be; everything around it (header / bit construction,
// segmented OR by head flags, compaction, term slices) is this file's kernels.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <vector>

#include "sa_common.cuh"

namespace {

struct DevMemB {
    std::vector<void *> ptrs;
    ~DevMemB() { for (void *p : ptrs) cudaFree(p); }
    template <typename T> T *alloc(size_t n) {
        void *p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T) + 64) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return (T *)p;
    }
};

__global__ void iota_kernel(u32 *__restrict__ a, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (u32)i;
}

// sorted position i: term = terms_sorted[i], triple = original index order[i].  A position starts a WORD when its
// (term, doc, posn // 18) differs from the previous position's; the head ORs the bits of its run (<= 18 entries:
// distinct positions of one block; repeated positions just OR the same bit again).
__global__ void word_head_kernel(const u32 *__restrict__ terms_sorted, const u32 *__restrict__ order,
                                 const u32 *__restrict__ docs, const u32 *__restrict__ posns, u64 n,
                                 u32 *__restrict__ head) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool h = i == 0;
    if (!h) {
        const u32 a = order[i], b = order[i - 1];
        h = terms_sorted[i] != terms_sorted[i - 1] || docs[a] != docs[b] || posns[a] / SA_LSB_BITS != posns[b] / SA_LSB_BITS;
    }
    head[i] = h ? 1u : 0u;
}

// block-local exclusive scan of head flags + per-block totals (1024 entries per block)
__global__ void __launch_bounds__(256)
build_scan_kernel(const u32 *__restrict__ flags, u32 *__restrict__ offs, u64 n, u32 *__restrict__ bsum) {
    __shared__ u32 warp_sums[8];
    const u64 base = (u64)blockIdx.x * 1024 + (u64)threadIdx.x * 4;
    u32 v[4], sum = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) { v[e] = (base + e < n) ? flags[base + e] : 0u; sum += v[e]; }
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    u32 wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { if (w < (int)warp) wbase += warp_sums[w]; total += warp_sums[w]; }
    u32 run = wbase + incl - sum;
#pragma unroll
    for (int e = 0; e < 4; e++) { if (base + e < n) offs[base + e] = run; run += v[e]; }
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
build_bsum_kernel(u32 *__restrict__ bsum, u32 n_blocks, u32 *__restrict__ total_out) {
    __shared__ u32 warp_sums[32];
    __shared__ u32 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 b0 = 0; b0 < n_blocks; b0 += 1024) {
        const u32 i = b0 + threadIdx.x;
        const u32 v = i < n_blocks ? bsum[i] : 0u;
        const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        u32 incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        u32 wbase = 0, tot = 0;
        for (int w = 0; w < 32; w++) { if (w < (int)warp) wbase += warp_sums[w]; tot += warp_sums[w]; }
        const u32 c = carry;
        if (i < n_blocks) bsum[i] = c + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

// every head writes its word at its rank; the first word of a term records the term's slice start, and every
// head bumps its term's length
__global__ void word_write_kernel(const u32 *__restrict__ terms_sorted, const u32 *__restrict__ order,
                                  const u32 *__restrict__ docs, const u32 *__restrict__ posns, u64 n,
                                  const u32 *__restrict__ head, const u32 *__restrict__ offs, const u32 *__restrict__ bsum,
                                  u64 *__restrict__ words, u64 *__restrict__ term_off, u64 *__restrict__ term_len) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    const u32 rank = offs[i] + bsum[i / 1024];
    const u32 term = terms_sorted[i];
    const u32 a = order[i];
    const u32 doc = docs[a], blk = posns[a] / SA_LSB_BITS;
    u64 bits = 0;
    for (u64 j = i; j < n; j++) {
        if (j > i && head[j]) break;
        bits |= 1ull << (posns[order[j]] % SA_LSB_BITS);
    }
    words[rank] = ((u64)doc << SA_KEY_SHIFT) | ((u64)blk << SA_LSB_BITS) | bits;
    if (i == 0 || terms_sorted[i - 1] != term) term_off[term] = rank;
    atomicAdd((unsigned long long *)&term_len[term], 1ull);
}

}  // namespace

// (term, doc, posn) triples in document order -> the index's upload format: posting words of all terms
// concatenated in term-id order, sorted and header-unique inside a term, plus every term's slice.
// words_out needs room for n_triples words; term_off_out / term_len_out have n_terms entries (absent terms: 0, 0).
extern "C" int sa_op_build_index(const uint32_t *term_ids, const uint32_t *doc_ids, const uint32_t *posns, uint64_t n_triples,
                                 uint32_t n_terms, int device, uint64_t *words_out, uint64_t *n_words_out,
                                 uint64_t *term_off_out, uint64_t *term_len_out) {
    SA_CHECK(n_words_out && (n_terms == 0 || (term_off_out && term_len_out)), "NULL argument");
    *n_words_out = 0;
    for (u32 t = 0; t < n_terms; t++) { term_off_out[t] = 0; term_len_out[t] = 0; }
    if (n_triples == 0) return SA_OK;
    SA_CHECK(term_ids && doc_ids && posns && words_out, "NULL argument");
    SA_CHECK(n_triples < (1ull << 32), "too many tokens for one build call (batch them like the reference's batch_size)");
    SA_CUDA(cudaSetDevice(device));
    const u64 n = n_triples;
    DevMemB m;
    u32 *d_t = m.alloc<u32>(n), *d_ts = m.alloc<u32>(n), *d_i = m.alloc<u32>(n), *d_is = m.alloc<u32>(n);
    u32 *d_doc = m.alloc<u32>(n), *d_pos = m.alloc<u32>(n), *d_head = m.alloc<u32>(n), *d_offs = m.alloc<u32>(n);
    const u32 n_blocks = (u32)((n + 1023) / 1024);
    u32 *d_bsum = m.alloc<u32>(n_blocks + 1);
    u64 *d_words = m.alloc<u64>(n), *d_toff = m.alloc<u64>(n_terms), *d_tlen = m.alloc<u64>(n_terms);
    if (!(d_t && d_ts && d_i && d_is && d_doc && d_pos && d_head && d_offs && d_bsum && d_words && d_toff && d_tlen)) {
        sa_set_error("device allocation failed");
        return SA_ERR_NOMEM;
    }
    SA_CUDA(cudaMemcpy(d_t, term_ids, n * sizeof(u32), cudaMemcpyHostToDevice));
    SA_CUDA(cudaMemcpy(d_doc, doc_ids, n * sizeof(u32), cudaMemcpyHostToDevice));
    SA_CUDA(cudaMemcpy(d_pos, posns, n * sizeof(u32), cudaMemcpyHostToDevice));
    SA_CUDA(cudaMemset(d_toff, 0, std::max<u32>(n_terms, 1) * sizeof(u64)));
    SA_CUDA(cudaMemset(d_tlen, 0, std::max<u32>(n_terms, 1) * sizeof(u64)));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    iota_kernel<<<blocks, 256>>>(d_i, n);
    // stable LSD radix sort by term id (only the bits a term id can use)
    int end_bit = 1;
    while (end_bit < 32 && (1ull << end_bit) < (u64)std::max<u32>(n_terms, 2)) end_bit++;
    size_t tmp_bytes = 0;
    SA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_t, d_ts, d_i, d_is, (int)n, 0, end_bit));
    void *d_tmp = m.alloc<char>(tmp_bytes);
    if (!d_tmp) { sa_set_error("device allocation failed"); return SA_ERR_NOMEM; }
    SA_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_t, d_ts, d_i, d_is, (int)n, 0, end_bit));
    word_head_kernel<<<blocks, 256>>>(d_ts, d_is, d_doc, d_pos, n, d_head);
    build_scan_kernel<<<n_blocks, 256>>>(d_head, d_offs, n, d_bsum);
    build_bsum_kernel<<<1, 1024>>>(d_bsum, n_blocks, d_bsum + n_blocks);
    word_write_kernel<<<blocks, 256>>>(d_ts, d_is, d_doc, d_pos, n, d_head, d_offs, d_bsum, d_words, d_toff, d_tlen);
    SA_CUDA(cudaGetLastError());
    u32 total = 0;
    SA_CUDA(cudaMemcpy(&total, d_bsum + n_blocks, sizeof(u32), cudaMemcpyDeviceToHost));
    SA_CUDA(cudaMemcpy(words_out, d_words, (size_t)total * sizeof(u64), cudaMemcpyDeviceToHost));
    if (n_terms) {
        SA_CUDA(cudaMemcpy(term_off_out, d_toff, n_terms * sizeof(u64), cudaMemcpyDeviceToHost));
        SA_CUDA(cudaMemcpy(term_len_out, d_tlen, n_terms * sizeof(u64), cudaMemcpyDeviceToHost));
    }
    *n_words_out = total;
    return SA_OK;
}

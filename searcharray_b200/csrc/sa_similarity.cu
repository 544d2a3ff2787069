// sa_similarity.cu -- the reference's non-default similarities as device kernels.
//
// Replaces the numpy closures of searcharray/similarity.py:41-89 (bm25_impact,
// bm25_legacy_similarity, classic_similarity).  What parity depends on is numpy's dtype promotion:
// the Python-float parameters become float32 next to the float32 arrays (k1, b, `1 - b` and `k1 + 1`
// are computed in double and THEN rounded), the idf scalars are float64 and make the final product
// float64 (legacy, classic).  Every operation is individually rounded (-fmad=false, *_rn).
#include <cmath>

#include "sa_common.cuh"

struct SimArgs {
    const float *tf, *dl;
    u64 n;
    float k1, b, one_minus_b, k1_plus_1, avgdl;
    double idf;
    void *out;
};

__device__ __forceinline__ float saturation_denominator(float tf, float dl, const SimArgs &a) {
    // tf + k1 * (1 - b + b * doc_lens / avg_doc_lens), left to right as numpy evaluates it
    const float ratio = __fdiv_rn(__fmul_rn(a.b, dl), a.avgdl);
    return __fadd_rn(tf, __fmul_rn(a.k1, __fadd_rn(a.one_minus_b, ratio)));
}

template <int KIND>
__global__ void __launch_bounds__(256) similarity_kernel(const SimArgs a) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const float tf = a.tf[i], dl = a.dl[i];
    if (KIND == SA_SIM_BM25_IMPACT) {
        ((float *)a.out)[i] = __fdiv_rn(tf, saturation_denominator(tf, dl, a));
    } else if (KIND == SA_SIM_BM25_LEGACY) {
        const float sat = __fdiv_rn(__fmul_rn(tf, a.k1_plus_1), saturation_denominator(tf, dl, a));
        ((double *)a.out)[i] = __dmul_rn(a.idf, (double)sat);
    } else {
        const float length_norm = __fdiv_rn(1.0f, __fsqrt_rn(dl));
        ((double *)a.out)[i] = __dmul_rn(__dmul_rn(a.idf, (double)__fsqrt_rn(tf)), (double)length_norm);
    }
}

extern "C" int sa_op_similarity(int kind, const float *term_freqs, const float *doc_lens, uint64_t n,
                                double avg_doc_len, double idf, double k1, double b, int device, void *out) {
    if (n == 0) return SA_OK;
    SA_CHECK(term_freqs && doc_lens && out, "NULL argument");
    SA_CHECK(kind == SA_SIM_BM25_IMPACT || kind == SA_SIM_BM25_LEGACY || kind == SA_SIM_CLASSIC, "unknown similarity %d", kind);
    SA_CUDA(cudaSetDevice(device));
    const size_t out_bytes = n * (kind == SA_SIM_BM25_IMPACT ? sizeof(float) : sizeof(double));
    float *d_tf = nullptr, *d_dl = nullptr;
    void *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_tf, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_dl, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, out_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(d_tf, term_freqs, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_dl, doc_lens, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        SimArgs a;
        a.tf = d_tf; a.dl = d_dl; a.n = n;
        a.k1 = (float)k1; a.b = (float)b;
        a.one_minus_b = (float)(1 - b);          // Python: `1 - b` in double, rounded when it meets the array
        a.k1_plus_1 = (float)(k1 + 1);
        a.avgdl = (float)avg_doc_len;
        a.idf = idf;
        a.out = d_out;
        const unsigned blocks = (unsigned)((n + 255) / 256);
        if (kind == SA_SIM_BM25_IMPACT) similarity_kernel<SA_SIM_BM25_IMPACT><<<blocks, 256>>>(a);
        else if (kind == SA_SIM_BM25_LEGACY) similarity_kernel<SA_SIM_BM25_LEGACY><<<blocks, 256>>>(a);
        else similarity_kernel<SA_SIM_CLASSIC><<<blocks, 256>>>(a);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(out, d_out, out_bytes, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_tf);
    cudaFree(d_dl);
    cudaFree(d_out);
    if (e != cudaSuccess) { sa_set_error("sa_op_similarity: %s", cudaGetErrorString(e)); return SA_ERR_CUDA; }
    return SA_OK;
}

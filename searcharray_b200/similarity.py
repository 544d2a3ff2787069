"""Similarity plug-ins, mirroring reference searcharray/similarity.py.

`Similarity` is the same callable protocol (similarity.py:8-16).  `bm25_similarity(k1, b)`
returns a callable object that ALSO carries (k1, b): SearchArray.score recognises it and runs
the fused GPU kernel (postings -> tf -> BM25 in one launch) instead of calling it on a host
vector.  Any other callable is treated as a user plug-in and receives the GPU-computed dense
term-frequency vector on the host, exactly like the reference.
"""
from typing import Protocol

import numpy as np


class Similarity(Protocol):
    def __call__(self, term_freqs, doc_freqs, doc_lens, avg_doc_lens, num_docs) -> np.ndarray:
        ...


def compute_idf(num_docs, dfs):
    """reference similarity.py:19-21 -- float64 on the host (negligible)."""
    dfs = np.asarray(dfs)
    return np.sum(np.log(1 + (num_docs - dfs + 0.5) / (dfs + 0.5)))


class Bm25Similarity:
    """BM25 as in Lucene 9 (reference similarity.py:24-38), evaluated on the GPU."""

    def __init__(self, k1=1.2, b=0.75):
        self.k1 = k1
        self.b = b

    def __call__(self, term_freqs, doc_freqs, doc_lens, avg_doc_lens, num_docs):
        # Protocol entry point for callers that already hold a host tf vector (e.g. an
        # edismax-style combiner): still the CUDA op, never numpy.
        from . import ops
        if avg_doc_lens == 0:
            return np.zeros_like(term_freqs)
        idf = compute_idf(num_docs, doc_freqs)
        return ops.bm25_score(term_freqs, doc_lens, avg_doc_lens, idf, self.k1, self.b)

    def __repr__(self):
        return f"bm25_similarity(k1={self.k1}, b={self.b})"


def bm25_similarity(k1: float = 1.2, b: float = 0.75) -> Bm25Similarity:
    return Bm25Similarity(k1, b)


default_bm25 = bm25_similarity()

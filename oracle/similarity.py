"""CPU restatement of the reference's non-default similarities (searcharray/similarity.py:41-89).
TEST INFRASTRUCTURE ONLY.  Pinned by tests/golden/similarity.npz (real reference output).

What parity depends on is numpy's dtype promotion: Python floats (k1, b, 1 - b, k1 + 1) are weak
and become float32 next to the float32 arrays; the idf scalars are np.float64 and make the final
product float64 (legacy, classic); bm25_impact stays float32.
"""
import numpy as np


def idf_bm25(num_docs, dfs):
    """similarity.py:19-21."""
    dfs = np.asarray(dfs)
    return np.sum(np.log(1 + (num_docs - dfs + 0.5) / (dfs + 0.5)))


def _saturation_denominator(tf, doc_lens, avg_doc_lens, k1, b):
    f32 = np.float32
    return tf + f32(k1) * (f32(1 - b) + (f32(b) * doc_lens) / f32(avg_doc_lens))


def bm25_impact(tf, dfs, doc_lens, avg_doc_lens, num_docs, k1=1.2, b=0.75):
    """similarity.py:41-54: BM25 without the idf (float32)."""
    if avg_doc_lens == 0:
        return np.zeros_like(tf)
    return tf / _saturation_denominator(tf, doc_lens, avg_doc_lens, k1, b)


def bm25_legacy(tf, dfs, doc_lens, avg_doc_lens, num_docs, k1=1.2, b=0.75):
    """similarity.py:57-72: pre-LUCENE-8563 BM25 with (k1 + 1) in the numerator (float64 result)."""
    if avg_doc_lens == 0:
        return np.zeros_like(tf)
    sat = (tf * np.float32(k1 + 1)) / _saturation_denominator(tf, doc_lens, avg_doc_lens, k1, b)
    return np.float64(idf_bm25(num_docs, dfs)) * sat.astype(np.float64)


def classic(tf, dfs, doc_lens, avg_doc_lens, num_docs):
    """similarity.py:75-89: Lucene classic TF-IDF (float64 result)."""
    idf = np.log((num_docs + 1) / (np.sum(np.asarray(dfs), axis=0) + 1)) + 1
    with np.errstate(divide="ignore", invalid="ignore"):
        length_norm = np.float32(1.0) / np.sqrt(doc_lens)
        return (np.float64(idf) * np.sqrt(tf).astype(np.float64)) * length_norm.astype(np.float64)

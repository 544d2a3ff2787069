"""Runs the REAL reference (oracle/_ref, built by oracle/build_ref.py) on an injected index -- test
infrastructure and the CPU arm of bench.py; never on the product path.

`reference_array(host, ...)` wraps a HostIndex (the upload format of the GPU index: ArrayDict.data +
metadata, reference phrase/memmap_arrays.py:15-53) in the reference's own `SearchArray` exactly the
way `SearchArray.index` injects its build (reference postings.py:293-299): term_mat / posns /
term_dict / avg_doc_length / doc_lens / corpus_size.  Everything that then runs -- `.score`,
`.termfreqs`, `.docfreq`, PosnBitArray caches, the Cython kernels -- is the reference's stock code.
"""
import numpy as np

from .build_ref import have_ref, import_reference


def available():
    return have_ref()


def reference_array(host, avg_doc_length=None, corpus_size=None, names=None):
    """The reference's SearchArray over `host`'s postings (words are shared, not copied)."""
    sa = import_reference()
    from searcharray.phrase.memmap_arrays import ArrayDict
    from searcharray.phrase.middle_out import PosnBitArray
    from searcharray.term_dict import TermDict
    from searcharray.utils.mat_set import SparseMatSet
    from searcharray.utils.row_viewable_matrix import RowViewableMatrix

    n_docs, n_terms = host.n_docs, host.n_terms
    arr = sa.SearchArray([])
    td = TermDict()
    if names is None:
        names = [host.term_dict.get_term(t) for t in range(n_terms)]
    for nm in names:
        td.add_term(nm)
    boundaries = np.concatenate((host.term_offsets, [host.term_offsets[-1] + host.term_lengths[-1]] if n_terms else [0]))
    # the ArrayDict layout wants each term's slice [boundaries[i], boundaries[i+1]); HostIndex lists tile `words`
    posns = ArrayDict.from_array_with_boundaries(host.words, np.arange(n_terms), boundaries.astype(np.uint64))
    arr.posns = PosnBitArray(posns, max_doc_id=max(n_docs - 1, 0))
    # the term matrix is only consulted for `.rows` / `.subset` on this path (postings.py:617-636)
    arr.term_mat = RowViewableMatrix(SparseMatSet(cols=np.zeros(0, dtype=np.uint32),
                                                  rows=np.zeros(n_docs + 1, dtype=np.uint32)))
    arr.term_dict = td
    arr.doc_lens = host.doc_lens
    arr.avg_doc_length = host.avg_doc_length if avg_doc_length is None else avg_doc_length
    arr.corpus_size = n_docs if corpus_size is None else int(corpus_size)
    return arr


def bm25(k1=1.2, b=0.75):
    import_reference()
    from searcharray.similarity import bm25_similarity
    return bm25_similarity(k1=k1, b=b)

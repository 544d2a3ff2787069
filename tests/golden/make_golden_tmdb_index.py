"""TMDB-derived postings fixture (BASELINE configs[0]) so that the real corpus runs through the CUDA
path on the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden_tmdb_index.py      (build container only)

Indexes fixtures/tmdb.json.gz (27,846 documents, title + overview, default whitespace tokenizer) with
this repo's host indexer -- which tests/test_tmdb_cpu.py proves identical, word for word, to the
reference's own index (digests in tmdb.json) -- and stores the upload format of both fields in
tests/golden/tmdb_index.npz: posting words (delta-coded per term for compression), term slices,
doc lengths, average doc length, the term strings in id order.  The expected query results are the
reference's, already committed as digests / top-10 lists in tmdb.json (make_golden_tmdb.py).
"""
import os

import numpy as np

from make_golden import HERE
from make_golden_tmdb import load_corpus

import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from searcharray_b200.indexing import build_index  # noqa: E402


def main():
    titles, overviews = load_corpus()
    out = {}
    for name, docs in (("title_tokens", titles), ("overview_tokens", overviews)):
        host = build_index(docs, str.split)
        # delta-code the words inside each term's slice (sorted ascending): small integers compress well
        delta = host.words.copy()
        delta[1:] -= host.words[:-1]
        starts = host.term_offsets[host.term_lengths > 0].astype(np.int64)
        delta[starts] = host.words[starts]
        out[name + ".delta"] = delta
        out[name + ".offsets"] = host.term_offsets
        out[name + ".lengths"] = host.term_lengths
        out[name + ".doc_lens"] = host.doc_lens
        out[name + ".avg_doc_length"] = np.asarray(host.avg_doc_length)
        terms = [host.term_dict.get_term(t) for t in range(host.n_terms)]
        out[name + ".terms"] = np.frombuffer("\n".join(terms).encode("utf-8"), dtype=np.uint8)
        assert all("\n" not in t for t in terms)
        print(name, host.n_terms, len(host.words), host.avg_doc_length)
    path = os.path.join(HERE, "tmdb_index.npz")
    np.savez_compressed(path, **out)
    print(os.path.getsize(path))


if __name__ == "__main__":
    main()

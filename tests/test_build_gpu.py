"""the sort + roaringish encode of the index build on the device (sa_op_build_index) against the host
(numpy) indexer, which tests/test_tmdb_cpu.py proves identical to the reference's own index word for word."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def same_index(a, b):
    assert a.n_terms == b.n_terms and a.n_docs == b.n_docs
    assert np.array_equal(a.term_lengths, b.term_lengths)
    for t in range(a.n_terms):
        assert np.array_equal(a.term_words(t), b.term_words(t)), t
    assert np.array_equal(a.doc_lens, b.doc_lens) and a.avg_doc_length == b.avg_doc_length
    assert a.term_dict.term_to_ids == b.term_dict.term_to_ids


def test_device_build_matches_host_build_on_text():
    from searcharray_b200.indexing import build_index
    docs = json.load(open(os.path.join(GOLDEN, "api.json")))["docs"]
    docs = docs + ["", "a a a a a a a a a a a a a a a a a a a a a a a a a a", None and "" or "zz " * 300] + docs[::-1]
    same_index(build_index(docs, str.split, gpu_build=0), build_index(docs, str.split))


def test_device_build_random_and_scoring():
    from searcharray_b200 import SearchArray
    rng = np.random.default_rng(3)
    vocab = [f"t{i}" for i in range(200)]
    docs = [" ".join(rng.choice(vocab, size=int(rng.integers(0, 120)), p=None)) for _ in range(3000)]
    a = SearchArray.index(docs, gpu_build=True)
    b = SearchArray.index(docs)
    same_index(a.host, b.host)
    for q in ("t0", "t199", ["t1", "t2"], ["t3", "t3"]):
        assert np.array_equal(a.score(q), b.score(q))


def test_device_build_empty():
    from searcharray_b200.indexing import build_index
    same_index(build_index(["", ""], str.split, gpu_build=0), build_index(["", ""], str.split))

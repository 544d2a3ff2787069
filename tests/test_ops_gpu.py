"""Device per-op exports against golden vectors of the reference's Cython ops (tests/golden/ops.npz,
made by make_golden.py from the real reference): popcount64_reduce and bm25_score."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "ops.npz"))


def n_cases(g):
    return len([f for f in g.files if f.endswith("_lhs")])


def test_popcount64_reduce_on_device(g):
    from searcharray_b200 import ops
    for c in range(n_cases(g)):
        k = f"c{c}_"
        ids, cnt = ops.popcount64_reduce(g[k + "lhs"])
        assert np.array_equal(ids, g[k + "pcr_ids"]) and np.array_equal(cnt, g[k + "pcr_cnt"]), c
        assert cnt.dtype == np.float32 or len(cnt) == 0


def test_bm25_score_on_device(g):
    """bit-exact incl. the NaN / inf positions of docs with doc_len == 0"""
    from searcharray_b200 import ops
    for c in range(n_cases(g)):
        k = f"c{c}_"
        tf = g[k + "dense"].copy()
        dl = g[k + "bm25_dl"]
        ops.bm25_score(tf, dl, float(np.mean(dl)), 2.345, 1.2, 0.75)
        assert np.array_equal(tf.view(np.uint32), g[k + "bm25"].view(np.uint32)), c

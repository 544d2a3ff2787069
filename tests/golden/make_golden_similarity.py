"""Golden vectors for the reference's non-default similarities (searcharray/similarity.py:41-89:
bm25_impact, bm25_legacy_similarity, classic_similarity), produced by the REAL reference.

    python tests/golden/make_golden_similarity.py

Same rules as make_golden.py (build container only; inputs + outputs, no reference source).
"""
import os

import numpy as np

from make_golden import import_reference, HERE


def main():
    import_reference()
    from searcharray.similarity import bm25_impact, bm25_legacy_similarity, classic_similarity
    rng = np.random.default_rng(20260926)
    n = 4000
    tf = rng.integers(0, 9, size=n).astype(np.float32)
    tf[rng.random(n) < 0.5] = 0
    dl = np.clip(rng.lognormal(3.5, 0.6, size=n), 1, 500).astype(np.float32)
    dl[:5] = [1, 2, 3, 400, 7]
    out = {"tf": tf, "doc_lens": dl}
    cases = []
    for ci, (k1, b, avgdl, dfs, num_docs) in enumerate([
            (1.2, 0.75, np.float32(np.mean(dl)), [37], 4000),
            (0.9, 0.4, np.float32(41.7), [1200, 3], 4000),
            (2.0, 0.0, np.float32(10.0), [5, 6, 7], 123456),
            (1.2, 0.75, 2.7322686, [14], 8516)]):
        dfa = np.asarray(dfs, dtype=np.uint64)
        out[f"c{ci}_params"] = np.asarray([k1, b, float(avgdl), num_docs], dtype=np.float64)
        out[f"c{ci}_dfs"] = dfa
        out[f"c{ci}_impact"] = bm25_impact(k1=k1, b=b)(tf.copy(), dfa, dl, avgdl, num_docs)
        out[f"c{ci}_legacy"] = bm25_legacy_similarity(k1=k1, b=b)(tf.copy(), dfa, dl, avgdl, num_docs)
        out[f"c{ci}_classic"] = classic_similarity()(tf.copy(), dfa, dl, avgdl, num_docs)
        cases.append(ci)
        print(ci, out[f"c{ci}_impact"].dtype, out[f"c{ci}_legacy"].dtype, out[f"c{ci}_classic"].dtype)
    out["n_cases"] = np.asarray([len(cases)])
    np.savez_compressed(os.path.join(HERE, "similarity.npz"), **out)


if __name__ == "__main__":
    main()

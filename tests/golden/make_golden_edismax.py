"""Golden vectors for edismax (reference searcharray/solr.py), produced by the REAL reference.

    python tests/golden/make_golden_edismax.py

Same rules as make_golden.py: runs only in the build container, writes inputs + outputs
(tests/golden/edismax.npz, edismax.json), copies no reference source.  Seeded.
"""
import json
import os

import numpy as np

from make_golden import import_reference, HERE

VOCAB = [f"w{i}" for i in range(40)]


def lower_one_token(text):
    """A tokenizer that yields ONE token: makes edismax take its field-centric path."""
    return [text.lower()]


def make_docs(rng, n_docs, lo, hi, plant):
    probs = 1.0 / np.arange(1, len(VOCAB) + 1) ** 1.0
    probs /= probs.sum()
    docs = []
    for _ in range(n_docs):
        n = int(rng.integers(lo, hi))
        toks = list(rng.choice(VOCAB, size=n, p=probs))
        if n >= 4 and rng.random() < plant:
            at = int(rng.integers(0, n - 3))
            toks[at:at + 3] = ["w2", "w5", "w1"]
        if n >= 3 and rng.random() < plant:
            at = int(rng.integers(0, n - 2))
            toks[at:at + 2] = ["w0", "w3"]
        docs.append(" ".join(toks))
    return docs


CASES = [
    # name, kwargs
    ("plain", dict(q="w2 w5", qf=["title", "body"])),
    ("boosts_tie", dict(q="w2 w5 w1", qf=["title^2.5", "body^0.7"], tie=0.3)),
    ("mm2", dict(q="w2 w5 w1", qf=["title", "body^0.5"], mm=2, tie=0.1)),
    ("mm_pct", dict(q="w0 w3 w7 w9", qf=["title^1.0", "body^0.5"], mm="75%", tie=0.25)),
    ("mm_cond", dict(q="w0 w3 w7 w9 w11", qf=["title", "body"], mm="2<75%")),
    ("q_and", dict(q="w0 w3", qf=["title", "body"], q_op="AND")),
    ("pf", dict(q="w2 w5 w1", qf=["title", "body^0.5"], pf=["title^3", "body"], tie=0.3)),
    ("pf2", dict(q="w2 w5 w1", qf=["title", "body^0.5"], pf2=["body^2"], tie=0.3)),
    ("pf3", dict(q="w2 w5 w1 w0", qf=["title", "body^0.5"], pf3=["body", "title^0.25"])),
    ("pf_all", dict(q="w0 w3 w2 w5 w1", qf=["title^1.0", "body^0.5"], pf=["body"], pf2=["title", "body^0.5"],
                    pf3=["body^1.5"], mm=2, tie=0.3)),
    ("pf_all_mm1", dict(q="w2 w5 w1", qf=["title^1.0", "body^0.5"], pf=["body", "title"], pf2=["body"], pf3=["title"])),
    ("repeat_terms", dict(q="w1 w1 w2", qf=["title", "body"], pf=["body"], pf2=["body"], tie=0.5)),
    ("unknown_term", dict(q="w2 zzz w5", qf=["title", "body"], pf2=["body"], mm=1)),
    ("all_unknown", dict(q="zzz yyy", qf=["title", "body"], pf=["body"])),
    ("one_field", dict(q="w2 w5 w1", qf=["body"], pf=["body^2"], pf2=["body"], pf3=["body"], mm="100%")),
    ("single_term", dict(q="w4", qf=["title^2", "body"], pf=["body"], tie=0.2)),
    ("field_centric", dict(q="w2 W5", qf=["title", "tag^2"], tie=0.1)),
    ("field_centric_mm", dict(q="w2 w5 w1", qf=["body^0.5", "tag", "title"], mm=2, tie=0.3)),
]


def main():
    import_reference()
    import pandas as pd
    from searcharray import SearchArray
    from searcharray.solr import edismax
    from searcharray.similarity import bm25_similarity
    rng = np.random.default_rng(20260925)
    n = 1200
    title = make_docs(rng, n, 1, 8, 0.15)
    body = make_docs(rng, n, 5, 60, 0.25)
    tag = [t.split()[0] + " " + t.split()[-1] if rng.random() < 0.5 else "w2 w5" for t in title]
    frame = pd.DataFrame({"title": SearchArray.index(title, autowarm=False),
                          "body": SearchArray.index(body, autowarm=False),
                          "tag": SearchArray.index(tag, tokenizer=lower_one_token, autowarm=False)})
    out, meta = {}, {"title": title, "body": body, "tag": tag, "cases": [], "terms": {}}
    # index dump per field (pins the oracle test to the reference's own postings)
    for fname in ("title", "body", "tag"):
        arr = frame[fname].array
        terms = sorted(arr.term_dict.term_to_ids.keys(), key=lambda t: arr.term_dict.term_to_ids[t])
        meta["terms"][fname] = terms
        ws = [np.asarray(arr.posns.encoded_term_posns[arr.term_dict.get_term_id(t)], dtype=np.uint64) for t in terms]
        out[f"ix_{fname}_words"] = np.concatenate(ws)
        out[f"ix_{fname}_lens"] = np.asarray([len(w) for w in ws], dtype=np.uint64)
        out[f"ix_{fname}_doc_lens"] = np.asarray(arr.doc_lens, dtype=np.float32)
        out[f"ix_{fname}_avgdl"] = np.asarray([arr.avg_doc_length], dtype=np.float32)
    for name, kw in CASES:
        scores, explain = edismax(frame, **kw)
        out[name] = scores
        meta["cases"].append({"name": name, "kwargs": kw, "explain": explain, "dtype": str(scores.dtype),
                              "matches": int(np.count_nonzero(scores))})
        print(name, scores.dtype, int(np.count_nonzero(scores)), float(scores.max()))
    # a per-field similarity dict
    sims = {"title": bm25_similarity(k1=0.9, b=0.4), "body": bm25_similarity()}
    kw = dict(q="w2 w5 w1", qf=["title", "body^0.5"], pf=["body"], pf2=["title"], tie=0.3)
    scores, explain = edismax(frame, similarity=sims, **kw)
    out["sim_dict"] = scores
    meta["cases"].append({"name": "sim_dict", "kwargs": kw, "explain": explain, "dtype": str(scores.dtype),
                          "similarity": {"title": [0.9, 0.4], "body": [1.2, 0.75]},
                          "matches": int(np.count_nonzero(scores))})
    np.savez_compressed(os.path.join(HERE, "edismax.npz"), **out)
    with open(os.path.join(HERE, "edismax.json"), "w") as f:
        json.dump(meta, f)
    print("edismax cases", len(meta["cases"]))


if __name__ == "__main__":
    main()

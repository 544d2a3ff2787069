"""TMDB fixture (BASELINE configs[0]: the reference's own CPU-runnable case) through the REAL
reference: index digests and query results, so that the host indexer and the oracle can be pinned on
27,846 real documents.  The corpus itself (fixtures/tmdb.json.gz) stays in the reference tree; this
writes only digests, counts and top-10 lists to tests/golden/tmdb.json.

    python tests/golden/make_golden_tmdb.py      (build container only)

Queries follow the reference's test/test_tmdb.py:167-191, 315-321 and its edismax call (:230-241).
"""
import gzip
import hashlib
import json
import os

import numpy as np

from make_golden import import_reference, HERE

FIXTURE = "/root/reference/fixtures/tmdb.json.gz"
TERMS = ["Star", "Black", "the", "Wars", "of", "a", "zzzzunknown"]
PHRASES = [["Star", "Wars"], ["the", "the"], ["Black", "Mirror:"], ["this", "doesnt", "match", "anything"],
           ["teeeeerms", "dooooont", "exiiiiist"], ["of", "the"], ["in", "the", "the", "of"]]
EDISMAX = ["Star Wars", "the next generation", "bartender fights a cow and", "to be or not to be",
           "the quick brown fox jumps over the lazy dog", "bill and ted's excellent adventure",
           "thirty years after defeating the galactic empire", "a film about a daughter of a refugee family"]


def load_corpus():
    with gzip.open(FIXTURE) as f:
        raw = json.load(f)
    titles, overviews = [], []
    for doc_id in raw.keys():
        titles.append(raw[doc_id].get("title", "") or "")
        overviews.append(raw[doc_id].get("overview", "") or "")
    return titles, overviews


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def vec_record(v):
    v = np.asarray(v)
    order = np.lexsort((np.arange(len(v)), -v.astype(np.float64)))[:10]
    order = order[v[order] > 0]
    return {"dtype": str(v.dtype), "sha256": sha(v), "nonzero": int(np.count_nonzero(v)),
            "top_ids": [int(i) for i in order], "top_scores": [float(v[i]) for i in order]}


def index_digest(arr):
    """Canonical (term-string order) digest of the whole positional index."""
    h = hashlib.sha256()
    terms = sorted(arr.term_dict.term_to_ids.keys())
    n_words = 0
    for t in terms:
        w = np.asarray(arr.posns.encoded_term_posns[arr.term_dict.get_term_id(t)], dtype=np.uint64)
        h.update(t.encode("utf-8"))
        h.update(w.tobytes())
        n_words += len(w)
    return {"n_terms": len(terms), "n_words": n_words, "sha256": h.hexdigest(),
            "doc_lens_sha256": sha(np.asarray(arr.doc_lens, dtype=np.float32)),
            "avg_doc_length": float(arr.avg_doc_length)}


def main():
    import_reference()
    import pandas as pd
    from searcharray import SearchArray
    from searcharray.solr import edismax
    titles, overviews = load_corpus()
    frame = pd.DataFrame({"title_tokens": SearchArray.index(titles, autowarm=False),
                          "overview_tokens": SearchArray.index(overviews, autowarm=False)})
    out = {"n_docs": len(titles), "fields": {}, "edismax": []}
    for field in ("title_tokens", "overview_tokens"):
        arr = frame[field].array
        rec = {"index": index_digest(arr), "terms": {}, "phrases": [], "slop": []}
        for t in TERMS:
            rec["terms"][t] = {"df": int(arr.docfreq(t)), "tf": vec_record(arr.termfreqs(t)), "score": vec_record(arr.score(t))}
        for ph in PHRASES:
            rec["phrases"].append({"phrase": ph, "tf": vec_record(arr.termfreqs(ph)), "score": vec_record(arr.score(ph))})
        if field == "title_tokens":
            for ph, slop in ((["of", "the"], 2), (["Star", "Wars"], 1), (["the", "of"], 3)):
                rec["slop"].append({"phrase": ph, "slop": slop, "tf": vec_record(arr.termfreqs(ph, slop=slop))})
        out["fields"][field] = rec
        print(field, rec["index"]["n_terms"], rec["index"]["n_words"], rec["index"]["avg_doc_length"])
    kw = dict(mm=2, qf=["title_tokens^1.0", "overview_tokens^0.5"], pf=["title_tokens^1.0", "overview_tokens^0.5"],
              pf2=["title_tokens^1.0", "overview_tokens^0.5"], pf3=["title_tokens^1.0", "overview_tokens^0.5"], tie=0.3)
    out["edismax_kwargs"] = kw
    for q in EDISMAX:
        scores, explain = edismax(frame, q=q, **kw)
        out["edismax"].append({"q": q, "explain": explain, "scores": vec_record(scores)})
        print(q, int(np.count_nonzero(scores)))
    with open(os.path.join(HERE, "tmdb.json"), "w") as f:
        json.dump(out, f)
    print(os.path.getsize(os.path.join(HERE, "tmdb.json")))


if __name__ == "__main__":
    main()

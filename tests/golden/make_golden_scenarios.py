"""Known-answer scenarios of the reference's OWN tests, extracted by importing
the reference's test modules and capturing what their `SearchArray.index(...)` lambdas index:

    test/test_phrase_matches.py:17-194   32 phrase scenarios (docs, phrase, expected counts)
    test/test_slop_matches.py:7-72       8 slop scenarios (phrase, doc, slop, match)
    test/test_minmax_posns.py:5-42       4 min/max position scenarios
    test/test_search.py                  the term-frequency / doc-frequency / score assertions

    python tests/golden/make_golden_scenarios.py      (build container only)

Writes tests/golden/reference_scenarios.json: the scenario inputs, the expected values the
reference's tests assert, and -- for cross-checks the tests only express as properties -- what the
real reference returns (slop 1..3 counts, odd-slice counts, slop scores).  No reference source is
copied: the tables are read from the imported modules.
"""
import json
import os
import sys

import numpy as np

from make_golden import import_reference, HERE, SCRATCH


def period_compress(docs):
    """docs == base * times for the smallest such base (the reference tables repeat 4 docs x N)."""
    n = len(docs)
    for p in range(1, n + 1):
        if n % p == 0 and docs == docs[:p] * (n // p):
            return {"base": docs[:p], "times": n // p}
    return {"base": docs, "times": 1}


def main():
    import_reference()
    sys.path.insert(0, os.path.join(SCRATCH, "test"))
    from searcharray.postings import SearchArray
    captured = []
    real_index = SearchArray.index.__func__

    def capturing_index(cls, array, *a, **kw):
        captured.append(list(array))
        return real_index(cls, array, *a, **kw)
    SearchArray.index = classmethod(capturing_index)

    import test_phrase_matches as tpm
    import test_slop_matches as tsm
    import test_minmax_posns as tmm
    out = {"phrase": [], "slop": [], "minmax": []}

    for name, sc in tpm.scenarios.items():
        captured.clear()
        arr = sc["docs"]()
        docs = captured[-1]
        phrase = list(sc["phrase"])
        expected = [float(x) for x in sc["expected"]]
        tf = arr.termfreqs(phrase if len(phrase) > 1 else phrase[0])
        assert (tf == np.asarray(expected)).all(), name
        rec = {"name": name, "docs": period_compress(docs), "phrase": phrase, "expected": period_compress(expected)}
        if len(phrase) > 1 and len(docs) <= 2000:
            rec["slop"] = {str(s): period_compress([float(x) for x in arr.termfreqs(phrase, slop=s)]) for s in (1, 2, 3)}
            rec["odd_slice"] = period_compress([float(x) for x in arr[1::2].termfreqs(phrase)])
        out["phrase"].append(rec)

    for name, sc in tsm.scenarios.items():
        docs = [sc["doc"], " empty ", sc["doc"] + " " + sc["doc"], " empty"] * 100
        arr = real_index(SearchArray, docs)
        toks = arr.tokenizer(sc["phrase"])
        scores = {}
        for s in range(sc["slop"], max(sc["slop"], 10)):
            v = arr.score(toks, slop=s)
            assert np.all((v[::2] > 0) == sc["match"]) and np.all(v[1::2] == 0), (name, s)
            if s in (sc["slop"], sc["slop"] + 1, 9):
                scores[str(s)] = [float(x) for x in v[:4]]          # the corpus repeats every 4 docs
                assert np.array_equal(v, np.tile(v[:4], 100))
        out["slop"].append({"name": name, "phrase": toks, "doc": sc["doc"], "slop": sc["slop"],
                            "match": sc["match"], "scores_first4": scores})

    for name, sc in tmm.scenarios.items():
        captured.clear()
        arr = sc["docs"]()
        docs = captured[-1]
        tf = arr.termfreqs(sc["phrase"], min_posn=sc["min_posn"], max_posn=sc["max_posn"])
        assert (tf == np.asarray(sc["expected"])).all(), name
        out["minmax"].append({"name": name, "docs": period_compress(docs), "phrase": list(sc["phrase"]),
                              "min_posn": sc["min_posn"], "max_posn": sc["max_posn"],
                              "expected": period_compress([float(x) for x in sc["expected"]])})

    with open(os.path.join(HERE, "reference_scenarios.json"), "w") as f:
        json.dump(out, f)
    print({k: len(v) for k, v in out.items()}, os.path.getsize(os.path.join(HERE, "reference_scenarios.json")))


if __name__ == "__main__":
    main()

"""CPU: the oracle's edismax restatement (oracle/solr.py) against golden vectors produced by the
real reference (tests/golden/make_golden_edismax.py), and the mm mini-language against the known
answers of the reference's own tests (test/test_solr.py:13-66)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def lower_one_token(text):
    return [text.lower()]


def load_fields():
    from oracle import search, solr
    g = np.load(os.path.join(GOLDEN, "edismax.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "edismax.json")))
    fields = {}
    for fname in ("title", "body", "tag"):
        lens = g[f"ix_{fname}_lens"].astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(lens)])
        words = g[f"ix_{fname}_words"]
        tw = {i: words[offs[i]:offs[i + 1]] for i in range(len(lens))}
        tid = {t: i for i, t in enumerate(meta["terms"][fname])}
        idx = search.OracleIndex(tw, g[f"ix_{fname}_doc_lens"], avg_doc_length=g[f"ix_{fname}_avgdl"][0])
        fields[fname] = solr.OracleField(idx, tid, tokenizer=lower_one_token if fname == "tag" else str.split)
    return g, meta, fields


def test_oracle_edismax_matches_reference():
    from oracle import solr
    g, meta, fields = load_fields()
    for case in meta["cases"]:
        kw = dict(case["kwargs"])
        for f, (k1, b) in case.get("similarity", {}).items():
            fields[f].k1, fields[f].b = k1, b
        got = solr.edismax(fields, **kw)
        for f in fields.values():
            f.k1, f.b = 1.2, 0.75
        want = g[case["name"]]
        assert str(got.dtype) == case["dtype"], case["name"]
        assert np.array_equal(got, want), (case["name"], float(np.abs(got - want).max()))


MM_KNOWN = [(10, "50%", 5), (10, "150%", 10), (10, "-50%", 5), (10, "3", 3), (10, "-3", 7), (10, "15", 10),
            (10, "5<70%", 7), (10, "15<70%", 10), (10, "3<50% 5<30%", 3), (10, "2<2 5<3 7<40%", 4)]


@pytest.mark.parametrize("n,spec,want", MM_KNOWN)
def test_mm_known_answers(n, spec, want):
    from oracle import solr
    from searcharray_b200 import solr as psolr
    assert solr.parse_min_should_match(n, spec) == want
    assert psolr.parse_min_should_match(n, spec) == want


@pytest.mark.parametrize("spec", ["five%", "five", "5<", ""])
def test_mm_invalid(spec):
    from oracle import solr
    from searcharray_b200 import solr as psolr
    with pytest.raises(ValueError):
        solr.parse_min_should_match(10, spec)
    with pytest.raises(ValueError):
        psolr.parse_min_should_match(10, spec)


def test_explain_strings_and_plan_on_host():
    """The host half of searcharray_b200.solr (query parsing, phase lists, explain strings) needs
    no GPU: compare the explain strings with the real reference's (golden)."""
    import pandas as pd
    from searcharray_b200 import SearchArray
    from searcharray_b200 import solr as psolr
    meta = json.load(open(os.path.join(GOLDEN, "edismax.json")))
    frame = pd.DataFrame({"title": SearchArray.index(meta["title"]),
                          "body": SearchArray.index(meta["body"]),
                          "tag": SearchArray.index(meta["tag"], tokenizer=lower_one_token)})
    for case in meta["cases"]:
        kw = dict(case["kwargs"])
        plan = psolr._Plan(frame, kw["q"], kw["qf"], kw.get("mm"), kw.get("pf"), kw.get("pf2"), kw.get("pf3"),
                           kw.get("tie", 0.0), kw.get("q_op", "OR"), psolr.default_bm25)
        assert plan.explain_qf() + plan.explain_phases() == case["explain"], case["name"]
        assert plan.term_centric == (case["dtype"] == "float64")
    assert psolr.parse_field_boosts(["title^2.5", "body"]) == {"title": 2.5, "body": None}
    with pytest.raises(ValueError):
        psolr.get_field(frame, "nope")

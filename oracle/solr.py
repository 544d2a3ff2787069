"""CPU restatement of the reference's edismax (searcharray/solr.py) on top of OracleIndex.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the product never imports this.  Pinned by
tests/golden/edismax.npz, generated from the real reference (tests/golden/make_golden_edismax.py).

What parity depends on (reference searcharray/solr.py):
  * :117-147 term-centric qf: per term position, over the fields in qf order, `score * boost` stays
    float32, the running sum and maximum are float64; term = max + (sum - max) * tie; a doc matches
    when at least `mm` terms score > 0; qf = sum of the term vectors in term order.
  * :150-178 field-centric qf (fields tokenise the query into different numbers of terms): all
    float32; mm applies per field (clamped to the field's term count).
  * :181-244 pf / pf2 / pf3 on the arrays SLICED to qf > 0 (df from the filtered postings, quirk
    iii); pf2 appends the last bigram's scores twice (:221, quirk vii); ps/ps2/ps3 are ignored.
  * :335-353 each phase's float32 sum is added to qf where qf != 0.
"""
import re

import numpy as np


def parse_min_should_match(num_clauses, spec):
    """solr.py:10-59 -- Solr's mm mini language: plain ints, negatives, percentages and
    `n<spec` conditionals (space separated, applied left to right)."""
    def to_int(text):
        try:
            return int(text)
        except ValueError:
            raise ValueError("Invalid 'mm' spec. Expecting an integer.")

    spec = spec.strip()
    if "<" in spec:
        spec = re.sub(r"\s*<\s*", "<", spec)
        result = num_clauses
        for clause in spec.split():
            head, sep, tail = clause.partition("<")
            if not sep:
                raise ValueError("Invalid 'mm' spec: '" + clause + "'. Expecting values before and after '<'")
            bound = to_int(head)
            if num_clauses <= bound:
                return result
            result = parse_min_should_match(num_clauses, tail)
        return result
    result = num_clauses
    if "%" in spec:
        pct = to_int(spec[:-1])
        calc = (result * pct) * (1 / 100)
        result = result + int(calc) if calc < 0 else int(calc)
    else:
        calc = to_int(spec)
        result = result + calc if calc < 0 else calc
    return min(num_clauses, max(result, 0))


def parse_field_boosts(field_lists):
    """solr.py:62-74: "title^2.5" -> {"title": 2.5}; no caret -> None."""
    out = {}
    for spec in field_lists or []:
        name, _, boost = spec.partition("^")
        out[name] = float(boost.split("^")[0]) if boost else None
    return out


class OracleField:
    """One searchable column: an OracleIndex plus the term dictionary and tokenizer."""

    def __init__(self, index, term_to_id, tokenizer=str.split, k1=1.2, b=0.75):
        self.index = index
        self.term_to_id = term_to_id
        self.tokenizer = tokenizer
        self.k1, self.b = k1, b

    def ids(self, tokens):
        return [self.term_to_id.get(t) for t in tokens]

    def score(self, tokens, index=None):
        index = self.index if index is None else index
        ids = self.ids(tokens)
        return index.score(ids[0] if len(ids) == 1 else ids, k1=self.k1, b=self.b)


def _boosted(score32, boost):
    return score32 if boost is None else score32 * np.float32(boost)


def edismax(fields, q, qf, mm=None, pf=None, pf2=None, pf3=None, tie=0.0, q_op="OR"):
    """fields: dict name -> OracleField.  Returns the score vector (float64 term-centric,
    float32 field-centric), solr.py:251-355."""
    listify = lambda x: x if isinstance(x, list) else [x]
    query_fields = parse_field_boosts(listify(qf))
    phrase_fields = parse_field_boosts(listify(pf)) if pf else {}
    bigram_fields = parse_field_boosts(pf2) if pf2 else {}
    trigram_fields = parse_field_boosts(pf3) if pf3 else {}
    mm = "1" if mm is None else (f"{mm}" if isinstance(mm, int) else mm)
    if q_op == "AND":
        mm = "100%"
    tokens = {f: list(fields[f].tokenizer(q)) for f in query_fields}
    counts = [len(t) for t in tokens.values()]
    n_terms = counts[0] if counts else 0
    term_centric = all(c == n_terms for c in counts)
    n_docs = len(next(iter(fields.values())).index)

    if term_centric:
        per_term = []
        for pos in range(n_terms):
            run_max = np.zeros(n_docs)
            run_sum = np.zeros(n_docs)
            for f, boost in query_fields.items():
                s = _boosted(fields[f].score([tokens[f][pos]]), boost)
                run_sum += s
                run_max = np.maximum(run_max, s)
            per_term.append(run_max + (run_sum - run_max) * tie)
        need = parse_min_should_match(n_terms, mm)
        stacked = np.asarray(per_term)
        enough = np.sum(stacked > 0, axis=0) >= need
        scores = np.sum(per_term, axis=0)
        scores[~enough] = 0
    else:
        per_field = []
        for f, boost in query_fields.items():
            ts = np.array([fields[f].score([t]) for t in tokens[f]])
            need = min(parse_min_should_match(len(tokens[f]), mm), len(tokens[f]))
            enough = np.sum(ts > 0, axis=0) >= need
            total = np.sum(ts, axis=0)
            total[~enough] = 0
            per_field.append(total * (1 if boost is None else boost))
        stacked = np.asarray(per_field)
        summed = np.sum(stacked, axis=0)
        best = np.max(stacked, axis=0)
        scores = best + (summed - best) * tie

    mask = scores > 0
    sliced = {f: fields[f].index.sliced(mask) for f in query_fields}

    def phase(field_boosts, gram, repeat_last=False):
        parts = []
        for f, boost in field_boosts.items():
            toks = tokens[f]
            if len(toks) < max(gram or 2, 2):
                continue
            grams = [toks] if gram is None else [toks[i:i + gram] for i in range(len(toks) - gram + 1)]
            last = None
            for g in grams:
                last = _boosted(fields[f].score(g, index=sliced[f]), boost)
                parts.append(last)
            if repeat_last:
                parts.append(last)
        return np.sum(parts, axis=0) if parts else None

    for add in (phase(phrase_fields, None), phase(bigram_fields, 2, repeat_last=True), phase(trigram_fields, 3)):
        if add is not None:
            scores[np.where(scores)[0]] += add
    return scores

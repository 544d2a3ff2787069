"""Generate golden vectors by running the REAL reference (softwaredoug/searcharray).

Run in the build container only (needs /root/reference); the GPU box never runs this.

    python tests/golden/make_golden.py

The reference is Python + Cython, so it is built in a scratch copy outside the repo
(/tmp/sa_oracle: `python setup.py build_ext --inplace`) and imported
from there.  Only inputs and outputs are written to tests/golden/*.npz|json -- no
reference source is copied.  Everything is seeded, so re-running reproduces the files.
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SCRATCH = "/tmp/sa_oracle"


def import_reference():
    if not os.path.exists(os.path.join(SCRATCH, "searcharray")):
        shutil.copytree("/root/reference", SCRATCH)
    import glob
    if not glob.glob(os.path.join(SCRATCH, "searcharray", "roaringish", "intersect*.so")):
        subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=SCRATCH)
    sys.path.insert(0, SCRATCH)
    import searcharray  # noqa
    return searcharray


U = np.uint64
HEADER_MASK = U(0xFFFFFFFFFFFC0000)


def rand_words(rng, n_docs, p_doc, max_blocks, p_block, bit_density):
    """Random header-unique sorted roaringish words: doc(28)|block(18)|bits(18)."""
    docs = np.nonzero(rng.random(n_docs) < p_doc)[0].astype(np.uint64)
    out = []
    for d in docs:
        nb = rng.integers(1, max_blocks + 1)
        blocks = np.nonzero(rng.random(nb) < p_block)[0].astype(np.uint64)
        for blk in blocks:
            bits = 0
            while bits == 0:
                bits = int(np.sum((rng.random(18) < bit_density) * (1 << np.arange(18))))
            out.append((int(d) << 36) | (int(blk) << 18) | bits)
    return np.asarray(out, dtype=np.uint64)


def gen_ops(sa, rng):
    from searcharray.roaringish import (intersect, adjacent, merge, unique, popcount64,
                                        popcount_reduce_at, key_sum_over, sort_merge_counts)
    from searcharray.roaringish.intersect import intersect_with_adjacents
    from searcharray.roaringish.popcount import popcount64_reduce
    from searcharray.roaringish.roaringish_ops import as_dense, payload_slice
    from searcharray.bm25 import bm25_score
    out = {}
    case = 0
    for (n_docs, p_doc, mb, pb, bd) in [(50, 0.5, 4, 0.7, 0.3), (400, 0.2, 6, 0.6, 0.2),
                                         (3000, 0.05, 3, 0.9, 0.5), (200, 0.9, 12, 0.8, 0.15),
                                         (10, 1.0, 30, 0.5, 0.4)]:
        for rep in range(3):
            lhs = rand_words(rng, n_docs, p_doc, mb, pb, bd)
            rhs = rand_words(rng, n_docs, p_doc * rng.uniform(0.3, 1.0), mb, pb, bd)
            if len(lhs) == 0 or len(rhs) == 0:
                continue
            k = f"c{case}_"
            out[k + "lhs"], out[k + "rhs"] = lhs, rhs
            a, b, c, d = intersect_with_adjacents(lhs, rhs, mask=HEADER_MASK)
            out[k + "iwa_li"], out[k + "iwa_ri"], out[k + "iwa_lai"], out[k + "iwa_rai"] = a, b, c, d
            a, b = intersect(lhs, rhs, mask=HEADER_MASK)
            out[k + "int_li"], out[k + "int_ri"] = a, b
            a, b = adjacent(lhs, rhs, mask=HEADER_MASK)
            out[k + "adj_li"], out[k + "adj_ri"] = a, b
            a, b = intersect(lhs >> U(36), rhs >> U(36), drop_duplicates=False)
            out[k + "keep_li"], out[k + "keep_ri"] = a, b
            a, b = intersect(lhs >> U(36), rhs >> U(36), drop_duplicates=True)
            out[k + "dropk_li"], out[k + "dropk_ri"] = a, b
            out[k + "merge"] = merge(lhs, rhs)
            out[k + "merge_drop"] = merge(lhs, rhs, drop_duplicates=True)
            out[k + "uniq_keys"] = unique(lhs, U(36))
            out[k + "uniq"] = unique(np.sort(np.concatenate([lhs, lhs[::2]])))
            ids, cnt = popcount64_reduce(lhs, U(36), U(0x3FFFF))
            out[k + "pcr_ids"], out[k + "pcr_cnt"] = ids, cnt
            out[k + "pc64"] = popcount64(lhs)
            ids2, cnt2 = popcount_reduce_at(lhs >> U(36), lhs & U(0x3FFFF) & U(0x15555))
            out[k + "pra_ids"], out[k + "pra_cnt"] = ids2, cnt2
            ids3, cnt3 = key_sum_over(lhs >> U(36), popcount64(lhs & U(0xFF)))
            out[k + "kso_ids"], out[k + "kso_cnt"] = ids3, cnt3
            rids, rcnt = popcount64_reduce(rhs, U(36), U(0x3FFFF))
            mi, mc = sort_merge_counts(ids, cnt, rids, rcnt)
            out[k + "smc_ids"], out[k + "smc_cnt"] = mi, mc
            out[k + "dense"] = as_dense(ids, cnt, n_docs)
            out[k + "pslice"] = payload_slice(lhs, U(0x0000000FFFFC0000), 1, 2)
            # bm25
            tf = as_dense(ids, cnt, n_docs)
            dl = rng.integers(0, 60, n_docs).astype(np.float32)
            out[k + "bm25_dl"] = dl
            tfc = tf.copy()
            bm25_score(tfc, dl, float(np.mean(dl)), 2.345, 1.2, 0.75)
            out[k + "bm25"] = tfc
            case += 1
    out["n_cases"] = np.asarray([case])
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **out)
    print("ops cases", case)


def gen_bigram(sa, rng):
    from searcharray.phrase.bigram_freqs import bigram_freqs, Continuation
    from searcharray.phrase.middle_out import compute_phrase_freqs
    from searcharray.phrase.spans import span_search
    out = {}
    case = 0
    cfgs = [(30, 0.8, 5, 0.8, 0.4), (300, 0.3, 3, 0.9, 0.3), (2000, 0.05, 2, 1.0, 0.5),
            (100, 0.6, 10, 0.7, 0.25), (5, 1.0, 40, 0.9, 0.5)]
    for cfg in cfgs:
        for rep in range(4):
            lhs = rand_words(rng, *cfg)
            rhs = rand_words(rng, cfg[0], cfg[1] * rng.uniform(0.2, 1.0), *cfg[2:])
            if rep == 3:
                rhs = lhs.copy()          # same-term path
            if len(lhs) == 0 or len(rhs) == 0:
                continue
            k = f"b{case}_"
            out[k + "lhs"], out[k + "rhs"] = lhs, rhs
            for cname, cont in (("R", Continuation.RHS), ("L", Continuation.LHS)):
                (ids, cnt), (ln, rn) = bigram_freqs(lhs.copy(), rhs.copy(), cont)
                out[k + cname + "_ids"] = np.asarray(ids, dtype=np.uint64)
                out[k + cname + "_cnt"] = np.asarray(cnt, dtype=np.float32)
                nxt = rn if cont == Continuation.RHS else ln
                out[k + cname + "_next"] = np.asarray(nxt, dtype=np.uint64)
            case += 1
    out["n_bigram"] = np.asarray([case])
    # multi-term phrases (lists shuffled so all three drivers are exercised)
    pcase = 0
    for n_terms in (2, 3, 4, 4, 5, 6, 7):
        for rep in range(4):
            cfg = cfgs[(pcase + rep) % len(cfgs)]
            enc = [rand_words(rng, cfg[0], cfg[1] * rng.uniform(0.15, 1.0), *cfg[2:]) for _ in range(n_terms)]
            if rep == 2 and n_terms >= 3:
                enc[1] = enc[0].copy()
            if rep == 3 and n_terms >= 3:
                enc[-1] = enc[-2].copy()
            if any(len(e) == 0 for e in enc):
                continue
            k = f"p{pcase}_"
            out[k + "n"] = np.asarray([n_terms])
            for i, e in enumerate(enc):
                out[k + f"t{i}"] = e
            ids, cnt = compute_phrase_freqs([e.copy() for e in enc])
            out[k + "ids"] = np.asarray(ids, dtype=np.uint64)
            out[k + "cnt"] = np.asarray(cnt, dtype=np.float32)
            for slop in (1, 2, 4):
                if n_terms <= 5:
                    sids, scnt = span_search([e.copy() for e in enc], slop)
                    out[k + f"s{slop}_ids"] = sids
                    out[k + f"s{slop}_cnt"] = scnt
            pcase += 1
    out["n_phrase"] = np.asarray([pcase])
    np.savez_compressed(os.path.join(HERE, "bigram.npz"), **out)
    print("bigram cases", case, "phrase cases", pcase)


VOCAB = ["w%d" % i for i in range(120)]


def make_corpus(rng, n_docs):
    """Zipf-ish random text; some empty docs, some long ones (cross 18-position blocks)."""
    probs = 1.0 / np.arange(1, len(VOCAB) + 1) ** 1.1
    probs /= probs.sum()
    docs = []
    for i in range(n_docs):
        r = rng.random()
        if r < 0.03:
            docs.append("")
            continue
        n = int(rng.integers(1, 12)) if r < 0.5 else int(rng.integers(12, 70))
        if r > 0.97:
            n = int(rng.integers(150, 420))
        toks = list(rng.choice(VOCAB, size=n, p=probs))
        # plant some phrases
        if rng.random() < 0.3 and n > 6:
            at = int(rng.integers(0, n - 4))
            toks[at:at + 4] = ["w3", "w7", "w1", "w9"]
        if rng.random() < 0.2 and n > 6:
            at = int(rng.integers(0, n - 3))
            toks[at:at + 3] = ["w2", "w2", "w2"]
        docs.append(" ".join(toks))
    return docs


def gen_api(sa, rng):
    from searcharray import SearchArray
    from searcharray.similarity import bm25_similarity
    docs = make_corpus(rng, 1500)
    arr = SearchArray.index(docs, autowarm=False)
    out = {}
    meta = {"docs": docs, "queries": []}
    # index dump (to pin the host indexer): per-term words
    terms = sorted(arr.term_dict.term_to_ids.keys(), key=lambda t: arr.term_dict.term_to_ids[t])
    meta["terms"] = terms
    lens = []
    allw = []
    for t in terms:
        w = arr.posns.encoded_term_posns[arr.term_dict.get_term_id(t)]
        lens.append(len(w))
        allw.append(np.asarray(w, dtype=np.uint64))
    out["index_words"] = np.concatenate(allw)
    out["index_lens"] = np.asarray(lens, dtype=np.uint64)
    out["doc_lens"] = np.asarray(arr.doc_lens, dtype=np.float32)
    out["avg_doc_length"] = np.asarray([arr.avg_doc_length], dtype=np.float32)

    queries = [
        ("w0",), ("w1",), ("w5",), ("w50",), ("w119",), ("nope",),
        ("w3", "w7"), ("w7", "w1"), ("w0", "w1"), ("w1", "w0"), ("w0", "w0"), ("w2", "w2"),
        ("w2", "w2", "w2"), ("w3", "w7", "w1"), ("w3", "w7", "w1", "w9"), ("w9", "w1", "w7", "w3"),
        ("w0", "w3", "w7", "w1"), ("w3", "w7", "w1", "w0"), ("w3", "nope"), ("w0", "w1", "w2", "w3", "w4"),
        ("w0", "w1", "w3", "w7", "w1", "w9"), ("w3", "w7", "w1", "w9", "w0", "w1"),
        ("w0", "w0", "w3", "w7", "w1", "w9", "w0"), ("w1", "w2", "w2", "w2", "w0"),
    ]
    sims = {"default": None, "k1b": bm25_similarity(k1=0.9, b=0.4)}
    qi = 0
    for q in queries:
        tok = q[0] if len(q) == 1 else list(q)
        rec = {"tokens": list(q), "idx": qi}
        arr.posns.clear_cache()
        out[f"q{qi}_tf"] = arr.termfreqs(tok)
        out[f"q{qi}_score"] = arr.score(tok)
        out[f"q{qi}_score_k1b"] = arr.score(tok, similarity=sims["k1b"])
        if len(q) == 1:
            out[f"q{qi}_df"] = np.asarray([arr.docfreq(q[0])], dtype=np.uint64)
            out[f"q{qi}_tf_max17"] = arr.termfreqs(tok, max_posn=17)
            out[f"q{qi}_tf_min18"] = arr.termfreqs(tok, min_posn=18)
        else:
            out[f"q{qi}_tf_max17"] = arr.termfreqs(tok, max_posn=17)
            out[f"q{qi}_tf_min18"] = arr.termfreqs(tok, min_posn=18)
            if len(q) <= 5:
                for slop in (1, 2, 3):
                    out[f"q{qi}_tf_slop{slop}"] = arr.termfreqs(tok, slop=slop)
                out[f"q{qi}_score_slop2"] = arr.score(tok, slop=2)
        # sliced arrays (FilteredPosns semantics, quirk iii)
        sl = arr[1::2]
        out[f"q{qi}_tf_odd"] = sl.termfreqs(tok)
        out[f"q{qi}_score_odd"] = sl.score(tok)
        sl2 = arr[100:700]
        out[f"q{qi}_score_mid"] = sl2.score(tok)
        arr.posns.clear_cache()
        meta["queries"].append(rec)
        qi += 1
    np.savez_compressed(os.path.join(HERE, "api.npz"), **out)
    with open(os.path.join(HERE, "api.json"), "w") as f:
        json.dump(meta, f)
    print("api queries", qi, "docs", len(docs), "terms", len(terms))


def main():
    sa = import_reference()
    rng = np.random.default_rng(20260924)
    gen_ops(sa, rng)
    gen_bigram(sa, rng)
    gen_api(sa, rng)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()

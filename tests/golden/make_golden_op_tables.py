"""Known-answer tables of the reference's op tests, extracted by importing the
reference's test modules: test/test_snp_ops.py:96-154 (intersect), :457-522 (adjacent), :537-548
(merge), :384-396 (unique), test/test_bitcount64.py:9-34, plus the reference's OUTPUT on its seven
saved posting pairs fixtures/{lhs,rhs,mask}_*.npy (test_snp_ops.py:324-350; the arrays themselves stay
in the reference tree -- only lengths and SHA-256 digests of the outputs are recorded here).

    python tests/golden/make_golden_op_tables.py      (build container only)
"""
import hashlib
import json
import os
import sys

import numpy as np

from make_golden import import_reference, HERE, SCRATCH

ALL = np.uint64(0xFFFFFFFFFFFFFFFF)


def ints(a):
    return [int(x) for x in np.asarray(a).ravel()]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def main():
    import_reference()
    sys.path.insert(0, os.path.join(SCRATCH, "test"))
    import types
    # dev-only dependencies of the reference's test module that this image lacks
    if "sortednp" not in sys.modules:
        snp = types.ModuleType("sortednp")
        snp.intersect = lambda a, b, **kw: np.intersect1d(a, b)
        snp.merge = lambda a, b, **kw: np.sort(np.concatenate([a, b]))
        sys.modules["sortednp"] = snp
    import test_snp_ops as tso
    import test_bitcount64 as tbc
    from searcharray.roaringish import intersect, adjacent, merge, unique, popcount64
    from searcharray.roaringish.intersect import intersect_with_adjacents
    out = {"intersect": [], "adjacent": [], "merge": [], "bitcount": [], "unique": [], "fixtures": []}
    for name, sc in tso.intersect_scenarios.items():
        mask = ALL if sc["mask"] is None else np.uint64(sc["mask"])
        lhs, rhs = sc["lhs"], sc["rhs"]
        li, ri = intersect(lhs, rhs, mask=mask)
        assert np.all((lhs[li] & mask) == sc["expected"]), name
        lk, rk = intersect(lhs, rhs, mask=mask, drop_duplicates=False)
        a0, a1, a2, a3 = intersect_with_adjacents(lhs, rhs, mask=mask)
        out["intersect"].append({"name": name, "lhs": ints(lhs), "rhs": ints(rhs), "mask": int(mask),
                                 "expected": ints(sc["expected"]), "lhs_idx": ints(li), "rhs_idx": ints(ri),
                                 "keep_lhs_idx": ints(lk), "keep_rhs_idx": ints(rk),
                                 "with_adjacents": [ints(a0), ints(a1), ints(a2), ints(a3)]})
    for name, sc in tso.adj_scenarios.items():
        mask = ALL if sc["mask"] is None else np.uint64(sc["mask"])
        if sc["delta"] == -1:
            ri, li = adjacent(sc["rhs"], sc["lhs"], mask)
        else:
            li, ri = adjacent(sc["lhs"], sc["rhs"], mask)
        if "lhs_idx_expected" in sc:
            assert np.all(li == sc["lhs_idx_expected"]) and np.all(ri == sc["rhs_idx_expected"]), name
        out["adjacent"].append({"name": name, "lhs": ints(sc["lhs"]), "rhs": ints(sc["rhs"]), "mask": int(mask),
                                "delta": sc["delta"], "lhs_idx": ints(li), "rhs_idx": ints(ri)})
    for name, sc in tso.merge_scenarios.items():
        m0 = merge(sc["lhs"], sc["rhs"])
        assert np.all(m0 == sc["expected"]), name
        out["merge"].append({"name": name, "lhs": ints(sc["lhs"]), "rhs": ints(sc["rhs"]), "merged": ints(m0),
                             "merged_dropdup": ints(merge(sc["lhs"], sc["rhs"], drop_duplicates=True))})
    for name, sc in tbc.scenarios.items():
        assert list(popcount64(sc["bits"].copy())) == sc["expected"], name
        out["bitcount"].append({"name": name, "bits": ints(sc["bits"]), "expected": ints(sc["expected"])})
    for arr, shift in ((np.asarray([0, 0, 11, 11, 11, 36, 41, 42], dtype=np.uint64), 0),
                       (np.asarray([0xEE00, 0xFF00, 0xFF01], dtype=np.uint64), 8)):
        out["unique"].append({"arr": ints(arr), "shift": shift, "expected": ints(unique(arr, shift))})
    for suffix in (128, 185, 24179, 27685, 44358, 45907, 90596):
        lhs = np.load(f"/root/reference/fixtures/lhs_{suffix}.npy")
        rhs = np.load(f"/root/reference/fixtures/rhs_{suffix}.npy")
        mask = np.load(f"/root/reference/fixtures/mask_{suffix}.npy")
        li, ri = intersect(lhs, rhs, mask=mask)
        a = intersect_with_adjacents(lhs, rhs, mask=mask)
        out["fixtures"].append({"suffix": suffix, "n_lhs": len(lhs), "n_rhs": len(rhs), "mask": int(mask),
                                "intersect": [len(li), digest(li), digest(ri)],
                                "with_adjacents": [[len(x), digest(x)] for x in a]})
        print(suffix, len(lhs), len(rhs), hex(int(mask)), len(li), [len(x) for x in a])
    with open(os.path.join(HERE, "op_tables.json"), "w") as f:
        json.dump(out, f)
    print({k: len(v) for k, v in out.items()}, os.path.getsize(os.path.join(HERE, "op_tables.json")))


if __name__ == "__main__":
    main()

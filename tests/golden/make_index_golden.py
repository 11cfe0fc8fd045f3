"""Generates tests/golden/index_golden.json from the CPU index oracle (oracle/index_oracle.c), which is pinned to the
properties index_test.go holds (tests/test_index_oracle.py; the reference has no numeric vectors for this path and, being
Go, cannot run in this image).  Restatement-derived vectors: a regression pin for the oracle (CPU suite) and a fixed
target for the CUDA path (GPU suite).
    python tests/golden/make_index_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import index_oracle as I  # noqa: E402


def inputs(seed, dim, n, nq):
    """the vectors are not stored: both sides regenerate them from the seed (PCG64 streams are stable across NumPy)"""
    rng = np.random.default_rng(seed)
    return np.round(rng.random((dim, n)), 6), np.round(rng.random((dim, nq)), 6)


CASES = [("CosineDistance", 7, 24, 500, 3, 5), ("CosineSimilarity", 8, 64, 300, 2, 4), ("AngularDistance", 9, 16, 200, 2, 3),
         ("EuclideanDistance", 10, 10, 400, 2, 6), ("ManhattenDistance", 11, 33, 250, 2, 4), ("VectorLenSimilarity", 12, 8, 100, 2, 3)]

if __name__ == "__main__":
    out = {"about": "oracle-derived (float64 C restatement of index.go / comparisons.go); matches sorted nearest first", "cases": []}
    for name, seed, dim, n, nq, k in CASES:
        m, q = inputs(seed, dim, n, nq)
        index = I.LinearScanIndex(name)
        for j in range(n):
            index.Index(m[:, j], 100 + j)
        res = []
        for i in range(nq):
            ms = sorted(index.Search(q[:, i], k))
            res.append({"ids": [x.ID for x in ms], "distances": [x.Distance for x in ms]})
        out["cases"].append({"comparer": name, "seed": seed, "dim": dim, "n": n, "nq": nq, "k": k, "checksum": float(m.sum() + q.sum()),
                             "results": res})
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "index_golden.json"), "w"), indent=1)
    print("wrote", len(out["cases"]), "cases")

"""CPU checks that pin the index oracle (oracle/index_oracle.c) to what the reference holds for this path: the
properties of index_test.go (no numeric golden vectors exist there) and closed-form values of the pairwise comparers
(measures/pairwise/comparisons.go)."""
import math

import numpy as np
import pytest

from oracle import index_oracle as I


def random_dense(rows, cols, seed):
    # sparse.Random(sparse.DenseFormat, 100, 10, 1.0) in index_test.go:14,48,99: a fully dense random matrix
    return np.random.default_rng(seed).random((rows, cols))


def test_indexer_index():
    # index_test.go:13-44 (LinearScanIndex row): every indexed vector finds itself at distance 0 +- 1e-7
    m = random_dense(100, 10, 1)
    index = I.LinearScanIndex("CosineDistance")
    for j in range(10):
        index.Index(m[:, j], j)
    for j in range(10):
        matches = index.Search(m[:, j], 1)
        assert len(matches) == 1 and matches[0].ID == j
        assert -1e-7 <= matches[0].Distance <= 1e-7


def test_indexer_search():
    # index_test.go:46-96: Search(v, numCols), sorted nearest first, is the argsort of the pairwise distances
    num = 10
    m = random_dense(100, num, 2)
    index = I.LinearScanIndex("CosineDistance")
    for j in range(num):
        index.Index(m[:, j], j)
    for j in range(num):
        d = [I.compare("CosineDistance", m[:, j], m[:, i]) for i in range(num)]
        matches = sorted(index.Search(m[:, j], num))
        assert len(matches) == num
        assert [x.ID for x in matches] == np.argsort(d, kind="stable").tolist()


def test_indexer_remove():
    # index_test.go:98-135
    m = random_dense(100, 10, 3)
    index = I.LinearScanIndex("CosineDistance")
    for j in range(10):
        index.Index(m[:, j], j)
    for j in range(10):
        index.Remove(j)
        matches = index.Search(m[:, j], 1)
        assert len(matches) <= 1
        if matches:
            assert matches[0].ID != j


def test_search_returns_fewer_than_k_and_keeps_the_k_nearest():
    rng = np.random.default_rng(4)
    m = rng.random((16, 200))
    index = I.LinearScanIndex("EuclideanDistance")
    for j in range(200):
        index.Index(m[:, j], 1000 + j)
    q = rng.random(16)
    assert len(index.Search(q, 500)) == 200          # index.go:101-103
    got = sorted(index.Search(q, 7))
    want = np.argsort(np.sqrt(((m - q[:, None]) ** 2).sum(0)))[:7]
    assert [x.ID - 1000 for x in got] == want.tolist()


def test_ties_at_the_cut_keep_the_later_vectors():
    # index.go:108 `dist <= heap top`: a later vector at the same distance replaces an earlier one
    index = I.LinearScanIndex("EuclideanDistance")
    for j in range(6):
        index.Index([1.0, 0.0], j)
    assert sorted(x.ID for x in index.Search([0.0, 0.0], 2)) == [4, 5]


def test_comparers_closed_form():
    a, b = np.array([1.0, 2.0, 2.0]), np.array([2.0, 0.0, 1.0])
    cos = 4.0 / (3.0 * math.sqrt(5.0))
    assert I.compare("CosineSimilarity", a, b) == pytest.approx(cos, abs=1e-15)
    assert I.compare("CosineDistance", a, b) == pytest.approx(1 - cos, abs=1e-15)
    assert I.compare("AngularDistance", a, b) == pytest.approx(math.acos(cos) / math.pi, abs=1e-15)
    assert I.compare("AngularSimilarity", a, b) == pytest.approx(1 - math.acos(cos) / math.pi, abs=1e-15)
    assert I.compare("HammingDistance", a, b) == 1.0 and I.compare("HammingDistance", a, a) == 0.0
    assert I.compare("HammingSimilarity", [1, 2, 3, 4], [1, 0, 3, 0]) == 0.5
    assert I.compare("EuclideanDistance", a, b) == pytest.approx(math.sqrt(1 + 4 + 1), abs=1e-15)
    assert I.compare("ManhattenDistance", a, b) == 4.0
    assert I.compare("VectorLenSimilarity", a, b) == 2.0
    # zero vectors: NaN (comparisons.go:24-26, 113-115); cos > 1 by rounding is clamped (comparisons.go:50-52)
    z = np.zeros(3)
    for name in ("CosineSimilarity", "CosineDistance", "AngularDistance", "VectorLenSimilarity"):
        assert math.isnan(I.compare(name, a, z))
    assert I.compare("AngularDistance", a, a) == pytest.approx(0.0, abs=2e-8)


def test_committed_golden_vectors_reproduce():
    # tests/golden/index_golden.json (tests/golden/make_index_golden.py): the oracle and the input generator are stable
    import json
    import os
    from tests.golden.make_index_golden import inputs
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "index_golden.json")))
    assert len(g["cases"]) >= 6
    for case in g["cases"]:
        m, q = inputs(case["seed"], case["dim"], case["n"], case["nq"])
        assert float(m.sum() + q.sum()) == case["checksum"]
        index = I.LinearScanIndex(case["comparer"])
        for j in range(case["n"]):
            index.Index(m[:, j], 100 + j)
        for i, want in enumerate(case["results"]):
            got = sorted(index.Search(q[:, i], case["k"]))
            assert [x.ID for x in got] == want["ids"]
            assert np.allclose([x.Distance for x in got], want["distances"], rtol=0, atol=1e-15)

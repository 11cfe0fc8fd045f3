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


def _go_linear_scan(dist, ids, k):
    """index.go:85-115 with Go's container/heap written out in Python (heap.Init / Pop / Push via up / down and
    resultHeap.Less = larger distance first): an independent restatement that the C oracle must agree with match for
    match, in heap order, ties included"""
    less = lambda h, i, j: h[i][0] > h[j][0]

    def up(h, j):
        while True:
            i = int((j - 1) / 2)   # Go's integer division truncates toward zero
            if i == j or not less(h, j, i):
                break
            h[i], h[j] = h[j], h[i]
            j = i

    def down(h, i, n):
        while True:
            j1 = 2 * i + 1
            if j1 >= n or j1 < 0:
                break
            j = j1
            if j1 + 1 < n and less(h, j1 + 1, j1):
                j = j1 + 1
            if not less(h, j, i):
                break
            h[i], h[j] = h[j], h[i]
            i = j

    h = [(dist[p], ids[p]) for p in range(min(k, len(dist)))]
    if len(h) < k:
        return h
    for i in range(len(h) // 2 - 1, -1, -1):
        down(h, i, len(h))
    for p in range(k, len(dist)):
        if dist[p] <= h[0][0]:
            n = len(h) - 1
            h[0], h[n] = h[n], h[0]
            down(h, 0, n)
            h[n] = (dist[p], ids[p])
            up(h, n)
    return h


@pytest.mark.parametrize("seed", range(6))
def test_c_oracle_agrees_with_a_python_restatement_of_the_heap(seed):
    rng = np.random.default_rng(seed)
    n, dim, k = int(rng.integers(1, 120)), 4, int(rng.integers(1, 40))
    m = rng.integers(0, 3, (dim, n)).astype(float)   # few distinct values: many exactly equal distances
    q = rng.integers(0, 3, dim).astype(float)
    index = I.LinearScanIndex("ManhattenDistance")
    for j in range(n):
        index.Index(m[:, j], 10 * j)
    dist = [I.compare("ManhattenDistance", q, m[:, j]) for j in range(n)]
    want = _go_linear_scan(dist, [10 * j for j in range(n)], k)
    got = index.Search(q, k)
    assert [(x.Distance, x.ID) for x in got] == want

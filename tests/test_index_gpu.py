"""CUDA tests of the exact nearest-neighbour scan (run on the B200 box with -m gpu): the reference's own index tests
(index_test.go, LinearScanIndex rows) against the CUDA path through the C ABI, and parity with the CPU oracle
(oracle/index_oracle.c) on the same seeded inputs.

Stated tolerance: comparer values within 1e-12 absolute of the float64 oracle (both sum in index order; the device
uses fused multiply-adds); ids identical whenever no two candidates are closer than that at the cut."""
import math

import numpy as np
import pytest

import nlp_b200
from nlp_b200 import _lib
from nlp_b200.measures import pairwise
from oracle import index_oracle as I

pytestmark = pytest.mark.gpu

DIST_ATOL = 1e-12


def random_dense(rows, cols, seed):
    return np.random.default_rng(seed).random((rows, cols))   # sparse.Random(sparse.DenseFormat, r, c, 1.0)


def col_do(m, fn):
    for j in range(m.shape[1]):   # utils.go ColDo
        fn(j, m[:, j])


# ---- index_test.go against the CUDA path ---------------------------------------------------------------------------
def test_indexer_index():
    m = random_dense(100, 10, 1)
    index = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    col_do(m, lambda j, v: index.Index(v, j))
    for j in range(10):
        matches = index.Search(m[:, j], 1)
        assert len(matches) == 1                                # index_test.go:33
        assert matches[0].ID == j                               # :36
        assert -1e-7 <= matches[0].Distance <= 1e-7             # :39


def test_indexer_search():
    num = 10
    m = random_dense(100, num, 2)
    index = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    col_do(m, lambda j, v: index.Index(v, j))
    for j in range(num):
        d = [I.compare("CosineDistance", m[:, j], m[:, i]) for i in range(num)]
        matches = index.Search(m[:, j], num)
        assert len(matches) == num                              # index_test.go:80
        assert [x.ID for x in matches] == np.argsort(d, kind="stable").tolist()   # :86-91


def test_indexer_remove():
    m = random_dense(100, 10, 3)
    index = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    col_do(m, lambda j, v: index.Index(v, j))
    for j in range(10):
        index.Remove(j)
        matches = index.Search(m[:, j], 1)
        assert len(matches) <= 1                                # index_test.go:121
        if matches:
            assert matches[0].ID != j                           # :125
    assert len(index) == 0 and index.Search(m[:, 0], 3) == []
    index.Remove(12345)                                         # unknown id: nothing happens (index.go:117-119)


# ---- parity with the oracle -----------------------------------------------------------------------------------------
def _check_against_oracle(comparer, m, queries, k, ids=None):
    dim, n = m.shape
    ids = list(range(n)) if ids is None else ids
    index = nlp_b200.NewLinearScanIndex(comparer)
    index.IndexColumns(m, ids)
    o = I.LinearScanIndex(comparer.name)
    for j in range(n):
        o.Index(m[:, j], ids[j])
    got_all = index.SearchColumns(queries, k)
    for qi in range(queries.shape[1]):
        want = sorted(o.Search(queries[:, qi], k), key=lambda x: (x.Distance if not math.isnan(x.Distance) else math.inf))
        got = got_all[qi]
        assert len(got) == len(want) == min(k, n)
        gd, wd = np.array([x.Distance for x in got]), np.array([x.Distance for x in want])
        finite = ~np.isnan(wd)
        assert np.array_equal(np.isnan(gd), np.isnan(wd)), comparer
        assert np.abs(gd[finite] - wd[finite]).max(initial=0) <= DIST_ATOL, comparer
        assert np.all(np.diff(gd[finite]) >= 0)                 # nearest first
        # same ids wherever the oracle's neighbouring distances are further apart than the tolerance
        full = np.array([I.compare(comparer.name, queries[:, qi], m[:, j]) for j in range(n)])
        order = np.sort(full[~np.isnan(full)])
        clear = len(order) <= k or order[k] - order[k - 1] > 10 * DIST_ATOL
        if clear and finite.all():
            assert sorted(x.ID for x in got) == sorted(x.ID for x in want), comparer
    index.Close()
    return got_all


@pytest.mark.parametrize("dim", [37, 64])   # 64: the dot-product comparers take the FP64 tensor-core kernel (dim % 4 == 0)
@pytest.mark.parametrize("comparer", pairwise.ALL, ids=lambda c: c.name)
def test_every_comparer_matches_the_oracle(comparer, dim):
    rng = np.random.default_rng(5)
    m = rng.random((dim, 300))
    if comparer.name.startswith("Hamming"):
        m = rng.integers(0, 3, (dim, 300)).astype(float)
    q = m[:, rng.integers(0, 300, 9)] if comparer.name.startswith("Hamming") else rng.random((dim, 9))
    q[:, 0] = m[:, 17]
    _check_against_oracle(comparer, m, q, 11)
    _check_against_oracle(comparer, m, q[:, :3], 4)     # few queries: the plain kernel


@pytest.mark.parametrize("dim,n,nq,k", [(1, 50, 3, 5), (3, 5000, 1, 1), (256, 3000, 16, 10), (256, 3000, 17, 64),
                                         (1000, 700, 5, 700), (600, 900, 2, 2000), (64, 40000, 4, 100), (5, 1, 2, 3)])
def test_shapes(dim, n, nq, k):
    rng = np.random.default_rng(dim * 131 + n)
    _check_against_oracle(pairwise.CosineDistance, rng.random((dim, n)), rng.random((dim, nq)), k,
                          ids=[7 * j + 3 for j in range(n)])


def test_theta_columns_as_returned_by_the_lda_path():
    """the K x D doc-topic matrix goes into the index as it is (columns sum to 1); every document is its own nearest
    neighbour and the neighbours agree with the oracle"""
    rng = np.random.default_rng(6)
    theta = rng.dirichlet(np.full(32, 0.2), size=2000).T.copy()          # K x D
    got = _check_against_oracle(pairwise.CosineDistance, theta, theta[:, :25], 8)
    assert [g[0].ID for g in got] == list(range(25))
    assert max(abs(g[0].Distance) for g in got) <= 1e-12


def test_ties_and_duplicates():
    # equal distances: the later-indexed vector ranks first and survives at the cut (index.go:108)
    index = nlp_b200.NewLinearScanIndex(pairwise.EuclideanDistance)
    for j in range(6):
        index.Index([1.0, 0.0], j)
    assert [x.ID for x in index.Search([0.0, 0.0], 2)] == [5, 4]
    assert [x.ID for x in index.Search([0.0, 0.0], 6)] == [5, 4, 3, 2, 1, 0]
    # many duplicates around the cut, mixed with distinct vectors
    rng = np.random.default_rng(7)
    m = rng.random((8, 5000))
    m[:, 1000:3000] = m[:, [999]]
    index2 = nlp_b200.NewLinearScanIndex(pairwise.ManhattenDistance)
    index2.IndexColumns(m)
    got = index2.Search(m[:, 999], 50)
    assert [x.ID for x in got] == list(range(2999, 2949, -1)) and all(x.Distance == 0 for x in got)
    # duplicate ids: Remove drops the first vector indexed under the id (index.go:121-122)
    index.Index([5.0, 5.0], 3)
    index.Remove(3)
    assert sorted(x.ID for x in index.Search([0.0, 0.0], 10)) == [0, 1, 2, 3, 4, 5]
    assert index.Search([5.0, 5.0], 1)[0] == nlp_b200.Match(0.0, 3)


def test_zero_vectors_rank_last_and_arbitrary_ids():
    index = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    index.Index([0.0, 0.0, 0.0], "zero")
    index.Index([1.0, 0.0, 0.0], ("tuple", 1))
    index.Index([0.0, 2.0, 0.0], "y")
    got = index.Search([1.0, 1.0, 0.0], 3)
    assert [x.ID for x in got[:2]] in (["y", ("tuple", 1)], [("tuple", 1), "y"]) and got[2].ID == "zero"
    assert math.isnan(got[2].Distance) and abs(got[0].Distance - (1 - 1 / math.sqrt(2))) <= 1e-12
    assert math.isnan(pairwise.CosineSimilarity([0, 0], [1, 2]))
    assert pairwise.EuclideanDistance([0, 0], [3, 4]) == 5.0


def test_growth_keeps_earlier_vectors():
    rng = np.random.default_rng(8)
    m = rng.random((12, 5000))
    index = nlp_b200.NewLinearScanIndex(pairwise.EuclideanDistance)
    for lo in range(0, 5000, 700):       # several capacity doublings
        index.IndexColumns(m[:, lo:lo + 700])
    assert len(index) == 5000
    ids, dist, found = index.SearchRaw(m[:, ::500], 1)
    assert ids[:, 0].tolist() == list(range(0, 5000, 500)) and np.all(dist[:, 0] == 0) and np.all(found == 1)
    for j in (0, 2500, 4999):
        index.Remove(j)
    ids, _, _ = index.SearchRaw(m[:, [0, 1, 2500, 4999, 4998]], 1)
    assert ids[1, 0] == 1 and ids[4, 0] == 4998 and ids[0, 0] != 0 and ids[2, 0] != 2500 and ids[3, 0] != 4999
    assert len(index) == 4997


def test_errors():
    index = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    index.Index([1.0, 2.0], 0)
    with pytest.raises(_lib.LdaB200Error, match="elements"):
        index.Index([1.0, 2.0, 3.0], 1)
    with pytest.raises(ValueError):
        index.Search([1.0, 2.0, 3.0], 1)
    with pytest.raises(_lib.LdaB200Error, match="k must be"):
        index.Search([1.0, 2.0], 0)
    with pytest.raises(TypeError):
        nlp_b200.NewLinearScanIndex(lambda a, b: 0.0)
    assert index.Search([1.0, 2.0], 1)[0].ID == 0   # still usable


def test_committed_golden_vectors():
    """tests/golden/index_golden.json: fixed targets for the CUDA path (generated by tests/golden/make_index_golden.py)"""
    import json
    import os
    from tests.golden.make_index_golden import inputs
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "index_golden.json")))
    by_name = {c.name: c for c in pairwise.ALL}
    for case in g["cases"]:
        m, q = inputs(case["seed"], case["dim"], case["n"], case["nq"])
        assert float(m.sum() + q.sum()) == case["checksum"]
        index = nlp_b200.NewLinearScanIndex(by_name[case["comparer"]])
        index.IndexColumns(m, [100 + j for j in range(case["n"])])
        got = index.SearchColumns(q, case["k"])
        for i, want in enumerate(case["results"]):
            assert [x.ID for x in got[i]] == want["ids"], case["comparer"]
            assert np.abs(np.array([x.Distance for x in got[i]]) - np.array(want["distances"])).max() <= DIST_ATOL
        index.Close()


def test_transform_into_index_keeps_the_vectors_on_the_device():
    """lda_b200_transform_into_index: ColDo(lda.Transform(m), index.Index) fused -- the Transform pipeline writes its
    result straight into the index's storage.  Same vectors, hence bitwise the same search results, as the round trip
    through the host; several pipeline ranges, custom ids, and more vectors appended afterwards."""
    import corpus_synth
    from nlp_b200 import rand
    from nlp_b200.lda import CSC
    W, D, K = 3000, 70_000, 16
    fit = CSC(W, 400, *corpus_synth.generate(400, W, block=100, mean_distinct=30, n_topics=16))
    held = CSC(W, D, *corpus_synth.generate(D, W, seed=77, block=8192, mean_distinct=8, n_topics=16))
    lda = nlp_b200.NewLatentDirichletAllocation(K)
    lda.Rnd = rand.New(rand.NewSource(0))
    lda.Processes, lda.Iterations = 1, 5
    lda.Fit(fit)
    t0 = np.random.default_rng(3).random((D, K))
    theta = lda.Transform(held, theta0=t0)
    ids = [("doc", j) for j in range(D)]
    via_host = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    via_host.IndexColumns(theta, ids)
    fused = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    fused.IndexTransform(lda, held, ids=ids, theta0=t0)
    assert len(fused) == len(via_host) == D
    q = theta[:, [0, 1, 65_535, 65_536, D - 1, 12_345]]
    a, b = via_host.SearchRaw(q, 7), fused.SearchRaw(q, 7)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert [[m.ID for m in r] for r in via_host.SearchColumns(q, 7)] == [[m.ID for m in r] for r in fused.SearchColumns(q, 7)]
    assert fused.Search(theta[:, 65_536], 1)[0].ID == ("doc", 65_536)
    # the index keeps growing after a fused append, and a second fused append lands behind the first
    fused.Index(theta[:, 5] * 3.0, "scaled copy of 5")
    assert {m.ID for m in fused.Search(theta[:, 5], 2)} == {("doc", 5), "scaled copy of 5"}
    fused.IndexTransform(lda, held, ids=[("again", j) for j in range(D)], theta0=t0)
    assert len(fused) == 2 * D + 1
    assert {m.ID for m in fused.Search(theta[:, 40_000], 2)} == {("doc", 40_000), ("again", 40_000)}
    # a model / index dimension mismatch is an error, not a crash
    other = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    other.Index(np.ones(K + 1), 0)
    with pytest.raises(_lib.LdaB200Error, match="elements"):
        other.IndexTransform(lda, held, theta0=t0)
    for x in (via_host, fused, other):
        x.Close()
    lda.Close()

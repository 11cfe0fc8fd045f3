"""Pins the CPU oracle (oracle/lda_oracle.c) against every golden value the reference's
own tests hold for the LDA path (lda_test.go:16-235) and against the known-answer values
of SURVEY.md Appendix B.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.fixtures import EXPECTED_DOCS, FIT_CASES, FIXTURE_A, FIXTURE_B


def test_pcg_known_answers():
    # SURVEY.md Appendix B (x/exp/rand PCG XSL-RR 128/64, Seed: low = high = seed)
    p = O.PCG(0)
    assert [p.uint64() for _ in range(4)] == [0xCBF98931523D4EEF, 0x4D98B91B8D356870, 0x01070196E695F8F1,
                                              0x703EC840C59F4493]
    p = O.PCG(1)
    assert [p.uint64() for _ in range(2)] == [0x34E936394905D167, 0x817C0EF62FE4C731]
    p = O.PCG(0)
    assert [p.int() % 27 for _ in range(12)] == [1, 5, 1, 7, 2, 24, 8, 18, 4, 14, 10, 12]


@pytest.mark.parametrize("data,expected", FIT_CASES)
def test_lda_fit(data, expected):
    # TestLDAFit, lda_test.go:65-88
    lda = O.LatentDirichletAllocation(3, seed=0)
    lda.Fit(data)
    components = lda.Components()
    assert components.shape == (3, 9)
    assert np.abs(components - expected).max() <= 0.01
    assert np.abs(1 - components.sum(axis=1)).max() <= 1e-8


@pytest.mark.parametrize("data", [FIXTURE_A, FIXTURE_B])
def test_lda_fit_transform(data):
    # TestLDAFitTransform, lda_test.go:152-176: theta.At(topic, doc)
    lda = O.LatentDirichletAllocation(3, seed=0)
    theta = lda.FitTransform(data)
    assert theta.shape == (3, 9)
    assert np.abs(theta.T - EXPECTED_DOCS).max() <= 0.01
    assert np.abs(1 - theta.sum(axis=0)).max() <= 1e-8


@pytest.mark.parametrize("data", [FIXTURE_A, FIXTURE_B])
def test_lda_transform(data):
    # TestLDATransform, lda_test.go:217-234: mat.EqualApprox(theta, tTheta, 0.035)
    lda = O.LatentDirichletAllocation(3, seed=0)
    lda.PerplexityEvaluationFrequency = 2
    theta = lda.FitTransform(data)
    t_theta = lda.Transform(data)
    assert np.abs(theta - t_theta).max() <= 0.035


def test_appendix_b_known_answers():
    lda = O.LatentDirichletAllocation(3, seed=0)
    lda.Fit(FIXTURE_A)
    assert lda.iterationsRun == 180
    np.testing.assert_allclose(lda.perplexity_trace[3:],
                               [3.0410271022125888, 3.0229426294727886, 3.0132846416856203], rtol=1e-12)
    lda = O.LatentDirichletAllocation(3, seed=0)
    lda.PerplexityEvaluationFrequency = 1
    lda.PerplexityTolerance = 0.0
    lda.Iterations = 5
    theta = lda.FitTransform(FIXTURE_A)
    np.testing.assert_allclose(lda.perplexity_trace, [8.898717368731296, 8.12220922656947, 7.332987602684897,
                                                      6.729362017028571, 6.323670338755834], rtol=1e-12)
    np.testing.assert_allclose(theta[:, 0], [0.9466819569492324, 0.04010285059873779, 0.013215192452029862],
                               rtol=1e-12)
    np.testing.assert_allclose(lda.get_state()[1], [10.585668278116637, 13.815478338657178, 13.439367737000566],
                               rtol=1e-12)


def test_transform_pass_counts():
    # burnInDoc early exit only at multiples of ChangeEvaluationFrequency (lda.go:224-259)
    lda = O.LatentDirichletAllocation(3, seed=0)
    lda.PerplexityEvaluationFrequency = 2
    lda.FitTransform(FIXTURE_B)
    assert lda.iterationsRun == 58
    _, passes = lda.Transform(FIXTURE_B, return_passes=True)
    assert passes.tolist() == [90, 90, 90, 90, 60, 450, 60, 60, 60]


def test_sparse_and_dense_inputs_agree():
    # utils.go:45-59: CSC iteration == dense scan skipping exact zeros
    import scipy.sparse as sp
    a = O.LatentDirichletAllocation(3, seed=0)
    b = O.LatentDirichletAllocation(3, seed=0)
    a.Iterations = b.Iterations = 20
    ta = a.FitTransform(FIXTURE_B)
    tb = b.FitTransform(sp.csr_matrix(FIXTURE_B))
    assert np.array_equal(ta, tb)


def test_refit_quirks():
    # wordsInCorpus is accumulated, never reset (lda.go:481); rhoThetaT keeps counting (lda.go:497,502)
    lda = O.LatentDirichletAllocation(3, seed=0)
    lda.Iterations = 4
    lda.Fit(FIXTURE_A)
    assert lda.wordsInCorpus == FIXTURE_A.sum()
    assert lda.rhoThetaT == 5
    lda.Fit(FIXTURE_A)
    assert lda.wordsInCorpus == 2 * FIXTURE_A.sum()
    assert lda.rhoThetaT == 9


def test_processes_gt_one_runs():
    # goroutine pool restatement (lda.go:504-528): 3 minibatches over 2 workers
    lda = O.LatentDirichletAllocation(3, seed=0, processes=2)
    lda.BatchSize = 4
    lda.Iterations = 30
    theta = lda.FitTransform(FIXTURE_A)
    assert np.abs(1 - theta.sum(axis=0)).max() <= 1e-8
    assert lda.tokenVisits == 2 * 27 * 30


def _golden():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lda_golden.json")))


def _golden_input(g, name):
    import corpus_synth
    if name.startswith("fixture_A"):
        return FIXTURE_A
    if name.startswith("fixture_B"):
        return FIXTURE_B
    s = g["synthetic"]
    indptr, ind, data = corpus_synth.generate(s["D"], s["W"], block=s["block"], mean_distinct=s["mean_distinct"],
                                              n_topics=s["n_topics"])
    assert (int(indptr[-1]), int(ind.sum()), float(data.sum())) == (s["nnz"], s["ind_checksum"], s["data_checksum"])
    return (s["W"], s["D"], indptr, ind, data)


def test_committed_golden_vectors_reproduce():
    # tests/golden/lda_golden.json (tests/golden/make_golden.py): the oracle and the corpus generator are stable
    g = _golden()
    for case in g["cases"]:
        o = O.LatentDirichletAllocation(case["K"], seed=0)
        for k, v in case["fields"].items():
            setattr(o, k, v)
        theta = o.FitTransform(_golden_input(g, case["name"]))
        assert o.iterationsRun == case["iterations_run"], case["name"]
        np.testing.assert_allclose(o.perplexity_trace, case["perplexity_trace"], rtol=1e-9)
        np.testing.assert_allclose(o.get_state()[1], case["nz"], rtol=1e-9)
        np.testing.assert_allclose(theta[:, :4], case["theta_first_docs"], rtol=1e-7, atol=1e-12)
        assert [o.rhoPhiT, o.rhoThetaT, o.wordsInCorpus] == case["counters"]
        assert list(o.rnd_state) == case["rnd_state"]


def test_hoisted_pow_keeps_every_bit():
    """HoistPow evaluates pow((1 - rhoTheta), v) once per token instead of 2K times (lda.go:247-248); pow is pure, so
    FitTransform, the perplexity trace, Transform and the model state are bit-identical -- which is what lets the large
    parity tests use it"""
    import corpus_synth
    D, W, K = 300, 900, 17
    m = (W, D) + tuple(corpus_synth.generate(D, W, block=64, mean_distinct=30, n_topics=16))
    outs = []
    for hoist in (0, 1):
        o = O.LatentDirichletAllocation(K, seed=0)
        o.Iterations, o.BatchSize, o.PerplexityEvaluationFrequency, o.PerplexityTolerance, o.HoistPow = 3, 64, 1, 0.0, hoist
        theta = o.FitTransform(m)
        outs.append((theta, o.perplexity_trace, o.Transform(m), o.get_state()[0], o.get_state()[1]))
    assert all(np.array_equal(a, b) for a, b in zip(*outs))

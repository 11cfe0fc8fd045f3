"""CUDA parity tests (run on the B200 box with -m gpu).  Every test calls the product through the C ABI of
liblda_b200.so (via nlp_b200.lda) and checks it against the CPU oracle (oracle/) on the same seeded inputs, against
the reference's own fixtures (lda_test.go), or through size-independent properties.

Stated tolerances (fp32 device state vs the reference's float64):
  doc-topic matrix theta       max |delta| <= 2e-3 absolute   (values in [0,1])
  per-iteration perplexity     relative    <= 1e-3            (north_star's bound; observed ~1e-5)
  token indices / CSR handling bit-exact
"""
import numpy as np
import pytest

import corpus_synth
import nlp_b200
from nlp_b200 import _lib, rand
from nlp_b200.lda import CSC, CSR, as_csc, as_matrix
from oracle import oracle as O
from tests.fixtures import EXPECTED_DOCS, FIT_CASES, FIXTURE_A, FIXTURE_B

pytestmark = pytest.mark.gpu

THETA_ATOL = 2e-3
PERP_RTOL = 1e-3


def new_lda(k, seed=0, **fields):
    lda = nlp_b200.NewLatentDirichletAllocation(k)
    lda.Rnd = rand.New(rand.NewSource(seed))
    lda.Processes = 1   # the deterministic parity mode (the default is the host's core count, lda.go:185)
    for f, v in fields.items():
        assert hasattr(lda, f), f
        setattr(lda, f, v)
    return lda


def new_oracle(k, seed=0, **fields):
    o = O.LatentDirichletAllocation(k, seed=seed)
    for f, v in fields.items():
        setattr(o, f, v)
    return o


# ---- the reference's own tests, run against the CUDA path (lda_test.go) ------------------------------------
@pytest.mark.parametrize("data,expected", FIT_CASES)
def test_lda_fit(data, expected):
    lda = new_lda(3)
    lda.Fit(data)
    components = lda.Components()
    assert components.shape == (3, 9)
    assert np.abs(components - expected).max() <= 0.01      # lda_test.go:80
    assert np.abs(1 - components.sum(axis=1)).max() <= 1e-8  # lda_test.go:84


@pytest.mark.parametrize("data", [FIXTURE_A, FIXTURE_B])
def test_lda_fit_transform(data):
    lda = new_lda(3)
    theta = lda.FitTransform(data)
    assert theta.shape == (3, 9)
    assert np.abs(theta.T - EXPECTED_DOCS).max() <= 0.01  # lda_test.go:168
    assert np.abs(1 - theta.sum(axis=0)).max() <= 1e-8    # lda_test.go:172


@pytest.mark.parametrize("data", [FIXTURE_A, FIXTURE_B])
def test_lda_transform(data):
    lda = new_lda(3, PerplexityEvaluationFrequency=2)
    theta = lda.FitTransform(data)
    t_theta = lda.Transform(data)
    assert np.abs(theta - t_theta).max() <= 0.035  # lda_test.go:231


# ---- parity with the oracle -------------------------------------------------------------------------------
@pytest.mark.parametrize("data", [FIXTURE_A, FIXTURE_B])
def test_fixture_parity_trace(data):
    f = dict(PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0, Iterations=60)
    o = new_oracle(3, **f)
    want = o.FitTransform(data)
    lda = new_lda(3, **f)
    got = lda.FitTransform(data)
    assert lda.IterationsRun == o.iterationsRun == 60
    np.testing.assert_allclose(lda.PerplexityTrace, o.perplexity_trace, rtol=PERP_RTOL)
    assert np.abs(got - want).max() <= THETA_ATOL
    np.testing.assert_allclose(lda.Components(), o.Components(), atol=THETA_ATOL)
    nphi_o, nz_o = o.get_state()
    nphi, nz = lda.GetState()
    np.testing.assert_allclose(nz, nz_o, rtol=1e-3)
    np.testing.assert_allclose(nphi, nphi_o, rtol=1e-2, atol=1e-3)
    assert (lda.rhoPhiT, lda.rhoThetaT, lda.wordsInCorpus) == (o.rhoPhiT, o.rhoThetaT, o.wordsInCorpus)


def test_default_early_stop_matches_oracle():
    # defaults: evaluation every 30 iterations, stop when |delta| < 1e-2 -> iteration 180
    o = new_oracle(3)
    o.Fit(FIXTURE_A)
    lda = new_lda(3)
    lda.Fit(FIXTURE_A)
    assert lda.IterationsRun == o.iterationsRun == 180
    np.testing.assert_allclose(lda.PerplexityTrace, o.perplexity_trace, rtol=PERP_RTOL)


def synth(D, W, mean_distinct, block=100, seed=12345):
    indptr, ind, data = corpus_synth.generate(D, W, seed=seed, block=block, mean_distinct=mean_distinct, n_topics=16)
    return CSC(W, D, indptr, ind, data)


def run_pair(m, k, iters, seed=0, theta_atol=THETA_ATOL, cuda_only=None, **fields):
    f = dict(PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0, Iterations=iters)
    f.update(fields)
    o = new_oracle(k, seed=seed, **f)
    want = o.FitTransform(tuple(m))
    lda = new_lda(k, seed=seed, **dict(f, **(cuda_only or {})))
    got = lda.FitTransform(m)
    np.testing.assert_allclose(lda.PerplexityTrace, o.perplexity_trace, rtol=PERP_RTOL)
    assert np.abs(got - want).max() <= theta_atol
    assert np.abs(1 - got.sum(axis=0)).max() <= 1e-8
    return lda, o


def test_config1_parity():
    """BASELINE configs[0]: K=10, 1k docs x 5k vocab (~100k tokens), Processes=1, fixed seed, BatchSize=100"""
    m = synth(1000, 5000, 100)
    lda, o = run_pair(m, 10, 20)
    rel = np.abs(np.array(lda.PerplexityTrace) / o.perplexity_trace - 1).max()
    print("config1: max rel perplexity delta %.3g, final perplexity %.4f" % (rel, lda.PerplexityTrace[-1]))
    assert lda.PerplexityTrace[-1] < lda.PerplexityTrace[0]
    nphi_o, nz_o = o.get_state()
    nphi, nz = lda.GetState()
    np.testing.assert_allclose(nz, nz_o, rtol=1e-3)


@pytest.mark.parametrize("k,D,W,b", [(1, 64, 200, 16), (3, 130, 300, 50), (16, 200, 500, 64), (33, 150, 400, 100),
                                     (64, 150, 400, 40), (100, 300, 800, 100), (200, 120, 600, 32),
                                     (256, 300, 1000, 128), (500, 60, 500, 32), (1024, 48, 400, 16)])
def test_topic_widths(k, D, W, b):
    """every (LANES, NV) kernel instantiation, pad lanes, ragged last minibatch"""
    m = synth(D, W, 40, block=b)
    run_pair(m, k, 4, BatchSize=b)


@pytest.mark.parametrize("k,procs,b,D", [(12, 3, 40, 410), (256, 4, 64, 300), (100, 16, 16, 200), (8, 40, 10, 95)])
def test_processes_concurrent_minibatches(k, procs, b, D):
    """Processes = P on one GPU: P minibatches per launch from one nPhi snapshot, blended in index order -- the
    oracle's SyncGroup = P restatement of lda.go:487-528 (at most 256 per launch)."""
    m = synth(D, 500, 30, block=b)
    f = dict(PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0, Iterations=4, BatchSize=b)
    o = new_oracle(k, SyncGroup=min(procs, 256), **f)
    want = o.FitTransform(tuple(m))
    lda = new_lda(k, Processes=procs, **f)
    got = lda.FitTransform(m)
    np.testing.assert_allclose(lda.PerplexityTrace, o.perplexity_trace, rtol=PERP_RTOL)
    assert np.abs(got - want).max() <= THETA_ATOL
    assert (lda.rhoPhiT, lda.rhoThetaT) == (o.rhoPhiT, o.rhoThetaT)
    np.testing.assert_allclose(lda.GetState()[1], o.get_state()[1], rtol=1e-3)


def test_committed_golden_vectors():
    """the CUDA path against tests/golden/lda_golden.json (oracle-derived, see tests/golden/make_golden.py)"""
    from tests.test_oracle_golden import _golden, _golden_input
    g = _golden()
    for case in g["cases"]:
        f = dict(case["fields"])
        sync = f.pop("SyncGroup", 1)
        lda = new_lda(case["K"], Processes=sync, **f)
        m = _golden_input(g, case["name"])
        theta = lda.FitTransform(CSC(*m) if isinstance(m, tuple) else m)
        assert lda.IterationsRun == case["iterations_run"], case["name"]
        np.testing.assert_allclose(lda.PerplexityTrace, case["perplexity_trace"], rtol=PERP_RTOL)
        assert np.abs(theta[:, :4] - np.array(case["theta_first_docs"])).max() <= THETA_ATOL
        np.testing.assert_allclose(lda.GetState()[1], case["nz"], rtol=1e-3)
        assert [lda.rhoPhiT, lda.rhoThetaT, lda.wordsInCorpus] == case["counters"]
        assert list(lda.Rnd.src.state) == case["rnd_state"]


def test_config1_full_fit_parity():
    """configs[0] run to completion: 100 iterations of K=10 on 1k docs x 5k vocab, BatchSize 100, Processes 1, seed 0;
    perplexity checked every 10 iterations.  fp32 device state must not drift from the float64 oracle."""
    m = synth(1000, 5000, 100)
    f = dict(PerplexityEvaluationFrequency=10, PerplexityTolerance=0.0, Iterations=100)
    o = new_oracle(10, **f)
    want = o.FitTransform(tuple(m))
    lda = new_lda(10, **f)
    got = lda.FitTransform(m)
    rel = np.abs(np.array(lda.PerplexityTrace) / o.perplexity_trace - 1)
    print("config1 full fit: perplexity", np.round(o.perplexity_trace, 3).tolist(), "max rel delta %.3g" % rel.max(),
          "theta max abs delta %.3g" % np.abs(got - want).max())
    assert rel.max() <= PERP_RTOL
    assert np.abs(got - want).max() <= 5e-3
    assert np.mean(got.argmax(0) == want.argmax(0)) >= 0.99


def test_burn_in_passes_and_real_values():
    m = synth(120, 300, 30)
    rng = np.random.default_rng(1)
    m = m._replace(data=m.data * rng.uniform(0.25, 2.5, size=m.data.size))  # TF-IDF-like non-integer values
    run_pair(m, 8, 5, BurnInPasses=3, BatchSize=50)
    run_pair(m, 8, 3, BurnInPasses=0, BatchSize=120)


def test_burn_in_change_check_in_training():
    # BurnInPasses >= ChangeEvaluationFrequency makes burnInDoc's early exit fire during training (lda.go:224-259)
    m = synth(60, 200, 20)
    run_pair(m, 5, 3, BurnInPasses=12, ChangeEvaluationFrequency=4, MeanChangeTolerance=1e-3, BatchSize=30)


def test_unsorted_token_order_is_preserved():
    # a CSC whose rows are NOT ascending inside a column: the sequential token chain must follow stored order
    m = synth(80, 200, 25)
    rng = np.random.default_rng(2)
    ind, data = m.ind.copy(), m.data.copy()
    for j in range(m.cols):
        s, e = m.indptr[j], m.indptr[j + 1]
        p = rng.permutation(e - s)
        ind[s:e], data[s:e] = ind[s:e][p], data[s:e][p]
    shuffled = m._replace(ind=ind, data=data)
    run_pair(shuffled, 6, 5)
    a = new_lda(6, Iterations=5).FitTransform(m)
    b = new_lda(6, Iterations=5).FitTransform(shuffled)
    assert np.abs(a - b).max() > 1e-6  # order matters (lda.go:236 -> 247), so the two inputs must differ


def test_empty_documents_and_edge_shapes():
    dense = np.zeros((12, 7))
    dense[[0, 3, 5], 0] = [1, 2, 3]
    dense[2, 2] = 4       # single-token document
    dense[:, 4] = 1       # document using every word
    # columns 1, 3, 5, 6 are empty documents
    f = dict(PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0, Iterations=6, BatchSize=3)
    o = new_oracle(4, **f)
    want = o.FitTransform(dense)
    lda = new_lda(4, **f)
    got = lda.FitTransform(dense)
    assert np.abs(got - want).max() <= THETA_ATOL
    np.testing.assert_allclose(lda.PerplexityTrace, o.perplexity_trace, rtol=PERP_RTOL)
    t_o, p_o = o.Transform(dense, return_passes=True)
    t, p = lda.Transform(dense, return_passes=True)
    assert np.abs(t - t_o).max() <= THETA_ATOL
    assert p.tolist() == p_o.tolist()
    # a single document / single word
    one = np.array([[3.0]])
    assert np.abs(new_lda(2, Iterations=3).FitTransform(one) - new_oracle(2, Iterations=3).FitTransform(one)).max() <= 1e-5


def test_sparse_inputs_match_dense():
    import scipy.sparse as sp
    a = new_lda(3, Iterations=10).FitTransform(FIXTURE_B)
    for conv in (sp.csr_matrix, sp.csc_matrix, sp.coo_matrix, sp.dok_matrix):
        b = new_lda(3, Iterations=10).FitTransform(conv(FIXTURE_B))
        assert np.array_equal(a, b), conv.__name__


def test_transform_and_perplexity_parity():
    m = synth(300, 800, 40)
    held = synth(90, 800, 40, seed=777)
    f = dict(Iterations=15, PerplexityEvaluationFrequency=5, PerplexityTolerance=0.0)
    lda, o = run_pair(m, 12, 15, **f)
    # same model on both sides so that Transform parity is isolated from fit drift
    nphi, nz = o.get_state()
    lda.SetState(nphi, nz)
    t_o, p_o = o.Transform(tuple(held), return_passes=True)
    t, p = lda.Transform(held, return_passes=True)
    same = p == p_o
    assert same.mean() >= 0.95                # early exit happens at multiples of 30 (lda.go:224-259); a borderline
    assert np.abs(t - t_o)[:, same].max() <= THETA_ATOL   # fp32/float64 decision moves a document by 30 passes,
    assert np.abs(t - t_o).max() <= 0.035                 # which lda_test.go:231 tolerates with 0.035
    assert set(np.unique(p)) <= set(range(30, 501, 30)) | {500}
    perp_o = o.Perplexity(tuple(held))
    perp = lda.Perplexity(held)
    assert abs(perp / perp_o - 1) <= PERP_RTOL
    # the RNG streams advanced identically: D*K draws per Transform / Perplexity call (lda.go:422-426)
    assert lda.Rnd.src.state == o.rnd_state


def test_transform_pass_counts_fixture():
    lda = new_lda(3, PerplexityEvaluationFrequency=2)
    lda.FitTransform(FIXTURE_B)
    assert lda.IterationsRun == 58
    _, passes = lda.Transform(FIXTURE_B, return_passes=True)
    assert passes.tolist() == [90, 90, 90, 90, 60, 450, 60, 60, 60]


def test_refit_quirks():
    # wordsInCorpus accumulates, rhoThetaT keeps counting, nPhi is re-drawn
    f = dict(Iterations=4, PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0)
    lda, o = new_lda(3, **f), new_oracle(3, **f)
    for _ in range(2):
        got, want = lda.FitTransform(FIXTURE_A), o.FitTransform(FIXTURE_A)
        assert np.abs(got - want).max() <= THETA_ATOL
        np.testing.assert_allclose(lda.PerplexityTrace, o.perplexity_trace, rtol=PERP_RTOL)
    assert lda.wordsInCorpus == o.wordsInCorpus == 2 * FIXTURE_A.sum()
    assert lda.rhoThetaT == o.rhoThetaT == 9
    assert lda.Rnd.src.state == o.rnd_state


def test_partial_fit_and_persistence(tmp_path):
    m = synth(240, 500, 30, block=60)
    # one PartialFit on a fresh model == Fit with Iterations = 1 (init draws, rhoPhiT from 1, same minibatches)
    o = new_oracle(7, Iterations=1, BatchSize=60, PerplexityEvaluationFrequency=0)
    want = o.FitTransform(tuple(m))
    lda = new_lda(7, BatchSize=60)
    got = lda.PartialFit(m, want_theta=True)
    assert np.abs(got - want).max() <= THETA_ATOL
    assert (lda.rhoPhiT, lda.rhoThetaT, lda.wordsInCorpus) == (o.rhoPhiT, o.rhoThetaT, o.wordsInCorpus)
    assert lda.Rnd.src.state == o.rnd_state
    np.testing.assert_allclose(lda.GetState()[1], o.get_state()[1], rtol=1e-3)
    # streaming: further PartialFit calls on new documents keep improving held-out perplexity
    held = synth(60, 500, 30, block=60, seed=4242)
    p0 = lda.Perplexity(held)
    for s in range(6):
        lda.PartialFit(synth(240, 500, 30, block=60, seed=100 + s))
    assert lda.rhoPhiT == 1 + 4 * 7 and lda.Perplexity(held) < p0
    with pytest.raises(_lib.LdaB200Error, match="same number of rows"):
        lda.PartialFit(synth(60, 400, 30, block=60))
    # Save / Load: a new object reproduces Transform exactly (same model, same RNG position)
    path = tmp_path / "lda.bin"
    with open(path, "wb") as f:
        lda.Save(f)
    lda2 = new_lda(3)
    with open(path, "rb") as f:
        lda2.Load(f)
    assert (lda2.K, lda2.w, lda2.rhoPhiT, lda2.rhoThetaT, lda2.wordsInCorpus) == (7, 500, lda.rhoPhiT, lda.rhoThetaT, lda.wordsInCorpus)
    lda2.Rnd.src.state = lda.Rnd.src.state
    assert np.array_equal(lda.Transform(held), lda2.Transform(held))
    assert np.array_equal(lda.Components(), lda2.Components())
    # the file is written by the library (lda_b200_save); the Python restatement of the layout gives the same bytes
    from nlp_b200.lda import serialise_model
    nphi, nz = lda.GetState()
    assert open(path, "rb").read() == serialise_model(7, (lda.rhoPhiT, lda.rhoThetaT, lda.wordsInCorpus), nphi, nz)


def test_model_file_fixture_round_trip(tmp_path):
    """Load of the committed fixture (tests/golden/lda_model_k3_w9.bin) and Save through the C ABI reproduce it bit for
    bit; truncated and foreign files are refused"""
    import io
    import os
    raw = open(os.path.join(os.path.dirname(__file__), "golden", "lda_model_k3_w9.bin"), "rb").read()
    lda = new_lda(5)
    lda.Load(io.BytesIO(raw))
    assert (lda.K, lda.w, lda.rhoPhiT, lda.rhoThetaT, lda.wordsInCorpus) == (3, 9, 5.0, 37.0, 1234.0)
    out = io.BytesIO()
    lda.Save(out)
    assert out.getvalue() == raw
    assert lda.Transform(FIXTURE_A).shape == (3, 9)
    for bad in (raw[:100], raw[:-1], b"\x00" * len(raw), raw[:32] + b"XXXX" + raw[36:]):
        with pytest.raises((_lib.LdaB200Error, ValueError), match="EOF|mat.Dense|dimensions"):
            new_lda(3).Load(io.BytesIO(bad))


# ---- bit-exact pieces -------------------------------------------------------------------------------------------
def test_ingest_roundtrip_bit_exact():
    import ctypes as C
    m = synth(257, 1000, 30)
    lda = new_lda(4)
    h = lda._handle()
    nnz = m.ind.size
    indptr = np.zeros(m.cols + 1, dtype=np.int64)
    ind = np.zeros(nnz, dtype=np.int32)
    val = np.zeros(nnz, dtype=np.float32)
    wc = np.zeros(m.cols, dtype=np.float64)
    n = C.c_double(0)
    _lib.check(_lib.lib().lda_b200_ingest_roundtrip(h, m.rows, m.cols, m.indptr.ctypes.data, m.ind.ctypes.data,
                                                   m.data.ctypes.data, indptr.ctypes.data, ind.ctypes.data,
                                                   val.ctypes.data, wc.ctypes.data, C.byref(n)), h)
    assert np.array_equal(indptr, m.indptr)
    assert np.array_equal(ind.astype(np.int64), m.ind)
    assert np.array_equal(val.astype(np.float64), m.data)  # integer counts are exact in fp32
    assert np.array_equal(wc, np.add.reduceat(m.data, m.indptr[:-1]))  # lda.go:476-480
    assert n.value == m.data.sum()                                     # lda.go:481


def _device_view(lda, m):
    """(indptr, ind, val, wc, N) the device holds for matrix m after the *_matrix ingestion (CSC / CSR / dense)"""
    import ctypes as C
    desc, rows, cols, _keep = as_matrix(m)
    h = lda._handle()
    nnz = C.c_int64(0)
    L = _lib.lib()
    _lib.check(L.lda_b200_ingest_roundtrip_matrix(h, desc, C.byref(nnz), None, None, None, None, None), h)
    indptr = np.zeros(cols + 1, dtype=np.int64)
    ind = np.zeros(nnz.value, dtype=np.int32)
    val = np.zeros(nnz.value, dtype=np.float32)
    wc = np.zeros(cols, dtype=np.float64)
    n = C.c_double(0)
    _lib.check(L.lda_b200_ingest_roundtrip_matrix(h, desc, None, indptr.ctypes.data, ind.ctypes.data, val.ctypes.data,
                                                  wc.ctypes.data, C.byref(n)), h)
    return indptr, ind, val, wc, n.value


def test_matrix_ingestion_bit_exact():
    """CSR (sparse.ToCSC, lda.go:464-466) and dense (utils.go:51-57) inputs converted on the device give exactly the
    token stream the host-side conversion gives: same indptr, word ids, values, wc (lda.go:476-480) and N (:481)."""
    import scipy.sparse as sp
    lda = new_lda(3)
    rng = np.random.default_rng(11)
    for (W, D, dens) in [(40, 33, 0.2), (500, 260, 0.03), (9, 2000, 0.3), (200, 1, 0.5), (1, 50, 0.6), (30, 30, 0.0)]:
        a = sp.random(W, D, density=dens, format="csr", random_state=rng, data_rvs=lambda n: rng.integers(1, 9, n).astype(float))
        want = as_csc(a)
        wc = np.zeros(D)
        np.add.at(wc, np.repeat(np.arange(D), np.diff(want.indptr)), want.data)
        dense = a.toarray()
        dense[dense == 0] = 0.0
        if W > 2 and D > 2:
            dense[1, 2] = -0.0   # Go's `v != 0` skips a negative zero as well
            a = sp.csr_matrix(dense)
            want = as_csc(a)
            wc = np.zeros(D)
            np.add.at(wc, np.repeat(np.arange(D), np.diff(want.indptr)), want.data)
        for m in (a, CSR(W, D, a.indptr, a.indices, a.data), dense, want, sp.csc_matrix(a)):
            indptr, ind, val, wc_d, n = _device_view(lda, m)
            assert np.array_equal(indptr, want.indptr), type(m)
            assert np.array_equal(ind.astype(np.int64), want.ind), type(m)
            assert np.array_equal(val.astype(np.float64), want.data), type(m)
            assert np.array_equal(wc_d, wc) and n == want.data.sum(), type(m)
    # CSR rows whose document ids are not sorted: words still come out ascending inside a document
    a = sp.random(60, 45, density=0.2, format="csr", random_state=rng, data_rvs=lambda n: rng.integers(1, 5, n).astype(float))
    ptr, ci, dv = a.indptr.copy(), a.indices.copy(), a.data.copy()
    for r in range(60):
        q = rng.permutation(ptr[r + 1] - ptr[r])
        ci[ptr[r]:ptr[r + 1]], dv[ptr[r]:ptr[r + 1]] = ci[ptr[r]:ptr[r + 1]][q], dv[ptr[r]:ptr[r + 1]][q]
    want = as_csc(a)
    indptr, ind, val, _, _ = _device_view(lda, CSR(60, 45, ptr, ci, dv))
    assert np.array_equal(indptr, want.indptr) and np.array_equal(ind, want.ind) and np.array_equal(val, want.data)


def test_matrix_inputs_through_every_call():
    """Fit / Transform / Perplexity / PartialFit give the same results whether the conversion to CSC ran on the device
    (CSR, dense) or on the host, and they match the oracle."""
    import scipy.sparse as sp
    m = synth(150, 400, 30)
    csc = sp.csc_matrix((m.data, m.ind, m.indptr), shape=(m.rows, m.cols))
    f = dict(Iterations=4, PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0, BatchSize=64)
    base, o = run_pair(m, 7, 4, BatchSize=64)
    th0 = new_lda(7, **f).FitTransform(m)
    t0, p0 = base.Transform(m), base.Perplexity(m)
    for x in (csc.tocsr(), csc.toarray(), CSR(m.rows, m.cols, *[getattr(csc.tocsr(), a) for a in ("indptr", "indices", "data")])):
        lda = new_lda(7, **f)
        th = lda.FitTransform(x)
        assert np.abs(th - th0).max() <= 1e-4
        np.testing.assert_allclose(lda.PerplexityTrace, o.perplexity_trace, rtol=PERP_RTOL)
        np.testing.assert_allclose(lda.PerplexityTrace, base.PerplexityTrace, rtol=1e-5)
        # same model, same initial theta: the device holds the same token stream, so Transform is identical
        base.Rnd = rand.New(rand.NewSource(5))
        tx = base.Transform(x)
        base.Rnd = rand.New(rand.NewSource(5))
        assert np.array_equal(tx, base.Transform(m))
        base.Rnd = rand.New(rand.NewSource(5))
        px = base.Perplexity(x)
        base.Rnd = rand.New(rand.NewSource(5))
        np.testing.assert_allclose(px, base.Perplexity(m), rtol=1e-12)
        host = new_lda(7, DeviceConvert=False, **f)
        host.FitTransform(x)
        np.testing.assert_allclose(host.PerplexityTrace, lda.PerplexityTrace, rtol=1e-5)
    assert np.isfinite(t0).all() and np.isfinite(p0)
    pf = new_lda(7, BatchSize=64)
    pf.PartialFit(csc.tocsr(), iterations=2)
    pf2 = new_lda(7, BatchSize=64)
    pf2.PartialFit(m, iterations=2)
    np.testing.assert_allclose(pf.GetState()[1], pf2.GetState()[1], rtol=1e-5)
    # a matrix with the wrong number of rows is rejected instead of indexing outside nPhi
    with pytest.raises(_lib.LdaB200Error, match="rows"):
        base.Transform(np.ones((m.rows + 1, 3)))
    with pytest.raises(_lib.LdaB200Error, match="rows"):
        base.Perplexity(sp.csr_matrix(np.ones((m.rows - 1, 3))))


def test_device_csr_to_csc_bit_exact():
    # the reference converts with sparse.ToCSC() on the host (lda.go:464-466); the device transpose must give the
    # identical CSC (rows ascending inside every document), including empty rows / documents
    import scipy.sparse as sp
    lda = new_lda(3)
    rng = np.random.default_rng(3)
    for (W, D, dens) in [(50, 70, 0.1), (1000, 300, 0.02), (7, 5000, 0.3), (300, 1, 0.5), (1, 40, 0.5), (64, 64, 0.0)]:
        a = sp.random(W, D, density=dens, format="csr", random_state=rng, data_rvs=lambda n: rng.integers(1, 9, n).astype(float))
        got = lda.ToCSC(a)
        want = a.tocsc()
        want.sort_indices()
        assert (got.rows, got.cols) == (W, D)
        assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.ind, want.indices)
        assert np.array_equal(got.data, want.data)
    m = synth(400, 3000, 60)
    import scipy.sparse as sp2
    csc = sp2.csc_matrix((m.data, m.ind, m.indptr), shape=(m.rows, m.cols))
    got = lda.ToCSC(csc.tocsr())
    assert np.array_equal(got.indptr, m.indptr) and np.array_equal(got.ind, m.ind) and np.array_equal(got.data, m.data)
    bad = csc.tocsr().copy()
    bad.indices = bad.indices.astype(np.int64)
    bad.indices[3] = 400
    with pytest.raises(_lib.LdaB200Error, match="outside"):
        lda.ToCSC(bad)


def test_device_pcg_init_is_bit_exact():
    # Iterations = 0: the state read back is the initial draw of lda.go:199-205, produced on the GPU
    W, D, K = 37, 11, 7
    dense = np.ones((W, D))
    lda = new_lda(K, seed=42, Iterations=0)
    theta = lda.FitTransform(dense)
    nphi, nz = lda.GetState()
    ref = O.PCG(42)
    nphi_ref = ref.fill(W * K, W * K).reshape(W, K)
    ntheta_ref = ref.fill(D * K, D * K).reshape(D, K)
    assert np.array_equal(nphi, nphi_ref.astype(np.float32).astype(np.float64))
    np.testing.assert_allclose(nz, nphi_ref.astype(np.float32).astype(np.float64).sum(0), rtol=1e-12)
    th32 = ntheta_ref.astype(np.float32).astype(np.float64)
    np.testing.assert_allclose(theta, (th32 / th32.sum(1, keepdims=True)).T, rtol=1e-12)
    assert lda.Rnd.src.state == ref.state
    # host-drawn init (DeviceInit = False) gives the identical model
    lda2 = new_lda(K, seed=42, Iterations=0, DeviceInit=False)
    theta2 = lda2.FitTransform(dense)
    assert np.array_equal(theta, theta2) and lda2.Rnd.src.state == ref.state


def test_exported_init_replay():
    # "same seeded init of nPhi/nTheta exported from the reference": arrays in, no RNG involved
    m = synth(100, 300, 30)
    K = 9
    src = O.PCG(5)
    nphi0 = src.fill(m.rows * K, m.rows * K)
    ntheta0 = src.fill(m.cols * K, m.cols * K)
    f = dict(Iterations=6, PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0)
    o = new_oracle(K, seed=99, **f)
    want = o.FitTransform(tuple(m), nphi0=nphi0, ntheta0=ntheta0)
    lda = new_lda(K, seed=1234, **f)
    got = lda.FitTransform(m, nphi0=nphi0, ntheta0=ntheta0)
    assert np.abs(got - want).max() <= THETA_ATOL
    np.testing.assert_allclose(lda.PerplexityTrace, o.perplexity_trace, rtol=PERP_RTOL)
    assert lda.Rnd.src.state == rand.NewSource(1234).state  # untouched


# ---- error behaviour -----------------------------------------------------------------------------------------------
def test_errors():
    lda = new_lda(3, Iterations=2)
    with pytest.raises(_lib.LdaB200Error, match="before Fit"):
        lda.Transform(FIXTURE_A)
    bad = as_csc(FIXTURE_A)
    bad.ind[4] = 9  # row index out of range
    with pytest.raises(_lib.LdaB200Error, match="outside"):
        lda.Fit(bad)
    with pytest.raises(_lib.LdaB200Error, match="1024"):
        new_lda(2048).Fit(FIXTURE_A)
    with pytest.raises(_lib.LdaB200Error, match="BatchSize"):
        new_lda(3, BatchSize=0).Fit(FIXTURE_A)
    lda.Fit(FIXTURE_A)  # the handle is still usable after errors
    # the same checks on the Transform / Perplexity pipeline and on the device-side CSR conversion
    with pytest.raises(_lib.LdaB200Error, match="outside"):
        lda.Transform(bad)
    with pytest.raises(_lib.LdaB200Error, match="outside"):
        lda.Perplexity(bad)
    import scipy.sparse as sp
    csr = sp.csr_matrix(FIXTURE_A)
    bad_csr = CSR(9, 9, csr.indptr.astype(np.int64), csr.indices.astype(np.int64), csr.data.copy())
    bad_csr.ind[2] = 9  # document id out of range
    with pytest.raises(_lib.LdaB200Error, match="outside"):
        lda.Transform(bad_csr)
    with pytest.raises(_lib.LdaB200Error, match="non-decreasing"):
        lda.Fit(as_csc(FIXTURE_A)._replace(indptr=np.array([0, 3, 2, 9, 12, 15, 18, 21, 24, 27], dtype=np.int64)))
    t = lda.Transform(FIXTURE_A)
    assert np.abs(t.sum(0) - 1).max() <= 1e-8  # still usable


def test_changing_k_after_fit_drops_the_model():
    """lda.K is an exported field (lda.go:92).  nPhi / nZ on the device are laid out for the K of the fit, and callers
    size their result buffers by the new one, so a changed K must not reach Transform / Components / GetState."""
    lda = new_lda(3, Iterations=3)
    lda.Fit(FIXTURE_A)
    assert lda.Components().shape == (3, 9)
    lda.K = 2
    for call in (lambda: lda.Transform(FIXTURE_A), lambda: lda.Perplexity(FIXTURE_A), lda.Components, lda.GetState):
        with pytest.raises(_lib.LdaB200Error, match="before Fit|no model state"):
            call()
    lda.Fit(FIXTURE_A)      # a new fit with the new K works
    assert lda.Transform(FIXTURE_A).shape == (2, 9) and lda.Components().shape == (2, 9)
    lda.K = 5               # and so does loading a state of the new shape
    lda.SetState(np.random.default_rng(0).random((9, 5)), np.ones(5))
    assert lda.Transform(FIXTURE_A).shape == (5, 9)


def test_failed_fit_leaves_no_model():
    """out-of-range word ids are found while the first iteration already runs (they stream in behind the first
    launches); the fit fails and must not leave a half-trained model behind that Transform would accept"""
    lda = new_lda(3, Iterations=2)
    lda.Fit(FIXTURE_A)
    bad = as_csc(FIXTURE_A)
    bad.ind[4] = 9
    with pytest.raises(_lib.LdaB200Error, match="outside"):
        lda.Fit(bad)
    with pytest.raises(_lib.LdaB200Error, match="before Fit"):
        lda.Transform(FIXTURE_A)


# ---- oracle parity on a BASELINE configuration in full, and at K = 256 over several large minibatches ---------------------
@pytest.mark.parametrize("b", [100, 2048], ids=["BatchSize100-reference-default", "BatchSize2048"])
def test_config2_full_parity(b):
    """BASELINE configs[1] in full -- K = 100, 18,846 documents x 50k words, ~2.5M non-zeros -- for 5 iterations at
    Processes = 1 against the oracle: per-iteration perplexity within 1e-3 relative, doc-topic matrix within 2e-3
    (BASELINE.md section 3).  The oracle runs with HoistPow (bit-identical, see oracle/lda_oracle.c) to finish in seconds."""
    D, W, K, md = corpus_synth.shape("C2")
    m = synth(D, W, md, block=2048)
    f = dict(PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0, Iterations=5, BatchSize=b)
    o = new_oracle(K, HoistPow=1, **f)
    want = o.FitTransform(tuple(m))
    lda = new_lda(K, **f)
    got = lda.FitTransform(m)
    rel = np.abs(np.array(lda.PerplexityTrace) / o.perplexity_trace - 1).max()
    print("config2 b=%d: max rel perplexity delta %.3g, max |dtheta| %.3g, perplexity %s" % (
        b, rel, np.abs(got - want).max(), np.round(o.perplexity_trace, 1)))
    assert rel <= PERP_RTOL
    assert np.abs(got - want).max() <= THETA_ATOL
    np.testing.assert_allclose(lda.GetState()[1], o.get_state()[1], rtol=1e-3)
    assert (lda.rhoPhiT, lda.rhoThetaT, lda.wordsInCorpus) == (o.rhoPhiT, o.rhoThetaT, o.wordsInCorpus)


@pytest.mark.parametrize("procs", [1, 4], ids=["Processes1", "Processes4"])
def test_k256_large_minibatches_parity(procs):
    """the K = 256 throughput kernel on 4,500 documents in minibatches of 1,024 (five minibatches, the last one short), one
    at a time and four per launch, 4 iterations, against the oracle"""
    m = synth(4500, 20000, 120, block=1024)
    f = dict(PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0, Iterations=4, BatchSize=1024)
    o = new_oracle(256, HoistPow=1, SyncGroup=procs, **f)
    want = o.FitTransform(tuple(m))
    lda = new_lda(256, Processes=procs, **f)
    got = lda.FitTransform(m)
    np.testing.assert_allclose(lda.PerplexityTrace, o.perplexity_trace, rtol=PERP_RTOL)
    assert np.abs(got - want).max() <= THETA_ATOL
    np.testing.assert_allclose(lda.GetState()[1], o.get_state()[1], rtol=1e-3)
    held = synth(1500, 20000, 120, block=1024, seed=777)
    o.set_state(*lda.GetState())
    t, passes = lda.Transform(held, return_passes=True)
    t_want, p_want = o.Transform(tuple(held), return_passes=True)
    same = passes == p_want
    assert same.mean() >= 0.95 and np.abs(t - t_want)[:, same].max() <= THETA_ATOL and np.abs(t - t_want).max() <= 0.035


# ---- size-independent properties at a BASELINE-sized configuration -----------------------------------------------------
def test_config2_properties():
    """configs[1] shape (18,846 docs x 50k vocab, K=100): columns sum to 1, perplexity ends up falling, refit reproducible
    up to the order of the fp32 reductions into nPhiHat."""
    D, W, K, md = corpus_synth.shape("C2")
    indptr, ind, data = corpus_synth.generate(D, W, block=4096, mean_distinct=md)
    m = CSC(W, D, indptr, ind, data)
    f = dict(Iterations=12, BatchSize=1024, PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0)
    lda = new_lda(K, **f)
    theta = lda.FitTransform(m)
    assert theta.shape == (K, D) and np.isfinite(theta).all() and theta.min() >= 0
    assert np.abs(1 - theta.sum(axis=0)).max() <= 1e-8
    tr = np.array(lda.PerplexityTrace)
    print("config2 perplexity trace:", np.round(tr, 2).tolist())
    # SCVB0 from a random init may rise for the first few iterations; it must be falling by the end
    assert np.isfinite(tr).all() and np.all(np.diff(tr)[-3:] < 0) and tr[-1] < tr.max(), tr
    comp = lda.Components()
    assert np.abs(1 - comp.sum(axis=1)).max() <= 1e-8
    lda2 = new_lda(K, **f)
    theta2 = lda2.FitTransform(m)
    # Transform of the training documents lands near FitTransform's theta (lda_test.go:179-235 at scale)
    sub = CSC(W, 512, indptr[:513], ind[:indptr[512]], data[:indptr[512]])
    t = lda.Transform(sub)
    stats = dict(refit_trace_rel=float(np.abs(np.array(lda2.PerplexityTrace) / tr - 1).max()),
                 refit_theta_abs=float(np.abs(theta - theta2).max()),
                 transform_argmax_agree=float(np.mean(t.argmax(0) == theta[:, :512].argmax(0))))
    print("config2 stats:", stats)
    assert np.abs(t.sum(axis=0) - 1).max() <= 1e-8
    # two runs differ only by the order of the fp32 reductions into nPhiHat
    assert stats["refit_trace_rel"] <= 1e-3, stats
    assert stats["transform_argmax_agree"] > 0.5, stats


def test_medium_parity():
    """a mid-sized corpus with long documents (several 32-token tiles), hot Zipf-head rows under concurrent
    reductions, K = 100 (pad lanes), several minibatches, over 10 iterations: the CUDA path must follow the float64
    oracle's per-iteration perplexity (run_pair asserts PERP_RTOL / THETA_ATOL)."""
    indptr, ind, data = corpus_synth.generate(600, 20000, block=128, mean_distinct=133)
    m = CSC(20000, 600, indptr, ind, data)
    lda, o = run_pair(m, 100, 10, BatchSize=128)
    tr = o.perplexity_trace
    print("medium parity: oracle perplexity", np.round(tr, 1).tolist(), "max rel delta",
          np.abs(np.array(lda.PerplexityTrace) / tr - 1).max())
    assert tr[-1] < tr[0]


@pytest.mark.parametrize("k", [8, 128])
def test_transform_pipeline_ranges(k):
    """Transform streams the corpus through the device in ranges of 65,536 documents (run_transform): several ranges,
    the last one short, give exactly what separate calls on the pieces give (documents are independent,
    lda.go:429-435), Perplexity agrees, and a sample of documents matches the oracle."""
    D, W = 140_000, 3000
    fit = synth(400, W, 30)
    held = synth(D, W, 8, block=8192, seed=4242)
    lda = new_lda(k, Iterations=5, BatchSize=100)
    lda.Fit(fit)
    rng = np.random.default_rng(1)
    theta0 = rng.random((D, k))
    whole, passes = lda.Transform(held, theta0=theta0, return_passes=True)
    assert whole.shape == (k, D) and np.abs(whole.sum(0) - 1).max() <= 1e-8
    cuts = [0, 50_000, 65_536, 131_072, D]
    for a, b in zip(cuts[:-1], cuts[1:]):
        sub = CSC(W, b - a, held.indptr[a:b + 1] - held.indptr[a], held.ind[held.indptr[a]:held.indptr[b]],
                  held.data[held.indptr[a]:held.indptr[b]])
        part, pp = lda.Transform(sub, theta0=theta0[a:b], return_passes=True)
        assert np.array_equal(part, whole[:, a:b]) and np.array_equal(pp, passes[a:b]), (a, b)
    # the same through the oracle on a sample that spans the range boundaries
    nphi, nz = lda.GetState()
    o = new_oracle(k)
    o.set_state(nphi, nz)
    o.rhoThetaT = lda.rhoThetaT
    pick = np.r_[0:40, 65_516:65_556, D - 40:D]
    ip = np.concatenate(([0], np.cumsum(np.diff(held.indptr)[pick])))
    idx = np.concatenate([np.arange(held.indptr[j], held.indptr[j + 1]) for j in pick])
    sample = (W, len(pick), ip.astype(np.int64), held.ind[idx], held.data[idx])
    want, wp = o.Transform(sample, theta0=theta0[pick], return_passes=True)
    same = wp == passes[pick]
    assert same.mean() >= 0.95
    assert np.abs(whole[:, pick][:, same] - want[:, same]).max() <= THETA_ATOL
    # Perplexity runs the same pipeline without reading theta back
    p1 = lda.Perplexity(held, theta0=theta0)
    assert np.isfinite(p1) and p1 > 1


def test_deterministic_mode_is_bitwise_reproducible():
    """lda_b200_set_deterministic: fixed-point accumulation of nPhiHat + block-ordered reductions.  Many warps reduce
    into the same hot rows concurrently (Zipf head, 4 minibatches per launch), the setting in which the default fp32
    reductions land in a different order from run to run."""
    indptr, ind, data = corpus_synth.generate(6000, 20000, block=1024, mean_distinct=120)
    m = CSC(20000, 6000, indptr, ind, data)
    f = dict(Iterations=6, BatchSize=1024, Processes=4, PerplexityEvaluationFrequency=2, PerplexityTolerance=0.0)
    runs = []
    for _ in range(3):
        lda = new_lda(100, Deterministic=True, **f)
        theta = lda.FitTransform(m)
        nphi, nz = lda.GetState()
        runs.append((theta, nphi, nz, list(lda.PerplexityTrace), lda.Transform(m), lda.Components()))
        lda.Close()
    for other in runs[1:]:
        for a, b in zip(runs[0], other):
            assert np.array_equal(np.asarray(a), np.asarray(b))
    # the mode changes rounding only: same results as the default mode and as the oracle within the stated tolerances
    plain = new_lda(100, **f)
    theta_plain = plain.FitTransform(m)
    assert np.abs(theta_plain - runs[0][0]).max() <= THETA_ATOL
    np.testing.assert_allclose(plain.PerplexityTrace, runs[0][3], rtol=PERP_RTOL)
    for k, b in ((7, 16), (100, 64), (300, 32)):
        small = synth(150, 500, 30)
        run_pair(small, k, 4, BatchSize=b, cuda_only=dict(Deterministic=True))


def _full_size_properties(name, batch, processes, iters):
    import torch
    D, W, K, md = corpus_synth.shape(name)
    indptr, ind, data = corpus_synth.generate(D, W, block=8192, mean_distinct=md, device=torch.device("cuda", 0))
    m = CSC(W, D, indptr, ind, data)
    lda = new_lda(K, Iterations=iters, BatchSize=batch, Processes=processes, PerplexityEvaluationFrequency=5,
                  PerplexityTolerance=0.0)
    theta = lda.FitTransform(m)
    assert theta.shape == (K, D) and np.isfinite(theta).all() and theta.min() >= 0
    assert np.abs(1 - theta.sum(axis=0)).max() <= 1e-8            # lda_test.go:172 at scale
    assert lda.wordsInCorpus == float(data.sum())                  # lda.go:476-482
    tr = np.array(lda.PerplexityTrace)
    print(name, "perplexity every 5 iterations", np.round(tr, 1).tolist())
    # from a random init the training perplexity first rises for a few iterations; it must be falling by the end
    assert len(tr) == iters // 5 and np.isfinite(tr).all() and tr[-1] < tr[-2] and tr[-1] < tr.max(), tr
    nphi, nz = lda.GetState()
    assert np.isfinite(nphi).all() and nphi.min() >= 0
    # invariant of the algorithm: nZ[k] == sum_w nPhi[w,k] (lda.go:203, then Eqn 7 and 8 are the same affine map)
    np.testing.assert_allclose(nz, nphi.sum(axis=0), rtol=2e-4)
    comp = lda.Components()
    assert comp.shape == (K, W) and np.abs(1 - comp.sum(axis=1)).max() <= 1e-8   # lda_test.go:84 at scale


def test_config3_properties_full_size():
    """configs[2] at full size: 1M docs x 100k vocab, K=256, ~200M non-zeros"""
    _full_size_properties("C3", 8192, 4, 20)


def test_config4_properties_full_size():
    """configs[3] at full size: 200k docs x 100k vocab, K=1024 (nPhi 410 MB, beyond L2)"""
    _full_size_properties("C4", 2048, 4, 20)  # 25 M-steps per iteration: past the initial transient by iteration 20


def test_step_issued_as_sub_range_launches(monkeypatch):
    """A step whose launch range exceeds LDA_B200_SUB_RANGE_DOCS documents is tiled into sub-ranges: one launch over all
    of them in the middle iterations, one launch per sub-range in the first iteration (word ids still streaming in) and
    in the last one (doc-topic rows streaming out to a page-locked result).  Same model as the oracle's SyncGroup."""
    import torch
    monkeypatch.setenv("LDA_B200_SUB_RANGE_DOCS", "100")
    for k, b, procs, D in [(256, 50, 8, 1037), (12, 30, 16, 1000)]:
        m = synth(D, 700, 30, block=b)
        f = dict(PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0, Iterations=3, BatchSize=b)
        o = new_oracle(k, SyncGroup=procs, **f)
        want = o.FitTransform(tuple(m))
        lda = new_lda(k, Processes=procs, **f)
        out = torch.empty((k, D), dtype=torch.float64).pin_memory()
        got = lda.FitTransform(m, out=out.numpy())
        np.testing.assert_allclose(lda.PerplexityTrace, o.perplexity_trace, rtol=PERP_RTOL)
        assert np.abs(got - want).max() <= THETA_ATOL
        lda2 = new_lda(k, Processes=procs, **f)                       # pageable result: no streaming out, same answer
        assert np.abs(lda2.FitTransform(m) - want).max() <= THETA_ATOL


@pytest.mark.parametrize("k,D,W,b", [(100, 300, 800, 100), (200, 120, 600, 32), (256, 300, 1000, 128), (500, 60, 500, 32),
                                     (1024, 48, 400, 16)])
def test_throughput_kernels_on_small_corpora(k, D, W, b, monkeypatch):
    """Fit launch ranges of up to LDA_B200_PAIR_MAX_DOCS documents run the two-tokens-per-butterfly kernels, larger ones the
    throughput kernels (one token per butterfly).  Whatever the default threshold is, this test runs both families on the
    same small corpora against the oracle (fit, then Transform under each fitted model), and the two fitted models against
    each other."""
    monkeypatch.setenv("LDA_B200_PAIR_MAX_DOCS", "0")
    m = synth(D, W, 40, block=b)
    lda, o = run_pair(m, k, 4, BatchSize=b)
    held = synth(80, W, 40, block=b, seed=4242)
    o.set_state(*lda.GetState())
    t, passes = lda.Transform(held, return_passes=True)
    t_want, p_want = o.Transform(tuple(held), return_passes=True)
    same = passes == p_want
    assert same.mean() >= 0.9 and np.abs(t - t_want)[:, same].max() <= THETA_ATOL
    monkeypatch.setenv("LDA_B200_PAIR_MAX_DOCS", "8192")
    nphi_unpaired = lda.GetState()[0]
    paired, o2 = run_pair(m, k, 4, BatchSize=b)   # this launch size now takes the paired kernels
    o2.set_state(*paired.GetState())
    t2, passes2 = paired.Transform(held, return_passes=True)
    t2_want, p2_want = o2.Transform(tuple(held), return_passes=True)
    same2 = passes2 == p2_want
    assert same2.mean() >= 0.9 and np.abs(t2 - t2_want)[:, same2].max() <= THETA_ATOL
    nphi_paired = paired.GetState()[0]
    assert np.abs(nphi_paired - nphi_unpaired).max() <= 1e-4 * max(1.0, np.abs(nphi_unpaired).max())

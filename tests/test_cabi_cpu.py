"""CPU-only checks of the drop-in boundary: liblda_b200.so loads, exports every symbol include/lda_b200.h declares,
its host-side helpers agree with the oracle, and the product refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import nlp_b200
from nlp_b200 import _lib, rand
from nlp_b200.lda import CSC, as_csc
from oracle import oracle as O
from tests.fixtures import FIXTURE_A, FIXTURE_B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


HEADERS = ("lda_b200.h", "lda_b200_index.h")


def declared_symbols():
    names = set()
    for header in HEADERS:
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(lda_b200_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 25
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "liblda_b200.so does not export %s" % n
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    _lib.lib()


def test_no_torch_types_in_the_abi():
    for header in HEADERS:
        text = open(os.path.join(ROOT, "include", header)).read()
        assert 'extern "C"' in text
        code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        assert "torch" not in code and "at::" not in code and "std::" not in code and "Tensor" not in code


def test_default_config_matches_reference_defaults():
    # NewLatentDirichletAllocation, lda.go:160-186
    c = _lib.Config()
    _lib.lib().lda_b200_default_config(7, c)
    assert (c.K, c.iterations, c.batch_size, c.burn_in_passes, c.transformation_passes) == (7, 1000, 100, 1, 500)
    assert (c.perplexity_evaluation_frequency, c.change_evaluation_frequency) == (30, 30)
    assert (c.perplexity_tolerance, c.mean_change_tolerance, c.alpha, c.eta) == (1e-2, 1e-5, 0.1, 0.01)
    assert (c.rho_phi.S, c.rho_phi.Tau, c.rho_phi.Kappa) == (10, 1000, 0.9)
    assert (c.rho_theta.S, c.rho_theta.Tau, c.rho_theta.Kappa) == (1, 10, 0.9)
    lda = nlp_b200.NewLatentDirichletAllocation(7)
    py = lda._config()
    for f, _ in _lib.Config._fields_:
        if f in ("rho_phi", "rho_theta", "processes"):
            continue
        assert getattr(py, f) == getattr(c, f), f
    assert (lda.rhoPhiT, lda.rhoThetaT, lda.wordsInCorpus) == (1.0, 1.0, 0.0)  # lda.go:182-183
    assert c.processes == lda.Processes == os.cpu_count()                      # lda.go:185: runtime.GOMAXPROCS(0)


def test_pcg_matches_oracle_and_known_answers():
    r = rand.New(rand.NewSource(0))
    assert [r.Uint64() for _ in range(4)] == [0xCBF98931523D4EEF, 0x4D98B91B8D356870, 0x01070196E695F8F1,
                                              0x703EC840C59F4493]
    for seed in (0, 1, 42, 2**63 + 5):
        a, b = rand.New(rand.NewSource(seed)), O.PCG(seed)
        assert np.array_equal(a.fill_mod(1000, 27), b.fill(1000, 27))
        assert np.array_equal(a.fill_mod(10, 10**15 + 37), b.fill(10, 10**15 + 37))
        assert a.src.state == b.state


def test_pcg_jump_ahead():
    # the device draws its stream by jumping; the jump must equal stepping
    L = _lib.lib()
    for n in (0, 1, 2, 63, 64, 1000, 123457):
        a, b = rand.NewSource(9), rand.NewSource(9)
        for _ in range(n):
            a.Uint64()
        L.lda_b200_pcg_advance(b._s, n)
        assert a.state == b.state
    c = rand.NewSource(3)
    L.lda_b200_pcg_advance(c._s, 2**40 + 12345)
    d = rand.NewSource(3)
    L.lda_b200_pcg_advance(d._s, 2**40)
    L.lda_b200_pcg_advance(d._s, 12345)
    assert c.state == d.state


def test_plan_step_equals_sequential_blends():
    # lda.go:299-317 applied n_s times == one combined update
    L = _lib.lib()
    rng = np.random.default_rng(0)
    sched = _lib.Schedule(10, 1000, 0.9)
    for n_s in (1, 2, 3, 8):
        for t0 in (1.0, 57.0, 4000.0):
            nphi = rng.random(50)
            hats = rng.random((n_s, 50)) * 100
            want = nphi.copy()
            for g in range(n_s):
                rho = O.schedule_calc(10, 1000, 0.9, t0 + g)
                want = (1 - rho) * want + rho * hats[g]
            got = np.zeros(50)
            A = C.c_double()
            for g in range(n_s):
                cg = C.c_double()
                assert L.lda_b200_plan_step(sched, t0, n_s, g, C.byref(A), C.byref(cg)) == 0
                got += cg.value * hats[g]
            got += A.value * nphi
            np.testing.assert_allclose(got, want, rtol=1e-13)


@pytest.mark.parametrize("D,b,R", [(0, 4, 2), (1, 100, 8), (1000, 100, 1), (1000, 128, 2), (1003, 100, 4), (10, 3, 8)])
def test_plan_shard_partitions_documents(D, b, R):
    L = _lib.lib()
    seen = np.zeros(D, dtype=np.int64)
    M = (D + b - 1) // b
    for rank in range(R):
        n = L.lda_b200_plan_shard(D, b, rank, R, None, None, 0)
        beg, end = np.zeros(max(n, 1), dtype=np.int64), np.zeros(max(n, 1), dtype=np.int64)
        assert L.lda_b200_plan_shard(D, b, rank, R, beg.ctypes.data, end.ctypes.data, n) == n
        for i in range(n):
            seen[beg[i]:end[i]] += 1
        if R == 1:
            assert n == 1 and (beg[0], end[0]) == (0, D)   # a single GPU owns one contiguous range
        else:
            assert n == len(range(rank, M, R))
            for i in range(n):
                assert beg[i] % b == 0 and (beg[i] // b) % R == rank    # minibatch m -> rank m % R (lda.go:524-526)
                assert end[i] == min(D, beg[i] + b)
    assert np.all(seen == 1)


def test_as_csc_token_order():
    # dense: rows ascending, exact zeros skipped (utils.go:51-57); sparse: ToCSC (lda.go:464-466)
    import scipy.sparse as sp
    d = as_csc(FIXTURE_B)
    assert (d.rows, d.cols) == (9, 9)
    assert d.indptr.tolist() == [0, 3, 6, 9, 12, 15, 16, 19, 22, 25]
    assert d.ind[9:16].tolist() == [3, 4, 5, 3, 4, 5, 3] and d.data[9:16].tolist() == [3, 3, 3, 5, 5, 5, 1]
    ref = O.to_csc(FIXTURE_B)
    for conv in (sp.csr_matrix, sp.coo_matrix, sp.dok_matrix, sp.csc_matrix, sp.lil_matrix):
        s = as_csc(conv(FIXTURE_B))
        assert np.array_equal(s.indptr, ref[2]) and np.array_equal(s.ind, ref[3]) and np.array_equal(s.data, ref[4])
    # an explicit CSC keeps its stored (possibly unsorted) order
    raw = CSC(3, 1, np.array([0, 3]), np.array([2, 0, 1]), np.array([1.0, 2.0, 3.0]))
    assert as_csc(raw).ind.tolist() == [2, 0, 1]
    neg = np.array([[0.0, -1.5], [np.nan, 0.0]])  # negative / NaN values are not rejected
    n = as_csc(neg)
    assert n.indptr.tolist() == [0, 1, 2] and n.ind.tolist() == [1, 0]


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lda = nlp_b200.NewLatentDirichletAllocation(3)
    with pytest.raises(_lib.LdaB200Error, match="no CPU fallback"):
        lda.Fit(FIXTURE_A)
    with pytest.raises(_lib.LdaB200Error):
        lda.Transform(FIXTURE_A)
    with pytest.raises(_lib.LdaB200Error, match="no CPU fallback"):
        nlp_b200.NewLinearScanIndex(nlp_b200.pairwise.CosineDistance)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nlp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "liblda_oracle", "libindex_oracle", "oracle/"):
                    assert bad not in text, (os.path.join(dirpath, f), bad)


def test_model_serialisation_round_trip():
    # Save/Load record layout (library style: dimreduction.go:111-153): uint64 K + gonum mat.Dense records
    from nlp_b200.lda import deserialise_model, serialise_model
    rng = np.random.default_rng(0)
    nphi, nz = rng.random((37, 5)), rng.random(5)
    buf = serialise_model(5, (3.0, 7.0, 1234.0), nphi, nz)
    assert len(buf) == 32 + (40 + 37 * 5 * 8) + (40 + 5 * 8)
    assert buf[:8] == (5).to_bytes(8, "little") and buf[32:36] == (0xFE000001).to_bytes(4, "little") and buf[36:39] == b"GFA"
    k, ctr, a, b = deserialise_model(buf)
    assert k == 5 and ctr == (3.0, 7.0, 1234.0) and np.array_equal(a, nphi) and np.array_equal(b, nz)
    with pytest.raises(ValueError):
        deserialise_model(buf[:100])
    with pytest.raises(ValueError):
        deserialise_model(b"\x00" * 200)


def test_bench_reference_arm_prints_one_json_line():
    # the driver's contract: exactly one JSON line on stdout; the reference arm runs on host cores only
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["impl"] == "reference" and j["cpu_baseline"]["kind"] == "port" and j["value"] > 0
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    # the reference arm runs on the CUDA arm's config: same dict, built by the same function
    sys.path.insert(0, ROOT)
    import bench
    sys.argv = ["bench.py", "--workload", "C1"]
    a = bench.parse()
    assert j["config"] == bench.common_config(a, *bench.workload(a))
    assert j["config"]["batch_size"] == 100 and "BatchSize 100" in j["cpu_baseline"]["sample"]


def test_bench_cuda_arm_refuses_to_run_without_a_gpu():
    # no CPU fallback: the CUDA arm of bench.py must stop, not time something else
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "C1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout) and not r.stdout.strip().startswith("{")


def test_matrix_descriptors_keep_the_callers_storage():
    """as_matrix hands CSR / CSC / dense inputs to the library in the storage they already have (lda_b200_matrix); other
    sparse formats and DeviceConvert = False go through the host-side ToCSC restatement (as_csc)"""
    import scipy.sparse as sp
    from nlp_b200.lda import CSR, as_matrix
    a = np.array([[0.0, 2.0, 0.0], [1.0, 0.0, 3.0]])
    want = as_csc(a)
    d, rows, cols, keep = as_matrix(a)
    assert (d.format, rows, cols, d.rows, d.cols) == (_lib.FORMAT_DENSE, 2, 3, 2, 3) and d.ptr is None and d.data == keep[0].ctypes.data
    d, rows, cols, keep = as_matrix(sp.csr_matrix(a))
    assert d.format == _lib.FORMAT_CSR and keep[0].tolist() == [0, 1, 3] and keep[1].tolist() == [1, 0, 2]
    d, _, _, keep = as_matrix(CSR(2, 3, np.array([0, 1, 3]), np.array([1, 0, 2]), np.array([2.0, 1.0, 3.0])))
    assert d.format == _lib.FORMAT_CSR and keep[2].dtype == np.float64 and keep[0].dtype == np.int64
    for m in (sp.csc_matrix(a), want, sp.coo_matrix(a), sp.dok_matrix(a)):
        d, _, _, keep = as_matrix(m)
        assert d.format == _lib.FORMAT_CSC
        assert keep[0].tolist() == want.indptr.tolist() and keep[1].tolist() == want.ind.tolist() and keep[2].tolist() == want.data.tolist()
    d, _, _, keep = as_matrix(sp.csr_matrix(a), device_convert=False)
    assert d.format == _lib.FORMAT_CSC and keep[1].tolist() == want.ind.tolist()
    with pytest.raises(ValueError):
        as_matrix(np.zeros(3))


def test_pairwise_comparers_name_the_reference_functions():
    # measures/pairwise/comparisons.go: nine comparers; the codes are the lda_b200_comparer enum of the header
    from nlp_b200.measures import pairwise
    text = open(os.path.join(ROOT, "include", "lda_b200_index.h")).read()
    enum = dict(re.findall(r"LDA_B200_([A-Z_]+) = (\d+)", text))
    assert len(pairwise.ALL) == 9
    for c in pairwise.ALL:
        snake = re.sub(r"(?<!^)(?=[A-Z])", "_", c.name).upper()
        assert int(enum[snake]) == c.code, c.name
    with pytest.raises(TypeError):
        nlp_b200.NewLinearScanIndex(lambda a, b: 0.0)


def test_model_file_fixture_layout():
    """tests/golden/lda_model_k3_w9.bin (tests/golden/make_model_fixture.py): the bytes of the model file format, field by
    field, and the Python restatement of the writer reproduces them"""
    import struct
    from nlp_b200.lda import deserialise_model, serialise_model
    from tests.golden import make_model_fixture as F
    raw = open(os.path.join(ROOT, "tests", "golden", "lda_model_k3_w9.bin"), "rb").read()
    assert raw == F.model_bytes() and len(raw) == 32 + 40 + 27 * 8 + 40 + 3 * 8
    assert struct.unpack_from("<Qddd", raw, 0) == (3, 5.0, 37.0, 1234.0)
    assert struct.unpack_from("<IBBBBqqqq", raw, 32) == (0xFE000001, ord("G"), ord("F"), ord("A"), 0, 9, 3, 0, 0)
    assert struct.unpack_from("<IBBBBqqqq", raw, 32 + 40 + 27 * 8) == (0xFE000001, ord("G"), ord("F"), ord("A"), 0, 1, 3, 0, 0)
    nphi, nz = F.arrays()
    assert serialise_model(3, F.COUNTERS, nphi, nz) == raw
    k, ctr, a, b = deserialise_model(raw)
    assert k == 3 and ctr == F.COUNTERS and np.array_equal(a, nphi) and np.array_equal(b, nz)

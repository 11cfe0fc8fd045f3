"""Host-side mirror of `nlp.LatentDirichletAllocation` (reference lda.go:60-542) over the C ABI of liblda_b200.so.

Same exported field names, constructor defaults (lda.go:160-186), method names and result orientation as the Go
type, so that parity tests read like lda_test.go.  All arithmetic of the path runs in CUDA kernels; this file only
marshals arguments (the Go shim in INTEGRATION.md does the same through cgo).
"""
import ctypes as C
from collections import namedtuple

import numpy as np

from . import _lib, rand

# term x document matrix in compressed-sparse-column form: what sparse.CSC holds after ToCSC() (lda.go:464-466)
CSC = namedtuple("CSC", "rows cols indptr ind data")
# the same matrix in compressed-sparse-row form (sparse.CSR: what the TfidfTransformer emits, weightings.go:79-95):
# indptr[rows+1], ind = document ids, data
CSR = namedtuple("CSR", "rows cols indptr ind data")


class LearningSchedule:
    """lda.go:16-32"""

    def __init__(self, S, Tau, Kappa):
        self.S, self.Tau, self.Kappa = float(S), float(Tau), float(Kappa)

    def Calc(self, iteration):
        return self.S / (self.Tau + iteration) ** self.Kappa

    def _c(self):
        return _lib.Schedule(self.S, self.Tau, self.Kappa)


def as_csc(m):
    """ToCSC() / ColNonZeroElemDo semantics (lda.go:464-466, utils.go:45-59).

    sparse input: converted to CSC (rows ascending inside a column; an input that already is CSC keeps its stored
    order); dense input: every column lists its non-zero rows in ascending order, exact zeros skipped."""
    if isinstance(m, CSC):
        rows, cols, indptr, ind, data = m
    elif hasattr(m, "tocsc"):
        s = m if getattr(m, "format", None) == "csc" else m.tocsc()
        rows, cols = s.shape
        indptr, ind, data = s.indptr, s.indices, s.data
    else:
        a = np.asarray(m, dtype=np.float64)
        if a.ndim != 2:
            raise ValueError("nlp: matrix must be 2-dimensional")
        rows, cols = a.shape
        at = a.T
        jj, ii = np.nonzero(at)
        indptr = np.concatenate(([0], np.cumsum(np.bincount(jj, minlength=cols)))) if cols else np.zeros(1)
        ind, data = ii, at[jj, ii]
    return CSC(int(rows), int(cols), np.ascontiguousarray(indptr, dtype=np.int64),
               np.ascontiguousarray(ind, dtype=np.int64), np.ascontiguousarray(data, dtype=np.float64))


def _ptr(a):
    return None if a is None else a.ctypes.data


def as_matrix(m, device_convert=True):
    """lda_b200_matrix descriptor of m in the storage it already has (CSC / CSR / dense), so that ToCSC() and the dense
    scan run on the GPU; other sparse formats (COO, DOK, ...) and device_convert=False go through as_csc on the host.
    Returns (descriptor, rows, cols, arrays kept alive for the duration of the call)."""
    fmt = None
    if isinstance(m, CSR):
        fmt, (rows, cols, ptr, ind, data) = _lib.FORMAT_CSR, m
    elif isinstance(m, CSC):
        fmt, (rows, cols, ptr, ind, data) = _lib.FORMAT_CSC, m
    elif device_convert and getattr(m, "format", None) in ("csr", "csc"):
        fmt = _lib.FORMAT_CSR if m.format == "csr" else _lib.FORMAT_CSC
        (rows, cols), ptr, ind, data = m.shape, m.indptr, m.indices, m.data
    elif device_convert and not hasattr(m, "tocsc"):
        a = np.ascontiguousarray(m, dtype=np.float64)
        if a.ndim != 2:
            raise ValueError("nlp: matrix must be 2-dimensional")
        d = _lib.Matrix(_lib.FORMAT_DENSE, a.shape[0], a.shape[1], None, None, _ptr(a))
        return d, int(a.shape[0]), int(a.shape[1]), (a,)
    if fmt is None:
        fmt, (rows, cols, ptr, ind, data) = _lib.FORMAT_CSC, as_csc(m)
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    ind = np.ascontiguousarray(ind, dtype=np.int64)
    data = np.ascontiguousarray(data, dtype=np.float64)
    return _lib.Matrix(fmt, int(rows), int(cols), _ptr(ptr), _ptr(ind), _ptr(data)), int(rows), int(cols), (ptr, ind, data)


class LatentDirichletAllocation:
    """lda.go:68-141.  Use NewLatentDirichletAllocation(k)."""

    def __init__(self, k, device=0):
        # defaults: lda.go:160-186
        self.Iterations = 1000
        self.PerplexityTolerance = 1e-2
        self.PerplexityEvaluationFrequency = 30
        self.BatchSize = 100
        self.K = int(k)
        self.BurnInPasses = 1
        self.TransformationPasses = 500
        self.MeanChangeTolerance = 1e-5
        self.ChangeEvaluationFrequency = 30
        self.Alpha = 0.1
        self.Eta = 0.01
        self.RhoPhi = LearningSchedule(10, 1000, 0.9)
        self.RhoTheta = LearningSchedule(1, 10, 0.9)
        self.rhoPhiT = 1.0
        self.rhoThetaT = 1.0
        self.wordsInCorpus = 0.0
        self.w = self.d = 0
        import time
        self.Rnd = rand.New(rand.NewSource(time.time_ns() & 0xFFFFFFFFFFFFFFFF))
        # Processes (lda.go:134) = minibatches in flight per step.  Default as in lda.go:185 -- runtime.GOMAXPROCS(0), the
        # host's hardware threads -- and at least one per GPU of the torch.distributed job; the GPUs share them
        # (Processes / GPUs consecutive minibatches per launch, at most 256).  Processes = 1 is the deterministic parity mode.
        import os
        self.Processes = max(_world()[1], os.cpu_count() or 1)
        # --- not in the reference ---
        self.Device = device
        self.DeviceInit = True       # draw the initial nPhi / nTheta from Rnd on the GPU (bit-identical stream)
        self.DeviceConvert = True    # CSR -> CSC (ToCSC, lda.go:464-466) and the dense scan (utils.go:51-57) on the GPU
        self.Deterministic = False   # bitwise run-to-run reproducible fits (fixed-point nPhiHat, lda_b200_set_deterministic)
        self.PerplexityTrace = []    # perplexities evaluated inside the last fit (lda.go:530-539 hides them)
        self.IterationsRun = 0
        self._h = None
        self._comm_ranks = 1
        self._has_model = False

    # ---- handle management ------------------------------------------------------------------------------
    def _config(self):
        c = _lib.Config()
        c.K = self.K
        c.iterations = self.Iterations
        c.batch_size = self.BatchSize
        c.burn_in_passes = self.BurnInPasses
        c.transformation_passes = self.TransformationPasses
        c.perplexity_evaluation_frequency = self.PerplexityEvaluationFrequency
        c.change_evaluation_frequency = self.ChangeEvaluationFrequency
        c.processes = self.Processes
        c.perplexity_tolerance = self.PerplexityTolerance
        c.mean_change_tolerance = self.MeanChangeTolerance
        c.alpha, c.eta = self.Alpha, self.Eta
        c.rho_phi, c.rho_theta = self.RhoPhi._c(), self.RhoTheta._c()
        return c

    def _handle(self):
        L = _lib.lib()
        cfg = self._config()
        if self._h is None:
            h = C.c_void_p()
            _lib.check(L.lda_b200_create(cfg, self.Device, C.byref(h)))
            self._h = h
        else:
            _lib.check(L.lda_b200_set_config(self._h, cfg), self._h)
        if getattr(self, "_det_set", None) != bool(self.Deterministic):
            _lib.check(L.lda_b200_set_deterministic(self._h, 1 if self.Deterministic else 0), self._h)
            self._det_set = bool(self.Deterministic)
        want = max(1, min(self.Processes, _world()[1]))
        if want != self._comm_ranks:
            self._init_comm(want)
        return self._h

    def _init_comm(self, nranks):
        """one process per GPU: rank 0 creates the ncclUniqueId, torch.distributed carries it to the others"""
        L = _lib.lib()
        rank, world = _world()
        if nranks > 1:
            import torch
            import torch.distributed as dist
            buf = (C.c_ubyte * 128)()
            if rank == 0:
                _lib.check(L.lda_b200_comm_unique_id(buf))
            t = torch.tensor(list(buf), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda(self.Device)
            dist.broadcast(t, src=0)
            raw = bytes(t.cpu().tolist())
            _lib.check(L.lda_b200_comm_init(self._h, rank, nranks, raw), self._h)
        else:
            _lib.check(L.lda_b200_comm_init(self._h, 0, 1, None), self._h)
        self._comm_ranks = nranks

    def CommInfo(self):
        """(rank, ranks, path) of the handle's communicator; path = 2 when the multi-GPU step runs as one reduce + M-step
        kernel with the reduction and the broadcast done in the NVSwitch (NVLink multicast), 1 when that kernel uses
        unicast peer loads / stores, 0 for ncclAllReduce + blend (truthy = one kernel over NVLink)"""
        r, n, p = C.c_int32(0), C.c_int32(1), C.c_int32(0)
        if self._h is not None:
            _lib.lib().lda_b200_comm_info(self._h, C.byref(r), C.byref(n), C.byref(p))
        return r.value, n.value, p.value

    def ToCSC(self, m):
        """sparse.ToCSC() (lda.go:464-466) of a scipy CSR term x document matrix on the GPU -> nlp_b200.CSC"""
        csr = m.tocsr() if getattr(m, "format", None) != "csr" else m
        csr.sort_indices()
        rows, cols = csr.shape
        rp = np.ascontiguousarray(csr.indptr, dtype=np.int64)
        ci = np.ascontiguousarray(csr.indices, dtype=np.int64)
        dv = np.ascontiguousarray(csr.data, dtype=np.float64)
        indptr = np.zeros(cols + 1, dtype=np.int64)
        ind = np.zeros(ci.size, dtype=np.int64)
        data = np.zeros(ci.size, dtype=np.float64)
        h = self._handle()
        _lib.check(_lib.lib().lda_b200_csr_to_csc(h, rows, cols, _ptr(rp), _ptr(ci), _ptr(dv), _ptr(indptr), _ptr(ind),
                                                  _ptr(data)), h)
        return CSC(rows, cols, indptr, ind, data)

    def Close(self):
        if self._h is not None:
            _lib.lib().lda_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass

    def _counters(self):
        return _lib.Counters(self.rhoPhiT, self.rhoThetaT, self.wordsInCorpus)

    def _take(self, ctr):
        self.rhoPhiT, self.rhoThetaT, self.wordsInCorpus = ctr.rho_phi_t, ctr.rho_theta_t, ctr.words_in_corpus

    def _draws(self, rows, n_draws_mod):
        """host-side draws (DeviceInit = False): lda.go:199-205 / 472-475 / 422-426"""
        return self.Rnd.fill_mod(rows * self.K, n_draws_mod)

    # ---- reference API ------------------------------------------------------------------------------------
    def Fit(self, m):
        """lda.go:211-214"""
        self.FitTransform(m, _want_theta=False)
        return self

    def FitTransform(self, m, _want_theta=True, nphi0=None, ntheta0=None, out=None):
        """lda.go:463-542.  Returns the K x D doc-topic matrix (float64).  nphi0 / ntheta0 (not in the reference)
        replace the RNG draws with an exported init, e.g. the reference's own, for deterministic parity runs; out
        is an optional caller-owned K x D float64 result buffer (e.g. page-locked)."""
        desc, rows, cols, _keep = as_matrix(m, self.DeviceConvert)
        h = self._handle()
        L = _lib.lib()
        if not self.DeviceInit:
            if nphi0 is None:
                nphi0 = self._draws(rows, rows * self.K)
            if ntheta0 is None:
                ntheta0 = self._draws(cols, cols * self.K)
        nphi0 = None if nphi0 is None else np.ascontiguousarray(nphi0, dtype=np.float64)
        ntheta0 = None if ntheta0 is None else np.ascontiguousarray(ntheta0, dtype=np.float64)
        if nphi0 is not None and nphi0.size != rows * self.K:
            raise ValueError("nlp: nphi0 must hold W*K values")
        if ntheta0 is not None and ntheta0.size != cols * self.K:
            raise ValueError("nlp: ntheta0 must hold D*K values")
        theta = None
        if _want_theta:
            theta = np.zeros((self.K, cols), dtype=np.float64) if out is None else out
            if theta.shape != (self.K, cols) or theta.dtype != np.float64 or not theta.flags.c_contiguous:
                raise ValueError("nlp: out must be a C-contiguous float64 K x D array")
        cap = self.Iterations // self.PerplexityEvaluationFrequency + 1 if self.PerplexityEvaluationFrequency > 0 else 1
        trace = np.zeros(cap, dtype=np.float64)
        n_evals, iters = C.c_int32(0), C.c_int32(0)
        ctr = self._counters()
        rc = L.lda_b200_fit_matrix(h, desc, _ptr(nphi0), _ptr(ntheta0), self.Rnd.src._s, ctr, _ptr(theta), _ptr(trace), cap,
                                   C.byref(n_evals), C.byref(iters))
        _lib.check(rc, h)
        self.w, self.d = rows, cols
        self._take(ctr)
        self.PerplexityTrace = trace[:n_evals.value].tolist()
        self.IterationsRun = iters.value
        self._has_model = True
        return theta

    # ---- resident-corpus stepping (not in the reference; lda_b200_fit is exactly these three calls) -------------
    def FitBegin(self, m, nphi0=None, ntheta0=None):
        """ingest + init of lda.go:464-497; the corpus stays in HBM until FitEnd"""
        csc = as_csc(m)
        h = self._handle()
        if not self.DeviceInit:
            nphi0 = self._draws(csc.rows, csc.rows * self.K) if nphi0 is None else nphi0
            ntheta0 = self._draws(csc.cols, csc.cols * self.K) if ntheta0 is None else ntheta0
        nphi0 = None if nphi0 is None else np.ascontiguousarray(nphi0, dtype=np.float64)
        ntheta0 = None if ntheta0 is None else np.ascontiguousarray(ntheta0, dtype=np.float64)
        ctr = self._counters()
        rc = _lib.lib().lda_b200_fit_begin(h, csc.rows, csc.cols, _ptr(csc.indptr), _ptr(csc.ind), _ptr(csc.data),
                                           _ptr(nphi0), _ptr(ntheta0), self.Rnd.src._s, ctr)
        _lib.check(rc, h)
        self.w, self.d = csc.rows, csc.cols
        self._take(ctr)
        self._has_model = True
        self.PerplexityTrace = []
        self.IterationsRun = 0

    def FitIterate(self, n):
        """n more iterations of the loop at lda.go:501-540; returns (iterations run, stopped, device ms)"""
        h = self._handle()
        cap = max(1, n)
        trace = np.zeros(cap, dtype=np.float64)
        n_evals, iters, stopped, ms = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_float(0)
        ctr = self._counters()
        rc = _lib.lib().lda_b200_fit_iterate(h, n, ctr, _ptr(trace), cap, C.byref(n_evals), C.byref(iters),
                                             C.byref(stopped), C.byref(ms))
        _lib.check(rc, h)
        self._take(ctr)
        self.PerplexityTrace += trace[:n_evals.value].tolist()
        self.IterationsRun += iters.value
        return iters.value, bool(stopped.value), ms.value

    def FitEnd(self, want_theta=True):
        """lda.go:541: normalised, transposed K x D copy; releases the resident corpus"""
        h = self._handle()
        theta = np.zeros((self.K, self.d), dtype=np.float64) if want_theta else None
        _lib.check(_lib.lib().lda_b200_fit_end(h, _ptr(theta)), h)
        return theta

    def PartialFit(self, m, iterations=1, ntheta0=None, want_theta=False):
        """OnlineTransformer.PartialFit (vectorisers.go:36-39; TODO at lda.go:147): `iterations` SCVB0 iterations over the
        documents of m, continuing from the current model (a fresh random model on first use).  Returns self, or
        the K x D doc-topic matrix of these documents when want_theta is set."""
        desc, rows, cols, _keep = as_matrix(m, self.DeviceConvert)
        h = self._handle()
        if ntheta0 is None and not self.DeviceInit:
            if self.IterationsRun == 0 and not self._has_model:
                raise ValueError("nlp: DeviceInit=False PartialFit needs an existing model")
            ntheta0 = self._draws(cols, cols * self.K)
        ntheta0 = None if ntheta0 is None else np.ascontiguousarray(ntheta0, dtype=np.float64)
        theta = np.zeros((self.K, cols), dtype=np.float64) if want_theta else None
        ctr = self._counters()
        rc = _lib.lib().lda_b200_partial_fit_matrix(h, desc, _ptr(ntheta0), self.Rnd.src._s, ctr, iterations, _ptr(theta))
        _lib.check(rc, h)
        self.w, self.d = rows, cols
        self._take(ctr)
        self._has_model = True
        return theta if want_theta else self

    # ---- persistence in the library's binary style (weightings.go:97-116, dimreduction.go:111-153) ---------------
    def Save(self, w):
        """uint64 K, float64 rhoPhiT / rhoThetaT / wordsInCorpus (little endian), then nPhi (W x K) and nZ (1 x K) as
        gonum `mat.Dense.MarshalBinaryTo` records.  The bytes are produced by the library (lda_b200_save), so the Go,
        C++ and Python hosts write the same file; serialise_model below restates the layout for the tests."""
        h = self._handle()
        L = _lib.lib()
        n = C.c_int64(0)
        _lib.check(L.lda_b200_save_size(h, C.byref(n)), h)
        buf = np.empty(n.value, dtype=np.uint8)
        _lib.check(L.lda_b200_save(h, self._counters(), _ptr(buf), n.value, C.byref(n)), h)
        w.write(buf[:n.value].tobytes())

    def Load(self, r):
        raw = r.read()
        if len(raw) < 8:
            raise ValueError("nlp: unexpected EOF")  # io.ErrUnexpectedEOF in dimreduction.go:135
        buf = np.frombuffer(raw, dtype=np.uint8)
        h = self._handle()
        ctr = self._counters()
        _lib.check(_lib.lib().lda_b200_load(h, _ptr(buf), buf.size, ctr), h)
        self.w, self.K = self._dims_raw()
        self._take(ctr)
        self._has_model = True

    def Transform(self, m, theta0=None, return_passes=False):
        """lda.go:446-453: K x D doc-topic matrix for new documents under the fitted model"""
        desc, _rows, cols, _keep = as_matrix(m, self.DeviceConvert)
        h = self._handle()
        if theta0 is None and not self.DeviceInit:
            theta0 = self._draws(cols, cols * self.K)
        theta0 = None if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float64)
        theta = np.zeros((self.K, cols), dtype=np.float64)
        passes = np.zeros(cols, dtype=np.int32) if return_passes else None
        rc = _lib.lib().lda_b200_transform_matrix(h, desc, _ptr(theta0), self.Rnd.src._s, self.rhoThetaT, _ptr(theta),
                                                  _ptr(passes))
        _lib.check(rc, h)
        return (theta, passes) if return_passes else theta

    def Perplexity(self, m, theta0=None):
        """lda.go:368-389"""
        desc, _rows, cols, _keep = as_matrix(m, self.DeviceConvert)
        h = self._handle()
        if theta0 is None and not self.DeviceInit:
            theta0 = self._draws(cols, cols * self.K)
        theta0 = None if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float64)
        out = C.c_double(0)
        rc = _lib.lib().lda_b200_perplexity_matrix(h, desc, _ptr(theta0), self.Rnd.src._s, self.rhoThetaT, C.byref(out))
        _lib.check(rc, h)
        return out.value

    def _dims(self):
        self._handle()
        return self._dims_raw()

    def _dims_raw(self):
        w, k = C.c_int64(0), C.c_int32(0)
        _lib.check(_lib.lib().lda_b200_get_dims(self._h, C.byref(w), C.byref(k)), self._h)
        return int(w.value), int(k.value)

    def Components(self):
        """lda.go:414-416: K x W topic-word matrix"""
        h = self._handle()
        self.w = self._dims()[0]
        out = np.zeros((self.K, self.w), dtype=np.float64)
        _lib.check(_lib.lib().lda_b200_components(h, _ptr(out)), h)
        return out

    # ---- persistence (TODO in the reference, lda.go:158) ------------------------------------------------------
    def GetState(self):
        h = self._handle()
        self.w = self._dims()[0]
        nphi = np.zeros((self.w, self.K), dtype=np.float64)
        nz = np.zeros(self.K, dtype=np.float64)
        _lib.check(_lib.lib().lda_b200_get_state(h, _ptr(nphi), _ptr(nz)), h)
        return nphi, nz

    def SetState(self, nphi, nz):
        nphi = np.ascontiguousarray(nphi, dtype=np.float64)
        nz = np.ascontiguousarray(nz, dtype=np.float64)
        if nphi.ndim != 2 or nphi.shape[1] != self.K or nz.shape != (self.K,):
            raise ValueError("nlp: state must be nPhi[W,K], nZ[K]")
        h = self._handle()
        _lib.check(_lib.lib().lda_b200_set_state(h, nphi.shape[0], _ptr(nphi), _ptr(nz)), h)
        self.w = nphi.shape[0]


_GONUM_VERSION = 0xFE000001  # gonum.org/v1/gonum/mat io.go storage header: version, 'G','F','A', unit, rows, cols, ku, kl


def _dense_record(a):
    import struct
    a = np.ascontiguousarray(a, dtype="<f8")
    return struct.pack("<IBBBBqqqq", _GONUM_VERSION, ord("G"), ord("F"), ord("A"), 0, a.shape[0], a.shape[1], 0, 0) + a.tobytes()


def _read_dense(buf, off):
    import struct
    ver, form, packing, uplo, _unit, rows, cols, _ku, _kl = struct.unpack_from("<IBBBBqqqq", buf, off)
    if ver != _GONUM_VERSION or (form, packing, uplo) != (ord("G"), ord("F"), ord("A")):
        raise ValueError("nlp: not a gonum mat.Dense record")
    off += 40
    n = rows * cols
    if rows < 0 or cols < 0 or off + 8 * n > len(buf):
        raise ValueError("nlp: truncated mat.Dense record")
    return np.frombuffer(buf, dtype="<f8", count=n, offset=off).reshape(rows, cols).copy(), off + 8 * n


def serialise_model(k, counters, nphi, nz):
    import struct
    return struct.pack("<Qddd", k, *counters) + _dense_record(nphi) + _dense_record(np.asarray(nz).reshape(1, -1))


def deserialise_model(buf):
    import struct
    if len(buf) < 32:
        raise ValueError("nlp: unexpected EOF")  # io.ErrUnexpectedEOF in dimreduction.go:135
    k, a, b, c = struct.unpack_from("<Qddd", buf, 0)
    nphi, off = _read_dense(buf, 32)
    nz, off = _read_dense(buf, off)
    if nphi.shape[1] != k or nz.shape != (1, k):
        raise ValueError("nlp: model dimensions do not match K")
    return int(k), (a, b, c), nphi, nz[0]


def NewLatentDirichletAllocation(k, device=None):
    """lda.go:145"""
    if device is None:
        import os
        device = int(os.environ.get("LOCAL_RANK", "0")) if _world()[1] > 1 else 0
    return LatentDirichletAllocation(k, device=device)


def _world():
    """(rank, world size) of the torch.distributed job this process belongs to, (0, 1) outside one"""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1

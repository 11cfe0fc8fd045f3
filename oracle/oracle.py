"""ctypes wrapper around oracle/lda_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It mirrors the reference's `nlp.LatentDirichletAllocation` surface
(lda.go:68-141, 211, 368, 414, 446, 463) closely enough that parity tests read like
lda_test.go.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liblda_oracle.so")
    src = os.path.join(_HERE, "lda_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liblda_oracle.so"], stdout=subprocess.DEVNULL)
    return so


class _CSC(C.Structure):
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int64), ("indptr", C.c_void_p), ("ind", C.c_void_p),
                ("data", C.c_void_p)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_lda_new.restype = C.c_void_p
        L.oracle_lda_new.argtypes = [C.c_int64, C.c_uint64, C.c_int64]
        L.oracle_lda_free.argtypes = [C.c_void_p]
        L.oracle_lda_fit_transform.argtypes = [C.c_void_p, C.POINTER(_CSC), C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_lda_transform.argtypes = [C.c_void_p, C.POINTER(_CSC), C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_lda_perplexity.restype = C.c_double
        L.oracle_lda_perplexity.argtypes = [C.c_void_p, C.POINTER(_CSC), C.c_void_p]
        L.oracle_lda_components.argtypes = [C.c_void_p, C.c_void_p]
        for name, t in _FIELDS.items():
            getattr(L, "oracle_lda_set_" + name).argtypes = [C.c_void_p, t]
            getattr(L, "oracle_lda_get_" + name).argtypes = [C.c_void_p]
            getattr(L, "oracle_lda_get_" + name).restype = t
        for name in ("RhoPhi", "RhoTheta"):
            getattr(L, "oracle_lda_set_" + name).argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.oracle_lda_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_lda_get_rnd.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        for name in ("w", "d", "iterationsRun", "tokenVisits"):
            getattr(L, "oracle_lda_get_" + name).argtypes = [C.c_void_p]
            getattr(L, "oracle_lda_get_" + name).restype = C.c_int64
        L.oracle_lda_get_perpTrace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.oracle_lda_get_perpTrace.restype = C.c_int64
        L.oracle_lda_get_iterTime.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.oracle_lda_get_iterTime.restype = C.c_int64
        L.oracle_lda_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_lda_set_state.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.oracle_pcg_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_pcg_uint64.argtypes = [C.c_void_p]
        L.oracle_pcg_uint64.restype = C.c_uint64
        L.oracle_pcg_fill.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.oracle_schedule_calc.argtypes = [C.c_double] * 4
        L.oracle_schedule_calc.restype = C.c_double
        _LIB = L
    return _LIB


_FIELDS = {
    "Iterations": C.c_int64, "PerplexityTolerance": C.c_double, "PerplexityEvaluationFrequency": C.c_int64,
    "BatchSize": C.c_int64, "K": C.c_int64, "BurnInPasses": C.c_int64, "TransformationPasses": C.c_int64,
    "MeanChangeTolerance": C.c_double, "ChangeEvaluationFrequency": C.c_int64, "Alpha": C.c_double,
    "Eta": C.c_double, "Processes": C.c_int64, "SyncGroup": C.c_int64, "HoistPow": C.c_int64,
    "rhoPhiT": C.c_double, "rhoThetaT": C.c_double,
    "wordsInCorpus": C.c_double,
}


class PCG:
    """golang.org/x/exp/rand PCGSource restatement (state = (hi, lo))."""

    def __init__(self, seed=0):
        self._s = (C.c_uint64 * 2)()
        lib().oracle_pcg_seed(self._s, seed)

    @property
    def state(self):
        return int(self._s[0]), int(self._s[1])

    @state.setter
    def state(self, hl):
        self._s[0], self._s[1] = hl

    def uint64(self):
        return int(lib().oracle_pcg_uint64(self._s))

    def int(self):
        return self.uint64() & ((1 << 63) - 1)

    def fill(self, count, n):
        out = np.empty(count, dtype=np.float64)
        lib().oracle_pcg_fill(self._s, count, n, out.ctypes.data)
        return out


def to_csc(m):
    """Term x document input -> (rows, cols, indptr, ind, data) in the token order of
    utils.go:45-59: dense -> ascending rows, exact zeros skipped; scipy sparse -> CSC with
    rows ascending inside a column (what sparse.ToCSC() of a DOK/CSR/COO produces)."""
    if isinstance(m, tuple):
        rows, cols, indptr, ind, data = m
    elif hasattr(m, "tocsc"):
        s = m.tocsc()
        s.sort_indices()
        rows, cols = s.shape
        indptr, ind, data = s.indptr, s.indices, s.data
    else:
        a = np.asarray(m, dtype=np.float64)
        rows, cols = a.shape
        at = a.T
        jj, ii = np.nonzero(at)
        indptr = np.zeros(cols + 1, dtype=np.int64)
        np.add.at(indptr, jj + 1, 1)
        indptr = np.cumsum(indptr)
        ind, data = ii, at[jj, ii]
    return (int(rows), int(cols), np.ascontiguousarray(indptr, dtype=np.int64),
            np.ascontiguousarray(ind, dtype=np.int64), np.ascontiguousarray(data, dtype=np.float64))


def _csc_struct(m):
    rows, cols, indptr, ind, data = to_csc(m)
    s = _CSC(rows, cols, indptr.ctypes.data, ind.ctypes.data, data.ctypes.data)
    s._keep = (indptr, ind, data)
    return s


def _opt(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data


class LatentDirichletAllocation:
    """Oracle with the reference's field names; attribute access proxies to the C struct."""

    def __init__(self, k, seed=0, processes=1):
        object.__setattr__(self, "_h", lib().oracle_lda_new(k, seed, processes))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().oracle_lda_free(self._h)
                object.__setattr__(self, "_h", None)
        except Exception:   # interpreter shutdown: module globals are already gone
            pass

    def __getattr__(self, name):
        if name in _FIELDS or name in ("w", "d", "iterationsRun", "tokenVisits"):
            return getattr(lib(), "oracle_lda_get_" + name)(self._h)
        raise AttributeError(name)

    def __setattr__(self, name, v):
        if name in _FIELDS:
            getattr(lib(), "oracle_lda_set_" + name)(self._h, v)
        elif name in ("RhoPhi", "RhoTheta"):
            getattr(lib(), "oracle_lda_set_" + name)(self._h, *v)
        else:
            raise AttributeError(name)

    def seed(self, s):
        lib().oracle_lda_seed(self._h, s)

    @property
    def rnd_state(self):
        hi, lo = C.c_uint64(), C.c_uint64()
        lib().oracle_lda_get_rnd(self._h, C.byref(hi), C.byref(lo))
        return hi.value, lo.value

    @property
    def perplexity_trace(self):
        n = lib().oracle_lda_get_perpTrace(self._h, None, 0)
        out = np.empty(n, dtype=np.float64)
        lib().oracle_lda_get_perpTrace(self._h, out.ctypes.data, n)
        return out

    @property
    def iteration_times(self):
        """seconds since the start of the iteration loop at the end of each iteration of the last fit"""
        n = lib().oracle_lda_get_iterTime(self._h, None, 0)
        out = np.empty(n, dtype=np.float64)
        lib().oracle_lda_get_iterTime(self._h, out.ctypes.data, n)
        return out

    def FitTransform(self, m, nphi0=None, ntheta0=None):
        s = _csc_struct(m)
        out = np.empty((self.K, s.cols), dtype=np.float64)
        a, pa = _opt(nphi0)
        b, pb = _opt(ntheta0)
        lib().oracle_lda_fit_transform(self._h, C.byref(s), pa, pb, out.ctypes.data)
        return out

    def Fit(self, m, **kw):
        self.FitTransform(m, **kw)
        return self

    def Transform(self, m, theta0=None, return_passes=False):
        s = _csc_struct(m)
        out = np.empty((self.K, s.cols), dtype=np.float64)
        passes = np.empty(s.cols, dtype=np.int64)
        a, pa = _opt(theta0)
        lib().oracle_lda_transform(self._h, C.byref(s), pa, out.ctypes.data, passes.ctypes.data)
        return (out, passes) if return_passes else out

    def Perplexity(self, m, theta0=None):
        s = _csc_struct(m)
        a, pa = _opt(theta0)
        return float(lib().oracle_lda_perplexity(self._h, C.byref(s), pa))

    def Components(self):
        out = np.empty((self.K, self.w), dtype=np.float64)
        lib().oracle_lda_components(self._h, out.ctypes.data)
        return out

    def get_state(self):
        nphi = np.empty((self.w, self.K), dtype=np.float64)
        nz = np.empty(self.K, dtype=np.float64)
        lib().oracle_lda_get_state(self._h, nphi.ctypes.data, nz.ctypes.data)
        return nphi, nz

    def set_state(self, nphi, nz):
        nphi = np.ascontiguousarray(nphi, dtype=np.float64)
        nz = np.ascontiguousarray(nz, dtype=np.float64)
        lib().oracle_lda_set_state(self._h, nphi.shape[0], nphi.ctypes.data, nz.ctypes.data)


def schedule_calc(S, Tau, Kappa, iteration):
    return float(lib().oracle_schedule_calc(S, Tau, Kappa, iteration))

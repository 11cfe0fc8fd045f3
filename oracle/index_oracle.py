"""ctypes wrapper around oracle/index_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Mirrors `nlp.LinearScanIndex` (index.go:60-135) and the `pairwise` comparers (measures/pairwise/comparisons.go) closely
enough that tests read like index_test.go.  Only tests/ and tools/bench_search.py's cpu_baseline leg may import it.
"""
import ctypes as C
import os
import subprocess
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

Match = namedtuple("Match", "Distance ID")   # index.go:13-19

COMPARERS = ["CosineSimilarity", "CosineDistance", "AngularDistance", "AngularSimilarity", "HammingDistance",
             "HammingSimilarity", "EuclideanDistance", "ManhattenDistance", "VectorLenSimilarity"]


def build(force=False):
    so = os.path.join(_HERE, "libindex_oracle.so")
    src = os.path.join(_HERE, "index_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libindex_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_compare.restype = C.c_double
        L.oracle_compare.argtypes = [C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.oracle_linear_scan_search.restype = C.c_int64
        L.oracle_linear_scan_search.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                                C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def compare(name, a, b):
    """pairwise.<name>(a, b)"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape and a.ndim == 1
    return lib().oracle_compare(COMPARERS.index(name), a.size, a.ctypes.data, 1, b.ctypes.data, 1)


class LinearScanIndex:
    """index.go:60-135 over integer ids"""

    def __init__(self, comparer):
        self.comparer = COMPARERS.index(comparer)
        self.signatures, self.ids = [], []

    def Index(self, v, id):      # index.go:79-84
        self.signatures.append(np.array(v, dtype=np.float64))
        self.ids.append(int(id))

    def Remove(self, id):        # index.go:117-135: the first vector with that id
        if id in self.ids:
            i = self.ids.index(id)
            del self.signatures[i], self.ids[i]

    def Search(self, q, k):      # index.go:85-115, matches in the reference's (heap) order
        n = len(self.ids)
        q = np.ascontiguousarray(q, dtype=np.float64)
        m = np.ascontiguousarray(np.stack(self.signatures, axis=1)) if n else np.zeros((q.size, 0))
        ids = np.asarray(self.ids, dtype=np.int64)
        ids_out = np.zeros(max(k, 1), dtype=np.int64)
        dist_out = np.zeros(max(k, 1), dtype=np.float64)
        got = lib().oracle_linear_scan_search(self.comparer, q.size, n, m.ctypes.data, max(n, 1), ids.ctypes.data, q.ctypes.data, 1,
                                              k, ids_out.ctypes.data, dist_out.ctypes.data)
        return [Match(float(dist_out[i]), int(ids_out[i])) for i in range(got)]

"""Host-side mirror of `nlp.LinearScanIndex` (reference index.go:60-135) over the C ABI of liblda_b200.so
(include/lda_b200_index.h): exact nearest-neighbour search over document vectors -- e.g. the K x D doc-topic matrix of
the LDA path -- by a linear scan on the GPU.  Same method names, argument meaning and result type (`Match`) as the Go
type; `IndexColumns` / `SearchColumns` are the batched forms (`ColDo(m, index.Index)`, many queries per scan).  All
arithmetic runs in CUDA kernels; this file only marshals arguments and maps arbitrary Python ids to int64 handles."""
import ctypes as C
from collections import namedtuple

import numpy as np

from . import _lib
from .measures import pairwise

Match = namedtuple("Match", "Distance ID")   # index.go:13-19


def _ptr(a):
    return None if a is None else a.ctypes.data


class LinearScanIndex:
    """index.go:60-135.  Use NewLinearScanIndex(compareFN)."""

    def __init__(self, compareFN, device=0):
        if not isinstance(compareFN, pairwise.Comparer):
            raise TypeError("nlp: compareFN must be one of nlp_b200.measures.pairwise (the kernels are specialised on it)")
        self.distance = compareFN
        # Vectors are known to the library by int64 handles (their insertion numbers).  Caller-chosen ids (index.go:16:
        # ID interface{}) are kept in dictionaries; runs of default ids (consecutive integers) are kept as ranges, so that
        # indexing a million columns costs no per-vector Python work.
        self._ids = {}        # handle -> caller's id
        self._handles = {}    # caller's id -> handles in insertion order (Remove drops the first, index.go:121-122)
        self._auto = []       # (first handle, end handle, first id) of runs with default ids, in insertion order
        self._removed = set() # handles of removed default-id vectors
        self._next = 0
        self._next_default_id = 0   # default ids are never reused, whatever has been removed since
        self._dim = None      # fixed by the first vector
        h = C.c_void_p()
        _lib.check_index(_lib.lib().lda_b200_index_create(device, compareFN.code, C.byref(h)))
        self._h = h

    def Close(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib().lda_b200_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass

    def __len__(self):
        return int(_lib.lib().lda_b200_index_len(self._h))

    def _new_handles(self, n):
        out = np.arange(self._next, self._next + n, dtype=np.int64)
        self._next += n
        return out

    def _register(self, handles, ids):
        for hd, id in zip(handles.tolist(), ids):
            self._ids[hd] = id
            self._handles.setdefault(id, []).append(hd)

    def _register_auto(self, handles, first_id):
        if len(handles):
            self._auto.append((int(handles[0]), int(handles[-1]) + 1, int(first_id)))

    def _id_of(self, hd):
        if hd in self._ids:
            return self._ids[hd]
        import bisect
        i = bisect.bisect_right(self._auto, (hd, float("inf"), 0)) - 1
        h0, _h1, id0 = self._auto[i]
        return id0 + (hd - h0)

    def _first_handle_of(self, id):
        """the earliest-indexed live vector carrying this id, or None"""
        best = self._handles[id][0] if self._handles.get(id) else None
        if isinstance(id, (int, np.integer)) and not isinstance(id, bool):
            for h0, h1, id0 in self._auto:
                hd = h0 + (int(id) - id0)
                if h0 <= hd < h1 and hd not in self._removed and (best is None or hd < best):
                    best = hd
                    break
        return best

    # ---- reference API -----------------------------------------------------------------------------------
    def Index(self, v, id):
        """index.go:79-84"""
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
        handles = self._new_handles(1)
        _lib.check_index(_lib.lib().lda_b200_index_add(self._h, v.size, 1, _ptr(v), 1, _ptr(handles)), self._h)
        self._dim = v.size
        self._register(handles, [id])

    def Search(self, qv, k):
        """index.go:85-115: up to k Matches, nearest first (the reference returns the same matches in heap order)"""
        qv = np.ascontiguousarray(np.asarray(qv, dtype=np.float64).ravel())
        return self.SearchColumns(qv.reshape(-1, 1), k)[0]

    def Remove(self, id):
        """index.go:117-135"""
        hd = self._first_handle_of(id)
        if hd is None:
            return
        if hd in self._ids:
            hs = self._handles[id]
            hs.remove(hd)
            if not hs:
                del self._handles[id]
            del self._ids[hd]
        else:
            self._removed.add(hd)
        _lib.check_index(_lib.lib().lda_b200_index_remove(self._h, hd), self._h)

    # ---- batched forms -------------------------------------------------------------------------------------
    def IndexColumns(self, m, ids=None):
        """ColDo(m, func(j, v) { index.Index(v, ids[j]) }): every column of the dim x n matrix m (e.g. the K x D matrix
        Transform returns), ids default to consecutive numbers continuing from the last default id handed out (never reused after Remove)"""
        m = np.asarray(m, dtype=np.float64)
        if m.ndim != 2:
            raise ValueError("nlp: matrix must be 2-dimensional")
        if not m.flags.c_contiguous:
            m = np.ascontiguousarray(m)
        n = m.shape[1]
        if ids is not None and len(ids) != n:
            raise ValueError("nlp: one id per column")
        first_id = self._next_default_id
        handles = self._new_handles(n)
        _lib.check_index(_lib.lib().lda_b200_index_add(self._h, m.shape[0], n, _ptr(m), max(n, 1), _ptr(handles)), self._h)
        if n:
            self._dim = m.shape[0]
        if ids is None:
            self._register_auto(handles, first_id)
            self._next_default_id = first_id + n
        else:
            self._register(handles, ids)

    def IndexTransform(self, lda, m, ids=None, theta0=None):
        """ColDo(lda.Transform(m), func(j, v) { index.Index(v, ids[j]) }) without the vectors leaving the GPU: the
        Transform pipeline (lda.go:446-453) writes its K x D result straight into this index's storage.  ids default to
        the column numbers of m, offset by the default ids handed out before (0 on a fresh index)."""
        from .lda import as_matrix
        desc, _rows, cols, _keep = as_matrix(m, lda.DeviceConvert)
        h = lda._handle()
        if theta0 is None and not lda.DeviceInit:
            theta0 = lda._draws(cols, cols * lda.K)
        theta0 = None if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float64)
        if ids is not None and len(ids) != cols:
            raise ValueError("nlp: one id per column")
        handles = self._new_handles(cols)
        rc = _lib.lib().lda_b200_transform_into_index(h, desc, _ptr(theta0), lda.Rnd.src._s, lda.rhoThetaT, self._h, _ptr(handles))
        _lib.check(rc, h)
        if cols:
            self._dim = lda.K
        if ids is None:
            self._register_auto(handles, self._next_default_id)
            self._next_default_id += cols
        else:
            self._register(handles, list(ids))

    def SearchColumns(self, q, k):
        """Search for every column of the dim x nq matrix q in one pass over the index -> list of Match lists"""
        ids, dist, found = self.SearchRaw(q, k)
        out = []
        for i in range(ids.shape[0]):
            out.append([Match(float(dist[i, r]), self._id_of(int(ids[i, r]))) for r in range(int(found[i]))])
        return out

    def SearchRaw(self, q, k):
        """as SearchColumns, returning (handles int64[nq,k], distances float64[nq,k], found int32[nq]) arrays; handles are
        the positions-at-insertion numbers unless ids were given"""
        q = np.asarray(q, dtype=np.float64)
        if q.ndim != 2:
            raise ValueError("nlp: queries must be the columns of a 2-dimensional matrix")
        if not q.flags.c_contiguous:
            q = np.ascontiguousarray(q)
        nq = q.shape[1]
        k = int(k)
        ids = np.zeros((nq, max(k, 1)), dtype=np.int64)
        dist = np.zeros((nq, max(k, 1)), dtype=np.float64)
        found = np.zeros(nq, dtype=np.int32)
        if self._dim is not None and q.shape[0] != self._dim:
            raise ValueError("nlp: query has %d elements, the index holds vectors of %d" % (q.shape[0], self._dim))
        _lib.check_index(_lib.lib().lda_b200_index_search(self._h, nq, _ptr(q), max(nq, 1), k, _ptr(ids), _ptr(dist),
                                                           _ptr(found)), self._h)
        return ids, dist, found

    def LastSearchTime(self):
        """(scan ms, select ms, kernel launches, passes over the index) of the last search, CUDA events on the index's
        stream"""
        a, b, n, p = C.c_float(0), C.c_float(0), C.c_int32(0), C.c_int32(0)
        _lib.lib().lda_b200_index_last_search_time(self._h, C.byref(a), C.byref(b), C.byref(n), C.byref(p))
        return a.value, b.value, n.value, p.value


def merge_matches(per_shard, k):
    """The k nearest of the union of several shards' Search results (each a list of Match, nearest first): what a
    LinearScanIndex over all the vectors would have returned, up to the order among exactly equal distances.  NaN
    distances rank last, as in the index itself."""
    import math
    pool = [m for matches in per_shard for m in matches]
    pool.sort(key=lambda m: (math.isnan(m.Distance), m.Distance if not math.isnan(m.Distance) else 0.0))
    return pool[:k]


def SearchSharded(index, q, k):
    """Search over a sharded index: with one process per GPU (torch.distributed) every rank holds the vectors of its own
    documents -- e.g. after IndexTransform under a communicator -- searches them, and the ranks' top-k lists are gathered
    and merged.  The only exchange is k matches per query and rank; the vectors never move.  q: dim x nq matrix (the
    same on every rank); returns a list of Match lists, identical on every rank."""
    import torch.distributed as dist
    local = index.SearchColumns(q, k)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, [[tuple(m) for m in row] for row in local])
    return [merge_matches([[Match(*t) for t in shard[i]] for shard in gathered], k) for i in range(len(local))]


def NewLinearScanIndex(compareFN, device=0):
    """index.go:74-76"""
    return LinearScanIndex(compareFN, device=device)

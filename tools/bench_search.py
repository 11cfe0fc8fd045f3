"""Measurement of the nearest-neighbour scan over document vectors (not the driver's bench).

  python tools/bench_search.py [--docs 1000000] [--dim 256] [--k 10] [--queries 1,16,64] [--reps 5] [--skip-cpu]

Index = the K x D doc-topic matrix of BASELINE configs[2]'s shape (K = 256, D = 1M; synthetic Dirichlet columns), cosine
distance.  For every query-batch size prints one JSON line with
  device: scan ms / select ms (CUDA events on the index's stream, lda_b200_index_last_search_time), queries/s,
          roofline = algorithmic bytes of the scan (D x dim x 8 per pass over the index; a pass serves up to 64 queries
          with the FP64 tensor-core kernel, 16 otherwise) / scan time against the measured HBM copy bandwidth
          (MEASURED_PEAKS.json); fp64 = multiply-adds/s of the scan against the DMMA rate measured by
          tools/micro/fp64_peak.cu (63.6 per clock and SM);
  e2e:    wall clock of LinearScanIndex.SearchRaw on HOST query buffers, result ids / distances back on the host;
  cpu_baseline: the oracle's restatement of LinearScanIndex.Search (index.go:85-115, one thread -- the reference's
          Search is single-threaded) on a bounded sample of the same index.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import nlp_b200  # noqa: E402
from nlp_b200.measures import pairwise  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--docs", type=int, default=1_000_000)
    p.add_argument("--dim", type=int, default=256)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--queries", default="1,16,64")
    p.add_argument("--reps", type=int, default=5)
    p.add_argument("--cpu-docs", type=int, default=200_000)
    p.add_argument("--skip-cpu", action="store_true")
    a = p.parse_args()

    rng = np.random.default_rng(12345)
    # Dirichlet(0.1) columns from normalised gamma draws, generated in blocks to bound host memory traffic
    theta = np.empty((a.dim, a.docs), dtype=np.float64)
    for lo in range(0, a.docs, 100_000):
        hi = min(a.docs, lo + 100_000)
        g = rng.standard_gamma(0.1, size=(a.dim, hi - lo)) + 1e-12
        theta[:, lo:hi] = g / g.sum(0, keepdims=True)
    index = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    t0 = time.perf_counter()
    index.IndexColumns(theta)
    t_index = time.perf_counter() - t0
    pk, pk_src = peak()

    cpu = None
    if not a.skip_cpu:
        from oracle import index_oracle as I
        n_cpu = min(a.docs, a.cpu_docs)
        sub = np.ascontiguousarray(theta[:, :n_cpu])
        ids = np.arange(n_cpu, dtype=np.int64)
        ids_out, dist_out = np.zeros(a.k, dtype=np.int64), np.zeros(a.k)
        q = np.ascontiguousarray(theta[:, 7])
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 10.0:
            I.lib().oracle_linear_scan_search(1, a.dim, n_cpu, sub.ctypes.data, n_cpu, ids.ctypes.data, q.ctypes.data, 1, a.k,
                                              ids_out.ctypes.data, dist_out.ctypes.data)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": n_cpu / dt, "unit": "vector comparisons/s", "cores": 1, "kind": "port",
               "sample": "first %d of the %d indexed vectors, k=%d, %d searches of one query, C restatement of "
                         "LinearScanIndex.Search (index.go:85-115; single-threaded in the reference)" % (n_cpu, a.docs, a.k, reps)}

    for nq in [int(x) for x in a.queries.split(",")]:
        q = np.ascontiguousarray(theta[:, rng.integers(0, a.docs, nq)] * rng.uniform(0.5, 1.5, size=(a.dim, nq)))
        index.SearchRaw(q, a.k)                      # warm-up: pool growth
        index.SearchRaw(q, a.k)
        scan, sel, wall = [], [], []
        launches = passes = 0
        for _ in range(a.reps):
            t0 = time.perf_counter()
            ids, dist, found = index.SearchRaw(q, a.k)
            wall.append(time.perf_counter() - t0)
            s, t, launches, passes = index.LastSearchTime()
            scan.append(s)
            sel.append(t)
        alg = passes * a.docs * a.dim * 8.0
        scan_ms, sel_ms, wall_s = float(np.median(scan)), float(np.median(sel)), float(np.median(wall))
        line = {
            "metric": "linear_scan_vector_comparisons_per_sec", "unit": "vector comparisons/s",
            "value": nq * a.docs / ((scan_ms + sel_ms) * 1e-3),
            "config": {"workload": "LinearScanIndex(CosineDistance) over the K x D doc-topic matrix, K=%d, D=%d, k=%d" % (
                a.dim, a.docs, a.k), "queries_per_search": nq},
            "scan_ms": scan_ms, "select_ms": sel_ms, "queries_per_s_device": nq / ((scan_ms + sel_ms) * 1e-3),
            "e2e": {"value": nq * a.docs / wall_s, "unit": "vector comparisons/s", "seconds": wall_s,
                    "h2d_bytes_per_step": int(q.nbytes), "d2h_bytes_per_step": int(ids.nbytes + dist.nbytes + found.nbytes)},
            "gpu_launches": launches, "dtype": "f64",
            "roofline": {"bound": "hbm", "kernel": "index_scan_kernel", "achieved": alg / (scan_ms * 1e-3) / 1e9, "peak": pk,
                         "unit": "GB/s", "frac": alg / (scan_ms * 1e-3) / 1e9 / pk, "traffic": None, "peak_source": pk_src,
                         "algorithmic_bytes_per_search": alg, "passes_over_index": passes},
            "fp64": {"achieved_Tfma_per_s": nq * a.docs * a.dim / (scan_ms * 1e-3) / 1e12, "peak_Tfma_per_s": 63.6 * 148 * 1.965e9 / 1e12,
                     "frac": nq * a.docs * a.dim / (scan_ms * 1e-3) / (63.6 * 148 * 1.965e9)},
            "cpu_baseline": cpu, "index_build_s": t_index,
            "self_is_nearest": bool((dist[:, 0] >= 0).all()),
        }
        print(json.dumps(line), flush=True)
    index.Close()


if __name__ == "__main__":
    main()

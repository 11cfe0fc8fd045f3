"""Secondary measurements for DESIGN.md (not the driver's bench): the other BASELINE.json configurations.
  python tools/bench_extra.py c2|c4|c5|b100 [--docs N]
c2: fit K=100, 18,846 x 50k      c4: fit K=1024, 200k x 100k      c5: Transform K=256 (scaled: --docs held-out docs)
b100: C2 at the reference's default BatchSize=100 with Processes=16
Prints one JSON line: device ms per iteration (fit) or per call (transform), visits, token-topic updates/s."""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import corpus_synth  # noqa: E402
import nlp_b200  # noqa: E402
from nlp_b200 import _lib, rand  # noqa: E402
from nlp_b200.lda import CSC  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("case")
p.add_argument("--docs", type=int, default=0)
p.add_argument("--steps", type=int, default=5)
p.add_argument("--batch-size", type=int, default=0)
p.add_argument("--processes", type=int, default=0)
a = p.parse_args()
dev = torch.device("cuda", 0)


def fit_case(D, W, K, md, b, procs, steps):
    indptr, ind, data = corpus_synth.generate(D, W, block=max(b, 4096), mean_distinct=md, device=dev)
    lda = nlp_b200.NewLatentDirichletAllocation(K)
    lda.Rnd = rand.New(rand.NewSource(0))
    lda.BatchSize, lda.Processes, lda.Iterations = b, procs, 10 ** 6
    lda.FitBegin(CSC(W, D, indptr, ind, data))
    lda.FitIterate(2)
    _lib.lib().lda_b200_set_profiling(lda._h, 1)
    n, _, ms = lda.FitIterate(steps)
    mb_ms, mb_n, bl_ms, bl_n = C.c_float(), C.c_int32(), C.c_float(), C.c_int32()
    _lib.lib().lda_b200_last_kernel_time(lda._h, C.byref(mb_ms), C.byref(mb_n), C.byref(bl_ms), C.byref(bl_n))
    lda.PerplexityEvaluationFrequency, lda.PerplexityTolerance = 1, 0.0
    lda.FitIterate(1)
    nnz = int(indptr[-1])
    visits = 2 * nnz
    alg = nnz * ((4 * K + 8) + (8 * K + 8)) + D * (8 * K + 8)
    out = dict(case=a.case, K=K, D=D, W=W, nnz=nnz, batch_size=b, processes=procs, ms_per_iteration=ms / n,
               updates_per_s=visits * K * n / (ms * 1e-3), minibatch_kernel_ms=mb_ms.value / max(mb_n.value, 1),
               minibatch_kernel_GBps=alg * n / (mb_ms.value * 1e-3) / 1e9, blend_ms=bl_ms.value / max(bl_n.value, 1),
               kernel_share=mb_ms.value / ms, blend_share=bl_ms.value / ms, perplexity=lda.PerplexityTrace[-1])
    lda.FitEnd(False)
    return out


if a.case in ("c2", "b100"):
    D, W, K, md = corpus_synth.shape("C2")
    b, procs = (100, 16) if a.case == "b100" else (a.batch_size or 2048, a.processes or 4)
    print(json.dumps(fit_case(a.docs or D, W, K, md, b, procs, a.steps)))
elif a.case == "c4":
    D, W, K, md = corpus_synth.shape("C4")
    print(json.dumps(fit_case(a.docs or D, W, K, md, a.batch_size or 8192, a.processes or 4, a.steps)))
elif a.case == "c5":
    D, W, K, md = corpus_synth.shape("C3")
    Dt = a.docs or 262144
    tr = corpus_synth.generate(131072, W, block=8192, mean_distinct=md, device=dev)
    lda = nlp_b200.NewLatentDirichletAllocation(K)
    lda.Rnd = rand.New(rand.NewSource(0))
    lda.BatchSize, lda.Processes, lda.Iterations, lda.PerplexityEvaluationFrequency = 8192, 4, 20, 0
    lda.Fit(CSC(W, 131072, *tr))                      # "a fixed trained nPhi"
    held = corpus_synth.generate(Dt, W, seed=999, block=8192, mean_distinct=md, device=dev)
    m = CSC(W, Dt, *held)
    lda.Transform(CSC(W, 8192, held[0][:8193], held[1][:held[0][8192]], held[2][:held[0][8192]]))  # warm-up
    torch.cuda.synchronize()
    times = []
    for _ in range(3):       # the first call also grows the device memory pool to this corpus' size
        t0 = time.perf_counter()
        theta, passes = lda.Transform(m, return_passes=True)
        times.append(time.perf_counter() - t0)
    dt = min(times)
    lens = np.diff(held[0])
    visits = int((lens * np.maximum(passes, 0)).sum())
    print(json.dumps(dict(case="c5", K=K, docs=Dt, nnz=int(held[0][-1]), seconds_e2e=dt, seconds_each_call=times, mean_passes=float(passes.mean()),
                          visits=visits, updates_per_s_e2e=visits * K / dt, docs_per_s=Dt / dt,
                          ok=bool(np.abs(theta.sum(0) - 1).max() < 1e-8))))

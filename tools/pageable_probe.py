"""Diagnostic: FitTransform wall time with pageable (NumPy) vs page-locked host buffers.  python tools/pageable_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import corpus_synth  # noqa: E402
import nlp_b200  # noqa: E402
from nlp_b200 import rand  # noqa: E402
from nlp_b200.lda import CSC  # noqa: E402

D, W, K = 262144, 100000, 256
indptr, ind, data = corpus_synth.generate(D, W, block=8192, mean_distinct=200, device=torch.device("cuda", 0))
pinned = [torch.from_numpy(a).pin_memory() for a in (indptr, ind, data)]
out_pin = torch.empty((K, D), dtype=torch.float64, pin_memory=True).numpy()
cases = {"pageable": (CSC(W, D, indptr, ind, data), np.zeros((K, D))),
         "page-locked": (CSC(W, D, *[t.numpy() for t in pinned]), out_pin)}
lda = nlp_b200.NewLatentDirichletAllocation(K)
lda.BatchSize, lda.Processes, lda.Iterations, lda.PerplexityEvaluationFrequency = 8192, 4, 2, 0
res = {}
for rep in range(2):
    for name, (m, out) in cases.items():
        lda.Rnd = rand.New(rand.NewSource(0))
        lda.wordsInCorpus, lda.rhoThetaT = 0.0, 1.0
        t0 = time.perf_counter()
        theta = lda.FitTransform(m, out=out)
        res[name] = (time.perf_counter() - t0, theta.copy() if rep == 1 else None)
print({k: round(v[0], 4) for k, v in res.items()}, "max |diff| between the two runs (fp32 reduction order):", float(np.abs(res["pageable"][1] - res["page-locked"][1]).max()),
      "bytes in", indptr.nbytes + ind.nbytes + data.nbytes, "bytes out", K * D * 8)

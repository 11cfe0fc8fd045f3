"""Diagnostic: training perplexity of the C3 fit over iterations for Processes = 1 and 4 (same BatchSize), to document
that several minibatches in flight per GPU do not change what the fit converges to.  python tools/converge.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import corpus_synth  # noqa: E402
import nlp_b200  # noqa: E402
from nlp_b200 import rand  # noqa: E402
from nlp_b200.lda import CSC  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
W, K = 100_000, 256
indptr, ind, data = corpus_synth.generate(D, W, block=8192, mean_distinct=200, device=torch.device("cuda", 0))
m = CSC(W, D, indptr, ind, data)
out = {}
for procs in (1, 4):
    lda = nlp_b200.NewLatentDirichletAllocation(K)
    lda.Rnd = rand.New(rand.NewSource(0))
    lda.BatchSize, lda.Processes, lda.Iterations = 8192, procs, 60
    lda.PerplexityEvaluationFrequency, lda.PerplexityTolerance = 10, 0.0
    t0 = time.perf_counter()
    lda.Fit(m)
    out["processes_%d" % procs] = {"seconds": round(time.perf_counter() - t0, 2),
                                   "perplexity_every_10_iterations": [round(p, 1) for p in lda.PerplexityTrace]}
    lda.Close()
print(json.dumps(out))

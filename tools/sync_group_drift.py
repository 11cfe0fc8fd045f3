"""Does computing more minibatches from one nPhi snapshot (the reference's `Processes`, lda.go:487-528) slow convergence
per iteration in the REFERENCE algorithm itself?  CPU oracle only (float64 restatement of lda.go, SyncGroup = G): the
same corpus fitted for 26 iterations at G = 1, 4, 8, 16, 32 -- the question VERDICT r01 raised about the round-1 scaling
runs, whose schedule grew with the number of GPUs (G = 4N: perplexity after 26 iterations 51k / 55k / 65k / 69k at
N = 1 / 2 / 4 / 8).  python tools/sync_group_drift.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import corpus_synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

D, W, K, b, iters = 16384, 5000, 32, 128, 26
indptr, ind, data = corpus_synth.generate(D, W, block=b, mean_distinct=60, n_topics=16)
out = {"corpus": "%d docs x %d words, %d nnz, K=%d, BatchSize=%d, %d iterations, seed 0" % (D, W, int(indptr[-1]), K, b, iters),
       "perplexity_after_iterations_6_13_26": {}}
for G in (1, 4, 8, 16, 32):
    o = O.LatentDirichletAllocation(K, seed=0)
    o.BatchSize, o.Iterations, o.SyncGroup, o.HoistPow = b, iters, G, 1
    o.PerplexityEvaluationFrequency, o.PerplexityTolerance = 1, 0.0
    o.Fit((W, D, indptr, ind, data))
    t = o.perplexity_trace
    out["perplexity_after_iterations_6_13_26"]["SyncGroup_%d (%d M-steps per iteration)" % (G, -(-(D // b) // G))] = [
        round(float(t[5]), 1), round(float(t[12]), 1), round(float(t[25]), 1)]
print(json.dumps(out, indent=1))

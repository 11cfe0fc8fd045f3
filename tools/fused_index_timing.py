"""Timing of ColDo(lda.Transform(m), index.Index): through the host (Transform, then IndexColumns) against the fused
device-resident hand-over (LinearScanIndex.IndexTransform -> lda_b200_transform_into_index) on 524,288 held-out documents,
K = 256.  Needs a B200."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import corpus_synth
import nlp_b200
from nlp_b200 import rand
from nlp_b200.lda import CSC
from nlp_b200.measures import pairwise
dev = torch.device("cuda", 0)
_, W, K, md = corpus_synth.shape("C3")
lda = nlp_b200.NewLatentDirichletAllocation(K); lda.Rnd = rand.New(rand.NewSource(0))
lda.BatchSize, lda.Processes, lda.Iterations, lda.PerplexityEvaluationFrequency = 8192, 4, 10, 0
lda.Fit(CSC(W, 65536, *corpus_synth.generate(65536, W, block=8192, mean_distinct=md, device=dev)))
D = 524288
held = CSC(W, D, *corpus_synth.generate(D, W, seed=999, block=8192, mean_distinct=md, device=dev))
for rep in range(2):
    ix = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    t0 = time.perf_counter(); theta = lda.Transform(held); t1 = time.perf_counter(); ix.IndexColumns(theta); t2 = time.perf_counter()
    ix.Close()
    ix2 = nlp_b200.NewLinearScanIndex(pairwise.CosineDistance)
    t3 = time.perf_counter(); ix2.IndexTransform(lda, held); t4 = time.perf_counter()
    print("rep", rep, "Transform %.3f s + IndexColumns %.3f s = %.3f s ; fused IndexTransform %.3f s" % (t1 - t0, t2 - t1, t2 - t0, t4 - t3), flush=True)
    ix2.Close()

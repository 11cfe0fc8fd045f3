"""Where a step's time goes at the reference's default operating point (BatchSize 100, lda.go:164): per step, the minibatch
kernel, the M-step kernels and what is left (launch gaps, host), for Processes = 1 and = host cores, on a C3-shaped corpus.

    python tools/defaults_probe.py [--docs 131072] [--batch 100]
    python tools/defaults_probe.py --sweep     # launch ranges of 100 ... 16,384 documents, paired kernels off / on"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import corpus_synth  # noqa: E402
import nlp_b200  # noqa: E402
from nlp_b200 import _lib, rand  # noqa: E402
from nlp_b200.lda import CSC  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--docs", type=int, default=131072)
p.add_argument("--batch", type=int, default=100)
p.add_argument("--k", type=int, default=256)
p.add_argument("--processes", default="1,16")
p.add_argument("--sweep", action="store_true")
a = p.parse_args()
dev = torch.device("cuda", 0)
W, K = 100_000, a.k
indptr, ind, data = corpus_synth.generate(a.docs, W, block=2048, mean_distinct=200, device=dev)
lens = np.diff(indptr)
L = _lib.lib()
out = {"docs": a.docs, "K": K, "batch_size": a.batch, "doc_len_mean": float(lens.mean()), "doc_len_max": int(lens.max()),
       "doc_len_p99": float(np.percentile(lens, 99)), "runs": {}}


def run(batch, procs, pair_max=None):
    if pair_max is not None:   # otherwise whatever the caller's environment says (default: the library's own threshold)
        os.environ["LDA_B200_PAIR_MAX_DOCS"] = str(pair_max)   # read when the handle is created
    lda = nlp_b200.NewLatentDirichletAllocation(K, device=0)
    lda.Rnd = rand.New(rand.NewSource(0))
    lda.BatchSize, lda.Processes, lda.Iterations = batch, procs, 10 ** 6
    lda.PerplexityEvaluationFrequency = 0
    lda.FitBegin(CSC(W, a.docs, indptr, ind, data))
    lda.FitIterate(1)
    L.lda_b200_set_profiling(lda._h, 1)
    _, _, ms = lda.FitIterate(1)
    mb_ms, mb_n, bl_ms, bl_n = C.c_float(), C.c_int32(), C.c_float(), C.c_int32()
    L.lda_b200_last_kernel_time(lda._h, C.byref(mb_ms), C.byref(mb_n), C.byref(bl_ms), C.byref(bl_n))
    L.lda_b200_set_profiling(lda._h, 0)
    _, _, ms_noprof = lda.FitIterate(1)
    lda.PerplexityEvaluationFrequency = 1
    lda.FitIterate(1)
    perp = lda.PerplexityTrace[-1]
    lda.FitEnd(want_theta=False)
    lda.Close()
    steps = max(mb_n.value, 1)
    return {"docs_per_launch": batch * procs, "steps": mb_n.value, "ms_per_iteration": ms,
            "ms_per_iteration_without_event_pairs": ms_noprof,
            "ms_per_step": ms_noprof / steps, "minibatch_kernel_ms_per_step": mb_ms.value / steps,
            "m_step_kernels_ms_per_step": bl_ms.value / max(bl_n.value, 1),
            "rest_ms_per_step": (ms_noprof - mb_ms.value - bl_ms.value) / steps, "perplexity_after_4_iterations": perp}


if a.sweep:
    for batch, procs in ((100, 1), (100, 16), (256, 16), (512, 16), (1024, 16), (2048, 8), (2048, 16)):
        for name, pm in (("unpaired", 0), ("paired", 1 << 30)):
            out["runs"]["%dx%d_%s" % (batch, procs, name)] = run(batch, procs, pm)
else:
    for procs in [int(x) for x in a.processes.split(",")]:
        out["runs"]["processes_%d" % procs] = run(a.batch, procs)
print(json.dumps(out))

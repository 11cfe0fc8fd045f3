"""Time-to-perplexity of the configs[2] fit (1M documents x 100k words, K = 256) as a function of the schedule -- BatchSize b
and Processes G (minibatches that share an nPhi snapshot per M-step, lda.go:487-528) -- on N GPUs.

    python tools/converge.py [--schedules 8192x1,8192x4,2048x16,2048x32,2048x64,2048x128] [--iterations 60] [--every 5]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/converge.py --schedules 2048x128

Per schedule: device seconds per iteration (resident corpus, CUDA events, max over ranks), the training perplexity every
`--every` iterations, and for each target in --targets the seconds of fit time until the first evaluation at or below it
(perplexity evaluations are not counted).  Prints one JSON line.  The schedule decides the model trajectory, the number
of GPUs does not: bG minibatches are spread over the GPUs, so N = 8 reaches a given perplexity N-times-efficiency faster
along the same curve."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import corpus_synth  # noqa: E402
import nlp_b200  # noqa: E402
from nlp_b200 import rand  # noqa: E402
from nlp_b200.lda import CSC  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--docs", type=int, default=1_000_000)
p.add_argument("--schedules", default="8192x1,8192x4,2048x16,2048x32,2048x64,2048x128")
p.add_argument("--iterations", type=int, default=60)
p.add_argument("--every", type=int, default=5)
p.add_argument("--targets", default="100000,30000,15000")
a = p.parse_args()

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
W, K = 100_000, 256
targets = [float(t) for t in a.targets.split(",")]
out = {"n_gpus": world, "docs": a.docs, "iterations": a.iterations, "schedules": {}}
for sched in a.schedules.split(","):
    b, G = (int(x) for x in sched.split("x"))
    if G % world:
        continue
    indptr, ind, data = corpus_synth.generate(a.docs, W, block=2048, mean_distinct=200, device=dev,
                                              owned=(lambda m: (m * 2048 // b) % world == rank) if b >= 2048 else None)
    lda = nlp_b200.NewLatentDirichletAllocation(K, device=local)
    lda.Rnd = rand.New(rand.NewSource(0))
    lda.BatchSize, lda.Processes, lda.Iterations = b, G, 10 ** 6
    lda.PerplexityEvaluationFrequency, lda.PerplexityTolerance = 0, 0.0
    lda.FitBegin(CSC(W, a.docs, indptr, ind, data))
    fit_s, trace, reached = 0.0, [], {}
    for it in range(0, a.iterations, a.every):
        lda.PerplexityEvaluationFrequency = 0
        _, _, ms = lda.FitIterate(a.every - 1)
        lda.PerplexityEvaluationFrequency = 1       # the last iteration of the block is followed by an evaluation
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        _, _, ms1 = lda.FitIterate(1)
        # the evaluation itself is inside ms1: replace that iteration's time by the block's mean iteration time
        t += ms / max(a.every - 1, 1) if a.every > 1 else ms1
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fit_s += float(t.item()) * 1e-3
        perp = lda.PerplexityTrace[-1]
        trace.append(round(perp, 1))
        for tg in targets:
            if perp <= tg and tg not in reached:
                reached[tg] = round(fit_s, 3)
    lda.FitEnd(want_theta=False)
    lda.Close()
    out["schedules"][sched] = {"batch_size": b, "processes": G, "m_steps_per_iteration": -(-(-(-a.docs // b)) // G),
                               "ms_per_iteration": round(1e3 * fit_s / a.iterations, 2),
                               "perplexity_every_%d_iterations" % a.every: trace,
                               "seconds_to_perplexity": {str(int(k)): v for k, v in reached.items()}}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()

"""BASELINE configs[4]: Transform-only inference of held-out documents at K = 256 with a fixed trained nPhi, sharded over
the GPUs of one box (not the driver's bench).

  python tools/bench_transform.py [--docs 524288]                                   # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      tools/bench_transform.py --docs 4000000                                       # configs[4] as stated

Every rank owns the minibatches m with m % world == rank of the held-out corpus (no data-path collective: documents are
independent, lda.go:429-435), materialises only those, and calls LatentDirichletAllocation.Transform on HOST buffers
(pageable NumPy arrays, like Go slices); the K x D result lands in the rank's columns.  Prints one JSON line from rank 0:
end-to-end seconds (max over ranks, best of --reps calls), documents/s and token-topic updates/s over all GPUs.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import corpus_synth  # noqa: E402
import nlp_b200  # noqa: E402
from nlp_b200 import rand  # noqa: E402
from nlp_b200.lda import CSC  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--docs", type=int, default=524288)
    p.add_argument("--train-docs", type=int, default=131072)
    p.add_argument("--reps", type=int, default=3)
    a = p.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _, W, K, md = corpus_synth.shape("C3")
    b = 8192
    owned = (lambda m: m % world == rank)

    lda = nlp_b200.NewLatentDirichletAllocation(K, device=local)
    lda.Rnd = rand.New(rand.NewSource(0))
    lda.BatchSize, lda.Processes, lda.Iterations, lda.PerplexityEvaluationFrequency = b, 4 * world, 20, 0
    lda.Fit(CSC(W, a.train_docs, *corpus_synth.generate(a.train_docs, W, block=b, mean_distinct=md, device=dev, owned=owned)))
    held = corpus_synth.generate(a.docs, W, seed=999, block=b, mean_distinct=md, device=dev, owned=owned)
    m = CSC(W, a.docs, *held)
    nnz_local = int(held[0][-1])

    def allreduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    times, passes = [], None
    for _ in range(a.reps):       # the first call also grows the device memory pool to this shard's size
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        theta, passes = lda.Transform(m, return_passes=True)
        torch.cuda.synchronize()
        times.append(allreduce(time.perf_counter() - t0, dist.ReduceOp.MAX))
    lens = np.diff(held[0])
    visits = allreduce(float((lens * np.maximum(passes, 0)).sum()), dist.ReduceOp.SUM)
    nnz = allreduce(float(nnz_local), dist.ReduceOp.SUM)
    mine = np.flatnonzero(lens > 0)[:1000]
    ok = bool(np.abs(theta[:, mine].sum(0) - 1).max() < 1e-8)
    dt = min(times)
    if rank == 0:
        print(json.dumps({
            "metric": "transform_token_topic_updates_per_sec", "value": visits * K / dt, "unit": "token-topic updates/s",
            "n_gpus": world, "config": {"workload": "configs[4]: Transform of %d held-out docs, K=%d, W=%d, %d nnz, fixed nPhi from a "
                                        "%d-doc fit; sharded by minibatch of %d over %d GPU(s)" % (a.docs, K, W, int(nnz), a.train_docs, b, world)},
            "e2e": {"seconds": dt, "seconds_each_call": times, "docs_per_s": a.docs / dt, "h2d_bytes": int(nnz) * 16 + 8 * (a.docs + 1),
                    "d2h_bytes": a.docs * K * 8, "host_buffers": "pageable"},
            "token_visits": int(visits), "columns_sum_to_one": ok}), flush=True)
    lda.Close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

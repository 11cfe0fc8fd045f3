"""world_size-2 `gloo` test (CPU) of the host-side logic of the N > 1 path: document sharding and the algebra that
turns the ranks' blends into one allreduce(sum).  The CUDA kernels are not involved; the all-reduce
that NCCL performs on the GPUs is done here by gloo on the same coefficients the engine uses."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, D, b, W, K, steps, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nlp_b200 import _lib
        from nlp_b200.lda import _world
        from oracle import oracle as O
        L = _lib.lib()
        assert _world() == (rank, world)

        # 1. sharding: every document is owned by exactly one rank
        n = L.lda_b200_plan_shard(D, b, rank, world, None, None, 0)
        beg, end = np.zeros(max(n, 1), dtype=np.int64), np.zeros(max(n, 1), dtype=np.int64)
        L.lda_b200_plan_shard(D, b, rank, world, beg.ctypes.data, end.ctypes.data, n)
        owned = torch.zeros(D, dtype=torch.int64)
        for i in range(n):
            owned[beg[i]:end[i]] += 1
        dist.all_reduce(owned)
        assert bool((owned == 1).all())

        # 2. steps of (scale by c_g, allreduce, blend) == the sequential blends of lda.go:299-317
        sched = _lib.Schedule(10, 1000, 0.9)
        rng = np.random.default_rng(1234)           # same stream on every rank
        nphi = rng.random((W, K))
        want = nphi.copy()
        M = (D + b - 1) // b
        rho_phi_t = 1.0
        for s in range(steps):
            n_s = min(world, M - s * world)
            hats = rng.random((world, W, K)) * 50.0   # hats[g] = what rank g's minibatch kernel produced
            for g in range(n_s):
                rho = O.schedule_calc(10, 1000, 0.9, rho_phi_t + g)
                want = (1 - rho) * want + rho * hats[g]
            A, cg = C.c_double(), C.c_double()
            assert L.lda_b200_plan_step(sched, rho_phi_t, n_s, rank, C.byref(A), C.byref(cg)) == 0
            mine = torch.from_numpy(cg.value * hats[rank] if rank < n_s else np.zeros((W, K)))
            dist.all_reduce(mine)                     # ncclAllReduce(sum) on the GPUs
            nphi = A.value * nphi + mine.numpy()
            rho_phi_t += n_s
        np.testing.assert_allclose(nphi, want, rtol=1e-12)
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("D,b", [(1000, 128), (300, 100)])
def test_two_rank_step_algebra(D, b):
    world = 2
    M = (D + b - 1) // b
    steps = (M + world - 1) // world   # the last step may hold a single minibatch
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), D, b, 40, 6, steps, ret), nprocs=world, join=True)
    assert dict(ret) == {0: 1, 1: 1}


def _search_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import nlp_b200
        from nlp_b200 import Match

        class Shard:  # stands in for a LinearScanIndex holding this rank's documents (no GPU in the CPU suite)
            def SearchColumns(self, q, k):
                rng = np.random.default_rng(100 + rank)
                vecs, ids = rng.random((q.shape[0], 50)), [rank * 1000 + j for j in range(50)]
                out = []
                for i in range(q.shape[1]):
                    d = np.sqrt(((vecs - q[:, [i]]) ** 2).sum(0))
                    order = np.argsort(d, kind="stable")[:k]
                    out.append([Match(float(d[j]), ids[j]) for j in order])
                return out

        q = np.random.default_rng(7).random((6, 3))
        ret[rank] = nlp_b200.SearchSharded(Shard(), q, 5)
    finally:
        dist.destroy_process_group()


def test_sharded_search_merges_the_ranks_top_k():
    """SearchSharded: every rank searches its own shard, the top-k lists are gathered (k matches per query and rank) and
    merged -- the result equals a search over the union of the shards and is identical on every rank"""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_search_worker, args=(world, port, ret), nprocs=world, join=True)
        got = [ret[r] for r in range(world)]
    assert got[0] == got[1]
    q = np.random.default_rng(7).random((6, 3))
    vecs = np.concatenate([np.random.default_rng(100 + r).random((6, 50)) for r in range(world)], axis=1)
    ids = [r * 1000 + j for r in range(world) for j in range(50)]
    for i in range(3):
        d = np.sqrt(((vecs - q[:, [i]]) ** 2).sum(0))
        want = [ids[j] for j in np.argsort(d, kind="stable")[:5]]
        assert [m.ID for m in got[0][i]] == want
        assert all(a.Distance <= b.Distance for a, b in zip(got[0][i], got[0][i][1:]))

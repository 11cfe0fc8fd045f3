#!/usr/bin/env python
"""bench.py -- SCVB0 token-topic updates/sec of the LDA fit (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (C restatement, see below)

A *step* is one SCVB0 iteration of lda.go:501-528 over the whole synthetic corpus: every minibatch's fused
burn-in + sufficient-statistics kernel, the allreduce (N > 1) and the rho-blend M-step.  Default workload is
BASELINE.json configs[2] -- the configuration the metric is quoted on -- K = 256 topics, 1M documents x 100k words,
~200M non-zeros, Zipfian, BatchSize 8192 documents per GPU-minibatch; with N GPUs the same corpus is sharded by
minibatch (strong scaling), one process per GPU, nPhiHat reduced across the GPUs every step (one reduce + M-step kernel
over NVLink peer memory; ncclAllReduce as fallback).

value   = non-zero visits x K / time of K timed steps, inputs already resident in HBM, CUDA events on the launching
          stream, max over ranks.   unit: token-topic updates/s
e2e     = the same metric through the public API call a user makes -- LatentDirichletAllocation.FitTransform(m) on
          HOST buffers (page-locked), K iterations: includes the host->device copy of the CSC, the device-side
          ingestion / PCG init, and the device->host read of the K x D doc-topic matrix.
The reference is Go and cannot be built in this image (no Go toolchain); `--impl reference` and `cpu_baseline` time
the line-faithful C restatement under oracle/ (kind "port") with the reference's goroutine scheme on host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "scvb0_token_topic_updates_per_sec"
UNIT = "token-topic updates/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4"])
    p.add_argument("--batch-size", type=int, default=8192)
    p.add_argument("--processes-per-gpu", type=int, default=4,
                   help="minibatches in flight per GPU (the reference's Processes / number of GPUs)")
    p.add_argument("--docs", type=int, default=0, help="override the number of documents (debug)")
    p.add_argument("--cpu-docs-per-worker", type=int, default=256)
    p.add_argument("--cpu-max-threads", type=int, default=64)
    p.add_argument("--e2e-warmup", type=int, default=1, help="untimed FitTransform calls (1 iteration) before the timed one")
    p.add_argument("--deterministic", action="store_true",
                   help="bitwise reproducible fits: fixed-point accumulation of nPhiHat (lda_b200_set_deterministic)")
    p.add_argument("--skip-cpu", action="store_true")
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--skip-parity", action="store_true")
    return p.parse_args()


def workload(args):
    import corpus_synth
    D, W, K, md = corpus_synth.shape(args.workload)
    if args.docs:
        D = args.docs
    name = {"C1": "configs[0]", "C2": "configs[1]", "C3": "configs[2]", "C4": "configs[3]"}[args.workload]
    return D, W, K, md, name


def alg_bytes(nnz, docs, K, B):
    """fp32 state, int32 index + fp32 value: burn-in visit 4K+8 B, main visit 8K+8 B, document 8K+8 B"""
    return nnz * (B * (4 * K + 8) + (8 * K + 8)) + docs * (8 * K + 8)


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU during the timed region (NVML)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
            "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4, "hw_power_brake": 0x80,
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(K, docs_per_launch):
    """dram bytes per launch of the minibatch kernel from the committed `ncu --set full` capture, if any"""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            j = json.load(open(path))
            if int(j.get("K", -1)) == K and int(j.get("docs_per_launch", -1)) == docs_per_launch:
                return j.get("dram_bytes_per_launch")
        except Exception:
            pass
    return None


# ------------------------------------------------------------------------------------------------------------------
def cpu_run(args, D, W, K, md, steps, warmup, device):
    """the reference's CPU algorithm (C restatement, goroutine pool -> pthreads) on a bounded sample of the workload:
    the first cores x docs_per_worker documents, one minibatch per worker.  Returns the cpu_baseline dict."""
    import corpus_synth
    from oracle import oracle as O
    # one worker per host core, capped: every worker owns a W x K float64 nPhiHat (205 MB at this shape), lda.go:492-495
    cores = min(os.cpu_count() or 1, args.cpu_max_threads)
    per = args.cpu_docs_per_worker
    n_docs = min(D, cores * per)
    indptr, ind, data = corpus_synth.generate(n_docs, W, block=args.batch_size, mean_distinct=md, device=device)
    o = O.LatentDirichletAllocation(K, seed=0, processes=cores)
    o.BatchSize = per
    o.Iterations = steps + warmup
    o.PerplexityEvaluationFrequency = 0
    t0 = time.perf_counter()
    o.Fit((W, n_docs, indptr, ind, data))
    wall = time.perf_counter() - t0
    t = o.iteration_times
    timed = float(t[-1] - (t[warmup - 1] if warmup > 0 else 0.0))
    visits_per_step = int(indptr[-1]) * (int(o.BurnInPasses) + 1)
    value = visits_per_step * K * steps / timed
    return {
        "value": value, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": "first %d docs of the corpus (%d nnz), BatchSize=%d, Processes=%d (one minibatch per host thread), "
                  "%d timed + %d warm-up iterations, float64 C restatement of lda.go (Go toolchain absent)" % (
                      n_docs, int(indptr[-1]), per, cores, steps, warmup),
        "ms_per_step": 1e3 * timed / steps, "wall_s": wall,
    }


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    D, W, K, md, cfgname = workload(args)
    try:
        import torch
        device = "cuda" if torch.cuda.is_available() else "cpu"
    except Exception:
        device = "cpu"
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cb = cpu_run(args, D, W, K, md, steps, warmup, device)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: LDA SCVB0 K=%d, %d docs x %d vocab (bounded CPU sample, see cpu_baseline.sample)" % (
            cfgname, K, D, W), "batch_size": args.cpu_docs_per_worker},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist

    import corpus_synth
    import nlp_b200
    from nlp_b200 import _lib, rand
    from nlp_b200.lda import CSC

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D, W, K, md, cfgname = workload(args)
    b = args.batch_size
    steps, warmup = max(1, args.steps), max(0, args.warmup)

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic corpus: this rank materialises only the minibatches it owns -----------------------------------
    t_gen = time.perf_counter()
    indptr, ind, data = corpus_synth.generate(D, W, block=b, mean_distinct=md, device=dev,
                                              owned=(lambda m: m % world == rank))
    pinned = [torch.from_numpy(a).pin_memory() for a in (indptr, ind, data)]   # page-locked host buffers
    indptr, ind, data = [t.numpy() for t in pinned]
    m = CSC(W, D, indptr, ind, data)
    nnz_local = int(indptr[-1])
    nnz = int(allsum(nnz_local))
    docs_local = sum(min(b, D - mb * b) for mb in range((D + b - 1) // b) if mb % world == rank)
    t_gen = time.perf_counter() - t_gen

    def new_lda(iterations):
        lda = nlp_b200.NewLatentDirichletAllocation(K, device=local)
        lda.Rnd = rand.New(rand.NewSource(0))
        lda.BatchSize = b
        lda.Iterations = iterations
        lda.Processes = world * args.processes_per_gpu
        lda.Deterministic = args.deterministic
        return lda

    B = 1  # BurnInPasses default
    visits_per_step = nnz * (B + 1)

    # ---- value: K timed iterations on the resident corpus ---------------------------------------------------------
    lda = new_lda(10 ** 6)
    lda.FitBegin(m)
    h = lda._h
    L = _lib.lib()
    lda.FitIterate(warmup)
    L.lda_b200_set_profiling(h, 1)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.lda_b200_launch_count(h)
    t0 = time.perf_counter()
    ran, _, dev_ms = lda.FitIterate(steps)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    launches = L.lda_b200_launch_count(h) - launches0
    assert ran == steps, (ran, steps)
    import ctypes as C
    mb_ms, mb_n, bl_ms, bl_n = C.c_float(), C.c_int32(), C.c_float(), C.c_int32()
    L.lda_b200_last_kernel_time(h, C.byref(mb_ms), C.byref(mb_n), C.byref(bl_ms), C.byref(bl_n))
    ar_ms, ar_n = C.c_float(), C.c_int32()
    L.lda_b200_last_allreduce_time(h, C.byref(ar_ms), C.byref(ar_n))
    L.lda_b200_set_profiling(h, 0)
    kernel_ms_max, kernel_ms_min = allmax(mb_ms.value), -allmax(-mb_ms.value)
    ms = allmax(dev_ms)
    wall_ms = allmax(wall_ms)
    total_launches = int(allsum(launches))
    value = visits_per_step * K * steps / (ms * 1e-3)
    # one perplexity evaluation outside the timed region, for the record
    lda.PerplexityEvaluationFrequency = 1
    lda.PerplexityTolerance = 0.0
    lda.FitIterate(1)
    perplexity = lda.PerplexityTrace[-1] if lda.PerplexityTrace else None
    lda.FitEnd(want_theta=False)

    exchange = ("nPhiHat reduced + M-step in one kernel over NVLink peer memory per step" if lda.CommInfo()[2] else
                "ncclAllReduce of nPhiHat + M-step per step") if world > 1 else "no exchange"
    peak, peak_src = measured_peak()
    kbytes = alg_bytes(nnz_local, docs_local, K, B) * steps      # what this rank's minibatch kernels processed
    achieved = kbytes / (mb_ms.value * 1e-3) / 1e9 if mb_ms.value > 0 else None
    roofline = {
        "bound": "hbm", "kernel": "scvb0_docs_kernel (fused burn-in + sufficient-statistics pass, one launch per minibatch)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
        "traffic": committed_traffic(K, b * args.processes_per_gpu), "peak_source": peak_src,
        "note": "achieved counts ALGORITHMIC bytes (SURVEY 8d: a 4K-byte row of C per token visit, 4K more of reductions per "
                "main visit); at K=256 nPhi/C/nPhiHat (3 x 102 MB) are largely resident in the 126 MB L2, so most of those "
                "bytes never reach DRAM (`traffic` = measured DRAM bytes per launch) and frac can exceed 1: the kernel is "
                "bound by L2 bandwidth, see `l2`",
        # informational: the same bytes seen as L2 traffic (rows of C and the reductions are L2-resident, DRAM carries
        # only `traffic`) against the full-chip L2 cap of B300_MICROARCH.md (~6300 B/clk) at the sampled SM clock
        "l2": {"achieved_GBps": achieved, "cap_GBps": (6300.0 * clocks["sm_mhz"] * 1e6 / 1e9) if clocks.get("sm_mhz") else None,
               "frac": (achieved / (6300.0 * clocks["sm_mhz"] * 1e6 / 1e9)) if (achieved and clocks.get("sm_mhz")) else None},
        "launches": mb_n.value, "avg_launch_ms": (mb_ms.value / mb_n.value) if mb_n.value else None,
        "algorithmic_bytes_per_launch": kbytes / mb_n.value if mb_n.value else None,
        "kernel_share_of_step": mb_ms.value / dev_ms if dev_ms else None,
        "blend_kernel": {"launches": bl_n.value, "avg_launch_ms": (bl_ms.value / bl_n.value) if bl_n.value else None,
                         "achieved_GBps": (16.0 * K * W * bl_n.value / (bl_ms.value * 1e-3) / 1e9) if bl_ms.value else None},
        "allreduce": {"path": "peer-memory reduce+blend kernel (CUDA IPC over NVLink)" if lda.CommInfo()[2] else
                      ("ncclAllReduce + blend kernel" if world > 1 else "none (1 GPU)"), "calls": ar_n.value, "avg_ms_incl_wait_for_slowest_rank": (ar_ms.value / ar_n.value) if ar_n.value else None,
                      "bytes": 4 * K * W, "minibatch_kernel_ms_rank_min_max": [kernel_ms_min / max(mb_n.value, 1),
                                                                                 kernel_ms_max / max(mb_n.value, 1)]},
        "whole_step_achieved_GBps": (alg_bytes(nnz, D, K, B) + bl_n.value / steps * 16.0 * K * W * world) * steps / (ms * 1e-3) / 1e9 / world,
    }

    # ---- e2e: the public call on host buffers ---------------------------------------------------------------------------
    e2e = None
    if not args.skip_e2e:
        out_t = torch.empty((K, D), dtype=torch.float64, pin_memory=True)   # page-locked result buffer
        out = out_t.numpy()
        # the same model object a user would keep around (its communicator is already connected); a fresh RNG seed and
        # fresh counters make it the same call as on a new object
        lda2 = lda
        lda2.Iterations, lda2.PerplexityEvaluationFrequency, lda2.PerplexityTolerance = steps, 30, 1e-2
        lda2.Rnd = rand.New(rand.NewSource(0))
        for _ in range(args.e2e_warmup):          # untimed: first-touch of the page-locked result buffer, pool growth
            lda2.rhoPhiT, lda2.rhoThetaT, lda2.wordsInCorpus = 1.0, 1.0, 0.0
            lda2.Iterations = 1
            lda2.FitTransform(m, out=out)
        lda2.Iterations = steps
        lda2.Rnd = rand.New(rand.NewSource(0))
        lda2.rhoPhiT, lda2.rhoThetaT, lda2.wordsInCorpus = 1.0, 1.0, 0.0
        barrier()
        t0 = time.perf_counter()
        lda2.FitTransform(m, out=out)
        torch.cuda.synchronize()
        t_e2e = allmax(time.perf_counter() - t0)
        h2d = int(allsum(indptr.nbytes + ind.nbytes + data.nbytes))
        d2h = int(allsum(K * docs_local * 8))
        ok = bool(np.isfinite(out[:, :min(D, 64)]).all()) if rank == 0 else True
        e2e = {"value": visits_per_step * K * steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d // steps,
               "d2h_bytes_per_step": d2h // steps, "seconds": t_e2e, "iterations": steps, "finite": ok,
               "call": "LatentDirichletAllocation.FitTransform(CSC host buffers) -> K x D float64"}
        del out
    lda.Close()

    # ---- parity on a sample + CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------
    parity, cpu = None, None
    if rank == 0 and world == 1 and not args.skip_parity:
        from oracle import oracle as O
        n_s = min(D, 384)
        sub = CSC(W, n_s, indptr[:n_s + 1].copy(), ind[:indptr[n_s]].copy(), data[:indptr[n_s]].copy())
        f = dict(Iterations=2, BatchSize=128, PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0)
        a = new_lda(2)
        a.Processes = 1      # deterministic mode: same minibatch order as the reference at Processes = 1
        o = O.LatentDirichletAllocation(K, seed=0)
        for k_, v_ in f.items():
            setattr(a, k_, v_)
            setattr(o, k_, v_)
        got, want = a.FitTransform(sub), o.FitTransform(tuple(sub))
        parity = {"sample": "first %d docs, BatchSize=128, 2 iterations, seed 0, Processes=1" % n_s,
                  "theta_max_abs_delta": float(np.abs(got - want).max()),
                  "perplexity_rel_delta": float(np.abs(np.array(a.PerplexityTrace) / o.perplexity_trace - 1).max())}
        a.Close()
    if rank == 0 and world == 1 and not args.skip_cpu:
        cb = cpu_run(args, D, W, K, md, 1, 0, dev)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: LDA SCVB0 fit, K=%d, %d docs x %d vocab, %d nnz (Zipfian planted-LDA corpus, seed 12345)"
                       % (cfgname, K, D, W, nnz), "batch_size": b, "processes": world * args.processes_per_gpu,
                       "burn_in_passes": B, "deterministic": bool(args.deterministic),
                       "parallelism": "dp%d (minibatch m -> GPU m %% %d, %d minibatches per GPU per step, %s)" % (
                           world, world, args.processes_per_gpu, exchange),
                       "step": "one SCVB0 iteration over all %d minibatches (%d launches per GPU)" % (
                           (D + b - 1) // b, -(-((D + b - 1) // b) // (world * args.processes_per_gpu))),
                       "l2": "inputs larger than L2 (CSC + nTheta = %.1f GB per GPU per step)" % (
                           (nnz_local * 8 + docs_local * K * 4) / 1e9),
                       "token_visits_per_step": visits_per_step, "corpus_generation_s": round(t_gen, 1)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": total_launches, "roofline": roofline, "cpu_baseline": cpu,
            "wall_ms_per_step": wall_ms / steps, "perplexity_after_%d_iterations" % (warmup + steps + 1): perplexity,
            "parity": parity,
        }
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _emit(line):
    os.write(_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    # the contract is ONE JSON line on stdout: anything else a library prints there (e.g. NCCL's version banner) is
    # sent to stderr by pointing fd 1 at fd 2 for the duration of the run; the JSON line goes to the saved fd
    sys.stdout.flush()
    _STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)

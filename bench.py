#!/usr/bin/env python
"""bench.py -- SCVB0 token-topic updates/sec of the LDA fit (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (C restatement, see below)

A *step* is one SCVB0 iteration of lda.go:501-528 over the whole synthetic corpus: every minibatch's fused
burn-in + sufficient-statistics kernel, the exchange of nPhiHat (N > 1) and the rho-blend M-step.  Default workload is
BASELINE.json configs[2] -- the configuration the metric is quoted on -- K = 256 topics, 1M documents x 100k words,
~200M non-zeros, Zipfian; BatchSize 2048, and `Processes` = 128 minibatches computed from one nPhi snapshot per M-step
(lda.go:487-528: the reference's goroutine pool).  The schedule does not depend on N: with N GPUs the same 128
minibatches per step are spread over the GPUs (minibatch m -> GPU m mod N), so every N fits the same model -- the
perplexity in the line is the same at N = 1, 2, 4, 8 -- and the scaling is strong scaling of one fixed job.
`--processes-per-gpu P` gives the other convention (Processes = P x N, the schedule changes with N).

value   = non-zero visits x K / time of K timed steps, inputs already resident in HBM, CUDA events on the launching
          stream, max over ranks.   unit: token-topic updates/s
e2e     = the same metric through the public API call a user makes -- LatentDirichletAllocation.FitTransform(m) on
          HOST buffers (page-locked), K iterations: includes the host->device copy of the CSC, the device-side
          ingestion / PCG init, and the device->host read of the K x D doc-topic matrix.
e2e_pageable        = the same call on pageable NumPy buffers (what a Go caller's slices are): first call on a new
                      object (cold: handle, pools, staging ring, communicator) and second call (warm).
reference_defaults  = the fit at the reference's own defaults, BatchSize 100 (lda.go:164) and Processes = host cores
                      (lda.go:185), and at Processes = 1; plus the CPU port at that same operating point
                      (cpu_port_same_operating_point, SURVEY 8d: "the CPU oracle run at the same b").
roofline.memory_floor = the minibatch kernel's memory access mix replayed without arithmetic or dependencies
                      (lda_b200_probe_memory_mix), measured in the same process: what bounds the kernel is the L2's
                      atomic (red.add) throughput, not HBM.
The reference is Go and cannot be built in this image (no Go toolchain); `--impl reference` and `cpu_baseline` time
the line-faithful C restatement under oracle/ (kind "port") with the reference's goroutine scheme on host threads, at
the same BatchSize, on a bounded sample of the same corpus (one minibatch per host thread and step).
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
CFG_NAME = {"C1": "configs[0]", "C2": "configs[1]", "C3": "configs[2]", "C4": "configs[3]", "C5": "configs[4]"}
# BatchSize / minibatches per step (Processes) of the throughput setting per workload; C1 keeps the reference defaults
DEFAULTS = {"C1": (100, 1), "C2": (2048, 4), "C3": (2048, 128), "C4": (2048, 32), "C5": (2048, 1)}
B = 1  # BurnInPasses default (lda.go:168)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    p.add_argument("--batch-size", type=int, default=0, help="BatchSize (default per workload: C3 2048)")
    p.add_argument("--minibatches-per-step", type=int, default=0,
                   help="Processes: minibatches that share an nPhi snapshot, the same at every N (default per workload: C3 128)")
    p.add_argument("--processes-per-gpu", type=int, default=0,
                   help="alternative convention: Processes = this x number of GPUs (the schedule then changes with N)")
    p.add_argument("--docs", type=int, default=0, help="override the number of documents (debug)")
    p.add_argument("--cpu-max-threads", type=int, default=64)
    p.add_argument("--e2e-warmup", type=int, default=1, help="untimed FitTransform calls (1 iteration) before the timed one")
    p.add_argument("--deterministic", action="store_true",
                   help="bitwise reproducible fits: fixed-point accumulation of nPhiHat (lda_b200_set_deterministic)")
    p.add_argument("--skip-cpu", action="store_true")
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--skip-parity", action="store_true")
    p.add_argument("--skip-extras", action="store_true", help="no e2e_pageable / reference_defaults legs")
    a = p.parse_args()
    a.world = int(os.environ.get("WORLD_SIZE", "1")) if a.impl == "ours" else max(1, a.gpus)
    b, g = DEFAULTS[a.workload]
    a.batch_size = a.batch_size or b
    if a.processes_per_gpu:
        a.group = a.processes_per_gpu * a.world
    else:
        a.group = a.minibatches_per_step or g
        a.group = max(a.world, a.group // a.world * a.world)   # the same number on every GPU
    return a


def workload(args):
    import corpus_synth
    if args.workload == "C5":
        D, W, K, md = 4_000_000, 100_000, 256, 200
    else:
        D, W, K, md = corpus_synth.shape(args.workload)
    if args.docs:
        D = args.docs
    return D, W, K, md


def common_config(args, D, W, K, md):
    """the `config` of the line: identical for both arms (the reference arm runs on our arm's config)"""
    world, b, G = args.world, args.batch_size, args.group
    fit = args.workload != "C5"
    return {
        "workload": "%s: LDA SCVB0 %s, K=%d, %d docs x %d vocab, ~%d distinct words per document (Zipfian planted-LDA "
                    "corpus, seed 12345)" % (CFG_NAME[args.workload], "fit" if fit else "Transform of held-out documents",
                                             K, D, W, md),
        "batch_size": b, "processes": G if fit else 1, "burn_in_passes": B, "deterministic": bool(args.deterministic),
        "parallelism": "dp%d: minibatch m -> GPU m %% %d; %d minibatches share an nPhi snapshot per M-step (%d per GPU and "
                       "launch)%s" % (world, world, G, G // world, "; nPhiHat exchanged every step" if world > 1 else "")
                       if fit else "dp%d: minibatch m -> GPU m %% %d, no data-path collective" % (world, world),
        "step": ("one SCVB0 iteration over all %d minibatches (%d M-steps)" % ((D + b - 1) // b, -(-((D + b - 1) // b) // G)))
                if fit else "one Transform call over all documents (host buffers in, K x D float64 out)",
        "l2": "inputs larger than L2 (CSC + nTheta ~ %.1f GB per GPU per step)" % ((D * md * 8 + D * K * 4) / world / 1e9),
    }


def alg_bytes(nnz, docs, K, passes):
    """fp32 state, int32 index + fp32 value: burn-in visit 4K+8 B, main visit 8K+8 B, document 8K+8 B"""
    return nnz * (passes * (4 * K + 8) + (8 * K + 8)) + docs * (8 * K + 8)


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
            self._halt.wait(0.02)

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


def committed_traffic(K, docs_per_launch, world):
    """dram bytes per launch of the minibatch kernel from the committed `ncu --set full` capture of the same
    configuration on one GPU; null for any other configuration (it is not a live measurement)"""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if world == 1 and os.path.exists(path):
        try:
            j = json.load(open(path))
            if int(j.get("K", -1)) == K and int(j.get("docs_per_launch", -1)) == docs_per_launch:
                return j.get("dram_bytes_per_launch")
        except Exception:
            pass
    return None


# ------------------------------------------------------------------------------------------------------------------
def cpu_run(args, D, W, K, md, steps, warmup, device):
    """the reference's CPU algorithm (C restatement, goroutine pool -> pthreads) on a bounded sample of the workload at
    the workload's BatchSize: the first T minibatches, T = host threads used (at most Processes), one minibatch per
    thread and step.  Returns the cpu_baseline dict."""
    import corpus_synth
    from oracle import oracle as O
    # every worker owns a W x K float64 nPhiHat (205 MB at the C3 shape), lda.go:492-495
    b = args.batch_size
    threads = max(1, min(os.cpu_count() or 1, args.cpu_max_threads, args.group, (D + b - 1) // b))
    n_docs = min(D, threads * b)
    indptr, ind, data = corpus_synth.generate(n_docs, W, block=b, mean_distinct=md, device=device)
    o = O.LatentDirichletAllocation(K, seed=0, processes=threads)
    o.BatchSize = b
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
        "value": value, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": "first %d docs of the corpus (%d nnz) = %d minibatches of BatchSize %d, Processes = %d (one minibatch per "
                  "host thread and step; the host has %d hardware threads), %d timed + %d warm-up iterations, float64 C "
                  "restatement of lda.go (Go toolchain absent)" % (
                      n_docs, int(indptr[-1]), threads, b, threads, os.cpu_count() or 1, steps, warmup),
        "ms_per_step": 1e3 * timed / steps, "wall_s": wall,
    }


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    D, W, K, md = workload(args)
    try:
        import torch
        device = "cuda" if torch.cuda.is_available() else "cpu"
    except Exception:
        device = "cpu"
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.workload == "C5":
        _emit({"impl": "reference", "unavailable": "the Transform workload has no reference arm in bench.py (lda.go:420-437 is "
                                                   "single-threaded; see tools/bench_transform.py)"})
        return
    cb = cpu_run(args, D, W, K, md, steps, warmup, device)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": common_config(args, D, W, K, md),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------------------------
def ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import corpus_synth
    import nlp_b200
    from nlp_b200 import _lib, rand
    from nlp_b200.lda import CSC

    rank = int(os.environ.get("RANK", "0"))
    world = args.world
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D, W, K, md = workload(args)
    b, G = args.batch_size, args.group
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
    raw = corpus_synth.generate(D, W, block=b, mean_distinct=md, device=dev, owned=(lambda m: m % world == rank))
    pinned = [torch.from_numpy(a).pin_memory() for a in raw]   # page-locked host buffers
    indptr, ind, data = [t.numpy() for t in pinned]
    m = CSC(W, D, indptr, ind, data)
    nnz_local = int(indptr[-1])
    nnz = int(allsum(nnz_local))
    docs_local = sum(min(b, D - mb * b) for mb in range((D + b - 1) // b) if mb % world == rank)
    t_gen = time.perf_counter() - t_gen
    if args.workload == "C5":
        return transform_workload(args, locals())

    def new_lda(iterations, processes=G, batch=b):
        lda = nlp_b200.NewLatentDirichletAllocation(K, device=local)
        lda.Rnd = rand.New(rand.NewSource(0))
        lda.BatchSize = batch
        lda.Iterations = iterations
        lda.Processes = processes
        lda.Deterministic = args.deterministic
        return lda

    visits_per_step = nnz * (B + 1)
    L = _lib.lib()

    # ---- value: K timed iterations on the resident corpus ---------------------------------------------------------
    lda = new_lda(10 ** 6)
    lda.FitBegin(m)
    h = lda._h
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
    mb_ms, mb_n, bl_ms, bl_n = C.c_float(), C.c_int32(), C.c_float(), C.c_int32()
    L.lda_b200_last_kernel_time(h, C.byref(mb_ms), C.byref(mb_n), C.byref(bl_ms), C.byref(bl_n))
    ar_ms, ar_n = C.c_float(), C.c_int32()
    L.lda_b200_last_allreduce_time(h, C.byref(ar_ms), C.byref(ar_n))
    phases = (C.c_float * 4)()
    L.lda_b200_last_exchange_phases(h, phases)
    L.lda_b200_set_profiling(h, 0)
    kernel_ms_max, kernel_ms_min = allmax(mb_ms.value), -allmax(-mb_ms.value)
    ms = allmax(dev_ms)
    wall_ms = allmax(wall_ms)
    total_launches = int(allsum(launches))
    value = visits_per_step * K * steps / (ms * 1e-3)
    # memory floor of the minibatch kernel: its access mix replayed without arithmetic, same process, same clocks
    pe, pg, pr, pm = C.c_int64(), C.c_float(), C.c_float(), C.c_float()
    _lib.check(L.lda_b200_probe_memory_mix(h, C.byref(pe), C.byref(pg), C.byref(pr), C.byref(pm)), h)
    # one perplexity evaluation outside the timed region, for the record
    lda.PerplexityEvaluationFrequency = 1
    lda.PerplexityTolerance = 0.0
    lda.FitIterate(1)
    perplexity = lda.PerplexityTrace[-1] if lda.PerplexityTrace else None
    lda.FitEnd(want_theta=False)

    peer = lda.CommInfo()[2]
    peak, peak_src = measured_peak()
    kbytes = alg_bytes(nnz_local, docs_local, K, B) * steps      # what this rank's minibatch kernels processed
    achieved = kbytes / (mb_ms.value * 1e-3) / 1e9 if mb_ms.value > 0 else None
    row = 4 * K
    kernel_ms_per_entry = mb_ms.value / (nnz_local * steps) if nnz_local else None
    floor = {
        "what": "lda_b200_probe_memory_mix: the word ids of this rank's first launch range replayed with no arithmetic and no "
                "dependencies -- (BurnInPasses+1) coalesced reads of the row of C per entry (gather), one red.global.add.v4.f32 "
                "of a row into nPhiHat per entry (red), and both together (mix = the kernel's own traffic)",
        "entries": pe.value, "gather_ms": pg.value, "red_ms": pr.value, "mix_ms": pm.value,
        "gather_GBps": pe.value * (B + 1) * row / (pg.value * 1e-3) / 1e9 if pg.value else None,
        "red_GBps": pe.value * row / (pr.value * 1e-3) / 1e9 if pr.value else None,
        "mix_GBps": pe.value * (B + 2) * row / (pm.value * 1e-3) / 1e9 if pm.value else None,
        "kernel_ns_per_entry": 1e6 * kernel_ms_per_entry if kernel_ms_per_entry else None,
        "floor_ns_per_entry": 1e6 * pm.value / pe.value if pe.value else None,
        "kernel_frac_of_floor": (pm.value / pe.value) / kernel_ms_per_entry if (pe.value and kernel_ms_per_entry) else None,
    }
    roofline = {
        "bound": "hbm", "kernel": "scvb0_docs32_kernel (fused burn-in + sufficient-statistics pass; one launch covers the %d "
                                  "minibatches of a step on this GPU)" % (G // world),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
        "traffic": committed_traffic(K, b * (G // world), world), "peak_source": peak_src,
        "note": "achieved counts ALGORITHMIC bytes (SURVEY 8d: a 4K-byte row of C per token visit, 4K more of reductions per "
                "main visit).  C and nPhiHat (2 x 102 MB at K=256) live largely in the 126 MB L2, so most of those bytes never "
                "reach DRAM (`traffic`) and frac exceeds 1: HBM is not what bounds this kernel.  What does is in "
                "`memory_floor`: the L2 executes fp32 reductions at `red_GBps` whatever their form, and the kernel's own "
                "read/reduction mix cannot run faster than `mix_ms` per launch range; kernel_frac_of_floor = floor / measured",
        "memory_floor": floor,
        "launches": mb_n.value, "avg_launch_ms": (mb_ms.value / mb_n.value) if mb_n.value else None,
        "algorithmic_bytes_per_launch": kbytes / mb_n.value if mb_n.value else None,
        "kernel_share_of_step": mb_ms.value / dev_ms if dev_ms else None,
        "blend_kernel": {"launches": bl_n.value, "avg_launch_ms": (bl_ms.value / bl_n.value) if bl_n.value else None,
                         "achieved_GBps": (24.0 * K * W * bl_n.value / (bl_ms.value * 1e-3) / 1e9) if bl_ms.value else None},
        "exchange": {"path": ("reduce + M-step kernel, reduction and broadcast in the NVSwitch (NVLink multicast, multimem.ld_reduce / "
                              "multimem.st)" if peer == 2 else "reduce + M-step kernel over unicast peer memory (CUDA IPC over NVLink)")
                     if peer else ("ncclAllReduce + blend kernel" if world > 1 else "none (1 GPU)"), "calls": ar_n.value,
                     "avg_ms_incl_wait_for_slowest_rank": (ar_ms.value / ar_n.value) if ar_n.value else None,
                     "phases_ms_per_step_rank0": dict(zip(
                         ("opening_barrier_wait_for_slowest_rank", "reduce_blend_over_peer_memory",
                          "column_sums_closing_barrier_nz", "refresh_c"),
                         [x / max(ar_n.value, 1) for x in phases])) if peer else None,
                     "bytes": 4 * K * W, "minibatch_kernel_ms_per_launch_rank_min_max": [
                         kernel_ms_min / max(mb_n.value, 1), kernel_ms_max / max(mb_n.value, 1)]},
    }

    # ---- e2e: the public call on host buffers ---------------------------------------------------------------------------
    e2e = None
    if not args.skip_e2e:
        out_t = torch.empty((K, D), dtype=torch.float64, pin_memory=True)   # page-locked result buffer
        out = out_t.numpy()
        # the same model object a user would keep around (its communicator is already connected); a fresh RNG seed and
        # fresh counters make it the same call as on a new object
        lda.PerplexityEvaluationFrequency, lda.PerplexityTolerance = 30, 1e-2
        for _ in range(args.e2e_warmup):          # untimed: first touch of the page-locked result buffer, pool growth
            lda.Rnd = rand.New(rand.NewSource(0))
            lda.rhoPhiT, lda.rhoThetaT, lda.wordsInCorpus = 1.0, 1.0, 0.0
            lda.Iterations = 1
            lda.FitTransform(m, out=out)
        lda.Iterations = steps
        lda.Rnd = rand.New(rand.NewSource(0))
        lda.rhoPhiT, lda.rhoThetaT, lda.wordsInCorpus = 1.0, 1.0, 0.0
        barrier()
        t0 = time.perf_counter()
        lda.FitTransform(m, out=out)
        torch.cuda.synchronize()
        t_e2e = allmax(time.perf_counter() - t0)
        h2d = int(allsum(indptr.nbytes + ind.nbytes + data.nbytes))
        d2h = int(allsum(K * docs_local * 8))
        ok = bool(np.isfinite(out[:, :min(D, 64)]).all()) if rank == 0 else True
        e2e = {"value": visits_per_step * K * steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d // steps,
               "d2h_bytes_per_step": d2h // steps, "seconds": t_e2e, "iterations": steps, "finite": ok,
               "call": "LatentDirichletAllocation.FitTransform(CSC host buffers, page-locked) -> K x D float64, warm handle"}
        del out, out_t
    lda.Close()

    # ---- the same call on pageable buffers; the reference's default operating point ------------------------------------------
    e2e_pageable, ref_defaults = None, None
    if not args.skip_e2e and not args.skip_extras:
        pm_ = CSC(W, D, *[np.array(a) for a in (indptr, ind, data)])   # pageable copies (NumPy heap), like Go slices
        cold = new_lda(steps)
        barrier()
        t0 = time.perf_counter()
        cold.FitTransform(pm_)                                          # new object: handle, pools, staging ring, NCCL
        t_cold = allmax(time.perf_counter() - t0)
        cold.Rnd = rand.New(rand.NewSource(0))
        cold.rhoPhiT, cold.rhoThetaT, cold.wordsInCorpus = 1.0, 1.0, 0.0
        barrier()
        t0 = time.perf_counter()
        cold.FitTransform(pm_)                                          # result is a fresh np.zeros each call, as in Go
        t_warm = allmax(time.perf_counter() - t0)
        cold.Close()
        e2e_pageable = {"unit": UNIT, "iterations": steps,
                        "cold": {"value": visits_per_step * K * steps / t_cold, "seconds": t_cold},
                        "warm": {"value": visits_per_step * K * steps / t_warm, "seconds": t_warm},
                        "call": "FitTransform on pageable NumPy arrays (in and out): cold = first call on a new object "
                                "(CUDA context is up; handle, memory pools, page-locked staging ring%s are not), warm = second "
                                "call on it" % (", NCCL communicator, peer mappings" if world > 1 else "")}
        del pm_
    if world == 1 and not args.skip_extras and args.workload in ("C2", "C3", "C4"):
        ref_defaults = {"what": "the same corpus at the reference's default operating point: BatchSize 100 (lda.go:164), "
                                "resident corpus, one warm-up + one timed iteration; the dense M-step runs once per step"}
        for name, procs in (("processes_host_cores", os.cpu_count() or 1), ("processes_1", 1)):
            r = new_lda(10 ** 6, processes=procs, batch=100)
            r.FitBegin(m)
            r.FitIterate(1)
            _, _, t_ms = r.FitIterate(1)
            r.FitEnd(want_theta=False)
            r.Close()
            ref_defaults[name] = {"processes": procs, "batch_size": 100, "ms_per_iteration": t_ms,
                                  "m_steps_per_iteration": -(-((D + 99) // 100) // min(procs, 256)),
                                  "value": visits_per_step * K / (t_ms * 1e-3)}

    # ---- parity on a sample against the oracle (rank 0 checks) ------------------------------------------------------------------
    parity = None
    if not args.skip_parity:
        parity = parity_leg(args, new_lda, rank, world, dev, W, K, md)

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cb = cpu_run(args, D, W, K, md, 1, 0, dev)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if ref_defaults is not None:
            # the CPU port at the reference's default BatchSize too (Processes = host cores, one minibatch each)
            import copy
            a100 = copy.copy(args)
            a100.batch_size, a100.group = 100, os.cpu_count() or 1
            cb100 = cpu_run(a100, D, W, K, md, 2, 1, dev)
            ref_defaults["cpu_port_same_operating_point"] = {k: cb100[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": common_config(args, D, W, K, md),
            "clocks": clocks, "e2e": e2e, "gpu_launches": total_launches, "roofline": roofline, "cpu_baseline": cpu,
            "wall_ms_per_step": wall_ms / steps, "perplexity_after_%d_iterations" % (warmup + steps + 1): perplexity,
            "parity": parity, "e2e_pageable": e2e_pageable, "reference_defaults": ref_defaults,
            "details": {"nnz": nnz, "token_visits_per_step": visits_per_step, "corpus_generation_s": round(t_gen, 1),
                        "launches_per_gpu_per_step": -(-((D + b - 1) // b) // G), "host_threads": os.cpu_count()},
        }
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_leg(args, new_lda, rank, world, dev, W, K, md):
    """the CUDA path against the oracle on the first documents of the same corpus: Processes = 1 (the reference's own
    deterministic schedule) on one GPU, and G' = 4 minibatches per GPU from one snapshot against the oracle's
    SyncGroup = G' restatement of lda.go:487-528 -- on every N, through the exchange path the timed run used"""
    import torch
    import torch.distributed as dist

    import corpus_synth
    from nlp_b200.lda import CSC
    out = {}
    cases = []
    if world == 1:
        cases.append(("processes_1", 1, 128, 384))
    g = 4 * world
    bs = 128 if g < 16 else 64 if g < 32 else 32
    cases.append(("sync_group_%d" % g, g, bs, min(2 * g * bs - bs // 2, 10 ** 9)))
    for name, procs, bs, n_s in cases:
        indptr, ind, data = corpus_synth.generate(n_s, W, block=bs, mean_distinct=md, device=dev)
        sub = CSC(W, n_s, indptr, ind, data)
        f = dict(Iterations=2, BatchSize=bs, PerplexityEvaluationFrequency=1, PerplexityTolerance=0.0)
        # the sample is small, and small launch ranges would take the latency-optimised (two tokens per butterfly) kernels:
        # keep the parity leg on the kernels the timed run used (the handle reads the variable when it is created)
        os.environ["LDA_B200_PAIR_MAX_DOCS"] = "0"
        a = new_lda(2, processes=procs, batch=bs)
        for k_, v_ in f.items():
            setattr(a, k_, v_)
        got = a.FitTransform(sub)
        os.environ.pop("LDA_B200_PAIR_MAX_DOCS", None)
        trace = list(a.PerplexityTrace)
        path = a.CommInfo()[2]
        a.Close()
        if world > 1:
            t = torch.from_numpy(got).to(dev)      # every rank holds its own documents' columns, zeros elsewhere
            dist.all_reduce(t)
            got = t.cpu().numpy()
        if rank == 0:
            from oracle import oracle as O
            o = O.LatentDirichletAllocation(K, seed=0)
            for k_, v_ in f.items():
                setattr(o, k_, v_)
            o.SyncGroup = procs
            want = o.FitTransform(tuple(sub))
            out[name] = {"sample": "first %d docs, BatchSize=%d, Processes=%d, 2 iterations, seed 0%s" % (
                             n_s, bs, procs, "" if world == 1 else ", %d GPUs, %s" % (
                                 world, {2: "NVLink-multicast exchange", 1: "peer-memory exchange", 0: "ncclAllReduce exchange"}[path])),
                         "theta_max_abs_delta": float(np.abs(got - want).max()),
                         "perplexity_rel_delta": float(np.abs(np.array(trace) / o.perplexity_trace - 1).max())}
        if world > 1:
            dist.barrier()
    return out if rank == 0 else None


def transform_workload(args, env):
    """configs[4]: Transform of held-out documents under a fixed model, sharded by minibatch, no data-path collective.
    A step is one Transform call on host buffers (the call's pipeline streams the corpus through the device)."""
    import torch
    import torch.distributed as dist

    import nlp_b200
    from nlp_b200 import rand
    D, W, K, md, m, rank, world, local, dev = (env[k] for k in ("D", "W", "K", "md", "m", "rank", "world", "local", "dev"))
    steps, warmup, nnz, t_gen = env["steps"], env["warmup"], env["nnz"], env["t_gen"]
    allmax, allsum, barrier = env["allmax"], env["allsum"], env["barrier"]
    lda = nlp_b200.NewLatentDirichletAllocation(K, device=local)
    lda.Rnd = rand.New(rand.NewSource(0))
    lda.BatchSize = args.batch_size
    lda.Processes = world
    # "a fixed trained nPhi": a seeded synthetic model with the corpus' Zipfian word marginal per topic
    g = np.random.default_rng(7)
    base = 1.0 / (np.arange(W) + 2.7) ** 1.07
    nphi = (base[:, None] * np.exp(1.5 * g.standard_normal((W, K)))) * (nnz / K / base.sum())
    lda.SetState(nphi, nphi.sum(0))
    del nphi
    times, visits = [], 0
    sampler = None
    for it in range(warmup + steps):
        if it == warmup:
            sampler = ClockSampler(local)
            sampler.start()
        lda.Rnd = rand.New(rand.NewSource(0))
        barrier()
        t0 = time.perf_counter()
        theta, passes = lda.Transform(m, return_passes=True)
        torch.cuda.synchronize()
        times.append(allmax(time.perf_counter() - t0))
        if it == warmup:
            visits = int(allsum(float((np.maximum(passes, 0).astype(np.int64) * np.diff(m.indptr)).sum())))
    clocks = sampler.stop()
    t = float(np.sum(times[warmup:]))
    ok = bool(np.abs(theta[:, :64].sum(0)[np.diff(m.indptr)[:64] > 0] - 1).max() <= 1e-8) if rank == 0 else True
    launches = int(allsum(float(nlp_b200._lib.lib().lda_b200_launch_count(lda._h))))
    lda.Close()
    if rank == 0:
        v = visits * K * steps / t
        e2e = {"value": v, "unit": UNIT, "h2d_bytes_per_step": int(nnz * 16 + (D + 1) * 8), "d2h_bytes_per_step": int(K * D * 8),
               "seconds_per_call": t / steps, "documents_per_s": D * steps / t, "columns_sum_to_one": ok,
               "call": "LatentDirichletAllocation.Transform(CSC host buffers, pageable) -> K x D float64"}
        _emit({"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": common_config(args, D, W, K, md), "clocks": clocks, "e2e": e2e,
               "gpu_launches": launches, "roofline": None, "cpu_baseline": None,
               "note": "value = e2e for this workload: the Transform call streams the held-out corpus through the device in "
                       "ranges of 65,536 documents, there is no resident-corpus variant; token visits = sum over documents of "
                       "entries x passes executed (every document stops at the first change check that passes, lda.go:251-259)",
               "details": {"nnz": nnz, "token_visits_per_step": visits, "corpus_generation_s": round(t_gen, 1)}})
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

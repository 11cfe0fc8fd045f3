"""Diagnostic (not part of the product): time a W*K fp32 all-reduce through torch.distributed/NCCL at the C3 size,
to compare with the all-reduce inside lda_b200_fit_iterate.  torchrun --nproc-per-node N tools/nccl_probe.py"""
import os
import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (25_600_000, 6_400_000, 102_400_000):
    x = torch.ones(n, device="cuda")
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if dist.get_rank() == 0:
        w = dist.get_world_size()
        print("allreduce %d floats: %.3f ms  algbw %.1f GB/s  busbw %.1f GB/s" % (n, ms, 4 * n / ms / 1e6, 4 * n / ms / 1e6 * 2 * (w - 1) / w), flush=True)
dist.destroy_process_group()

"""Is NVLink multicast (NVLS: multimem.ld_reduce / multimem.st) usable on this box?  Driver attributes, and a
torch.distributed._symmetric_memory rendezvous (multicast pointer) under torchrun."""
import os
import sys

from cuda.bindings import driver as cu


def chk(r):
    if r[0] != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(str(r[0]))
    return r[1:] if len(r) > 2 else (r[1] if len(r) == 2 else None)


def main():
    chk(cu.cuInit(0))
    n = chk(cu.cuDeviceGetCount())
    for d in range(n):
        dev = chk(cu.cuDeviceGet(d))
        mc = chk(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev))
        fab = chk(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev))
        fd = chk(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev))
        print("device %d: multicast_supported=%d fabric_handles=%d posix_fd_handles=%d" % (d, mc, fab, fd))
    if "RANK" in os.environ:
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        rank = int(os.environ["RANK"])
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
        t = symm.empty(1 << 20, dtype=torch.float32, device="cuda")
        hdl = symm.rendezvous(t, dist.group.WORLD)
        print("rank %d: symmetric memory ok, multicast_ptr=%#x buffer_ptrs=%s" % (
            rank, getattr(hdl, "multicast_ptr", 0) or 0, [hex(p) for p in hdl.buffer_ptrs]))
        dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except Exception as e:  # report, do not fail the box
        print("probe failed:", repr(e))
        sys.exit(0)

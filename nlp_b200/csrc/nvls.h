// nvls.h -- NVLink multicast (NVLS) plumbing for the multi-GPU step: one multicast object over all ranks' GPUs, backed on
// every GPU by one physical allocation that holds the two nPhiHat buffers and the nPhi replica.  A kernel then reads the
// SUM of all ranks' nPhiHat with one multimem.ld_reduce (the NVSwitch adds) and writes a value into EVERY rank's nPhi
// with one multimem.st (the switch replicates), instead of R peer loads and R peer stores per element.
//
// The CUDA driver entry points are resolved at run time (cudaGetDriverEntryPoint), so liblda_b200.so keeps loading on
// hosts without libcuda.  One process per GPU: rank 0 creates the multicast object and hands its POSIX file descriptor
// to the other processes over abstract unix sockets (SCM_RIGHTS).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <string>

namespace ldab {

struct DriverApi {
    bool ok = false;
    std::string error;
    CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;

    template <typename F>
    bool get(const char *name, F *fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            cudaGetLastError();
            error = std::string("driver entry point ") + name + " is not available";
            return false;
        }
        *fn = (F)p;
        return true;
    }
    bool load() {
        if (ok) return true;
        ok = get("cuDeviceGet", &DeviceGet) && get("cuDeviceGetAttribute", &DeviceGetAttribute) &&
             get("cuMulticastCreate", &MulticastCreate) && get("cuMulticastAddDevice", &MulticastAddDevice) &&
             get("cuMulticastBindMem", &MulticastBindMem) && get("cuMulticastUnbind", &MulticastUnbind) &&
             get("cuMulticastGetGranularity", &MulticastGetGranularity) && get("cuMemCreate", &MemCreate) &&
             get("cuMemRelease", &MemRelease) && get("cuMemAddressReserve", &MemAddressReserve) &&
             get("cuMemAddressFree", &MemAddressFree) && get("cuMemMap", &MemMap) && get("cuMemUnmap", &MemUnmap) &&
             get("cuMemSetAccess", &MemSetAccess) && get("cuMemGetAllocationGranularity", &MemGetAllocationGranularity) &&
             get("cuMemExportToShareableHandle", &MemExportToShareableHandle) &&
             get("cuMemImportFromShareableHandle", &MemImportFromShareableHandle) && get("cuGetErrorString", &GetErrorString);
        return ok;
    }
    std::string str(CUresult r) {
        const char *s = nullptr;
        if (GetErrorString && GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
        return "CUDA driver error " + std::to_string((int)r);
    }
};

inline DriverApi &drv() {
    static DriverApi api;
    return api;
}

// ---- file descriptor passing between the ranks' processes (abstract unix sockets, SCM_RIGHTS) ------------------------
inline void nvls_sock_name(sockaddr_un *a, socklen_t *len, uint64_t token, int rank) {
    memset(a, 0, sizeof *a);
    a->sun_family = AF_UNIX;
    char name[64];
    const int n = snprintf(name, sizeof name, "lda_b200_mc_%016llx_%d", (unsigned long long)token, rank);
    memcpy(a->sun_path + 1, name, (size_t)n);  // sun_path[0] == 0: abstract namespace, nothing to unlink
    *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + (size_t)n);
}
inline int nvls_listen(uint64_t token, int rank) {
    const int s = socket(AF_UNIX, SOCK_STREAM, 0);
    if (s < 0) return -1;
    sockaddr_un a;
    socklen_t len;
    nvls_sock_name(&a, &len, token, rank);
    if (bind(s, (sockaddr *)&a, len) != 0 || listen(s, 1) != 0) {
        close(s);
        return -1;
    }
    return s;
}
inline bool nvls_send_fd(uint64_t token, int to_rank, int fd) {
    const int s = socket(AF_UNIX, SOCK_STREAM, 0);
    if (s < 0) return false;
    sockaddr_un a;
    socklen_t len;
    nvls_sock_name(&a, &len, token, to_rank);
    bool ok = false;
    for (int attempt = 0; attempt < 200 && !ok; ++attempt) {  // the listener exists before the senders start (barrier); be lenient
        ok = connect(s, (sockaddr *)&a, len) == 0;
        if (!ok) {
            timespec ts = {0, 10 * 1000 * 1000};
            nanosleep(&ts, nullptr);
        }
    }
    if (ok) {
        char byte = 'x';
        iovec iov = {&byte, 1};
        char ctl[CMSG_SPACE(sizeof(int))];
        memset(ctl, 0, sizeof ctl);
        msghdr msg;
        memset(&msg, 0, sizeof msg);
        msg.msg_iov = &iov;
        msg.msg_iovlen = 1;
        msg.msg_control = ctl;
        msg.msg_controllen = sizeof ctl;
        cmsghdr *c = CMSG_FIRSTHDR(&msg);
        c->cmsg_level = SOL_SOCKET;
        c->cmsg_type = SCM_RIGHTS;
        c->cmsg_len = CMSG_LEN(sizeof(int));
        memcpy(CMSG_DATA(c), &fd, sizeof(int));
        ok = sendmsg(s, &msg, 0) == 1;
    }
    close(s);
    return ok;
}
inline int nvls_recv_fd(int listen_sock) {
    timeval tv = {30, 0};
    setsockopt(listen_sock, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
    const int c = accept(listen_sock, nullptr, nullptr);
    if (c < 0) return -1;
    setsockopt(c, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
    char byte = 0;
    iovec iov = {&byte, 1};
    char ctl[CMSG_SPACE(sizeof(int))];
    memset(ctl, 0, sizeof ctl);
    msghdr msg;
    memset(&msg, 0, sizeof msg);
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctl;
    msg.msg_controllen = sizeof ctl;
    int fd = -1;
    if (recvmsg(c, &msg, 0) == 1) {
        for (cmsghdr *h = CMSG_FIRSTHDR(&msg); h; h = CMSG_NXTHDR(&msg, h))
            if (h->cmsg_level == SOL_SOCKET && h->cmsg_type == SCM_RIGHTS) memcpy(&fd, CMSG_DATA(h), sizeof(int));
    }
    close(c);
    return fd;
}

// the multicast-backed region of one rank
struct NvlsRegion {
    bool active = false;
    size_t size = 0;
    CUmemGenericAllocationHandle mc = 0, phys = 0;
    CUdeviceptr uc_va = 0, mc_va = 0;  // this GPU's own memory / the multicast view of every GPU's memory
    bool bound = false, uc_mapped = false, mc_mapped = false;
    int device = 0;
};

inline void nvls_release(NvlsRegion &r) {
    DriverApi &d = drv();
    if (!d.ok) {
        r = NvlsRegion();
        return;
    }
    if (r.mc_mapped) d.MemUnmap(r.mc_va, r.size);
    if (r.mc_va) d.MemAddressFree(r.mc_va, r.size);
    if (r.uc_mapped) d.MemUnmap(r.uc_va, r.size);
    if (r.uc_va) d.MemAddressFree(r.uc_va, r.size);
    if (r.bound) {
        CUdevice dev;
        if (d.DeviceGet(&dev, r.device) == CUDA_SUCCESS) d.MulticastUnbind(r.mc, dev, 0, r.size);
    }
    if (r.phys) d.MemRelease(r.phys);
    if (r.mc) d.MemRelease(r.mc);
    r = NvlsRegion();
}

}  // namespace ldab

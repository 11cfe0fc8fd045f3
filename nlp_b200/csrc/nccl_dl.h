// nccl_dl.h -- NCCL bound at run time with dlopen, so that liblda_b200.so loads (and its single-GPU path
// works) on hosts without libnccl, and so that inside a torch process the already-loaded libnccl.so.2 is
// reused instead of a second copy.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace ldab {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    const char *load_error = nullptr;

    bool load() {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) {
            load_error = "libnccl.so.2 could not be loaded (dlopen)";
            return false;
        }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !AllGather || !GetErrorString) {
            load_error = "libnccl.so.2 lacks a required symbol";
            lib = nullptr;
            return false;
        }
        return true;
    }
};

inline NcclApi &nccl() {
    static NcclApi api;
    return api;
}

}  // namespace ldab

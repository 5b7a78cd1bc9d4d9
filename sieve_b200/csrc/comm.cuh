// comm.cuh -- the one exchange step of the multi-GPU path: an all-gather of the 5-double shard results.
//
// Replaces the join of the reference's worker threads (likelihood/ThreadedScsTreeLikelihood.java:366-379: after
// pool.invokeAll, logPByThread[] and constSumByThread[] are read by the parent).  One process per GPU: NCCL all-gather
// on the handle's stream (NVLink / NVSwitch), then every rank applies the same fixed-order combine.  NCCL is resolved at
// run time with dlopen, so the library has no link-time dependency and single-GPU users never load it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

struct SvNccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static SvNccl& sv_nccl() {
    static SvNccl n;
    if (n.lib || n.ok) return n;
    // an already loaded libnccl.so.2 (e.g. the one PyTorch brings) is reused: same soname
    n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) n.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) return n;
    n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(n.lib, "ncclGetUniqueId");
    n.CommInitRank = (decltype(n.CommInitRank))dlsym(n.lib, "ncclCommInitRank");
    n.CommDestroy = (decltype(n.CommDestroy))dlsym(n.lib, "ncclCommDestroy");
    n.AllGather = (decltype(n.AllGather))dlsym(n.lib, "ncclAllGather");
    n.GetErrorString = (decltype(n.GetErrorString))dlsym(n.lib, "ncclGetErrorString");
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllGather && n.GetErrorString;
    return n;
}

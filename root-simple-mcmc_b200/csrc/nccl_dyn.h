// nccl_dyn.h -- NCCL bound at run time (dlopen), so that libsmcmc_b200.so has
// no link-time dependency on it: single-GPU users never load NCCL, and a
// process that already carries a libnccl.so.2 (PyTorch bundles one) re-uses it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

#include "common.cuh"

namespace smcmc {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    void* handle = nullptr;

    static NcclApi& get() {
        static NcclApi api;
        if (api.handle) return api;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // already in the process?
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) throw Error(SMCMC_ERR_RUNTIME, std::string("cannot load libnccl.so.2: ") + dlerror());
        auto sym = [&](const char* name) {
            void* p = dlsym(h, name);
            if (!p) throw Error(SMCMC_ERR_RUNTIME, std::string("NCCL symbol missing: ") + name);
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommSplit = (decltype(api.CommSplit))sym("ncclCommSplit");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.ReduceScatter = (decltype(api.ReduceScatter))sym("ncclReduceScatter");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.CommCount = (decltype(api.CommCount))sym("ncclCommCount");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.handle = h;
        return api;
    }
    void check(ncclResult_t r, const char* what) {
        if (r != ncclSuccess) throw Error(SMCMC_ERR_RUNTIME, std::string("NCCL ") + what + ": " + GetErrorString(r));
    }
};

}  // namespace smcmc

// nccl_dl.h -- NCCL entry points resolved at run time (dlopen), so that libcnavier_b200.so carries no link-time
// dependency on NCCL and binds to whichever libnccl.so.2 the process already holds (torch's bundled one when the
// Python layer is in use).  Only the calls of the slab halo exchange and of the continuity max / min all-reduce are needed.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>

namespace cnv {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

inline const NcclApi &nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
#define CNV_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
    CNV_SYM(GetUniqueId, "ncclGetUniqueId");
    CNV_SYM(CommInitRank, "ncclCommInitRank");
    CNV_SYM(CommDestroy, "ncclCommDestroy");
    CNV_SYM(GroupStart, "ncclGroupStart");
    CNV_SYM(GroupEnd, "ncclGroupEnd");
    CNV_SYM(Send, "ncclSend");
    CNV_SYM(Recv, "ncclRecv");
    CNV_SYM(AllReduce, "ncclAllReduce");
    CNV_SYM(GetErrorString, "ncclGetErrorString");
#undef CNV_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv && api.AllReduce;
    return api;
}

}  // namespace cnv

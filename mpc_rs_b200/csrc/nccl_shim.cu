// nccl_shim.cu — see nccl_shim.h.  Only the four NCCL entry points the MPPI exchange needs.
#include <dlfcn.h>

#include "common.cuh"
#include "nccl_shim.h"

namespace mpcb {
namespace {

// ABI of nccl.h (NCCL 2.x): ncclUniqueId is 128 opaque bytes, ncclFloat64 = 8, ncclSuccess = 0.
struct UniqueId {
    char internal[128];
};
using GetUniqueIdFn = int (*)(UniqueId*);
using CommInitRankFn = int (*)(void**, int, UniqueId, int);
using AllGatherFn = int (*)(const void*, void*, size_t, int, void*, cudaStream_t);
using CommDestroyFn = int (*)(void*);
using GetErrorStringFn = const char* (*)(int);

struct Api {
    void* lib = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    AllGatherFn all_gather = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    GetErrorStringFn get_error_string = nullptr;
    bool tried = false;
};
Api g_api;

mpcb_status load() {
    if (g_api.lib) return MPCB_OK;
    if (g_api.tried) {
        set_error("NCCL library not available");
        return MPCB_NCCL_ERROR;
    }
    g_api.tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.lib) break;
    }
    if (!g_api.lib) {
        set_error("dlopen(libnccl.so.2) failed: %s", dlerror());
        return MPCB_NCCL_ERROR;
    }
    g_api.get_unique_id = (GetUniqueIdFn)dlsym(g_api.lib, "ncclGetUniqueId");
    g_api.comm_init_rank = (CommInitRankFn)dlsym(g_api.lib, "ncclCommInitRank");
    g_api.all_gather = (AllGatherFn)dlsym(g_api.lib, "ncclAllGather");
    g_api.comm_destroy = (CommDestroyFn)dlsym(g_api.lib, "ncclCommDestroy");
    g_api.get_error_string = (GetErrorStringFn)dlsym(g_api.lib, "ncclGetErrorString");
    if (!g_api.get_unique_id || !g_api.comm_init_rank || !g_api.all_gather || !g_api.comm_destroy) {
        set_error("NCCL symbols missing");
        g_api.lib = nullptr;
        return MPCB_NCCL_ERROR;
    }
    return MPCB_OK;
}

mpcb_status check(int rc, const char* what) {
    if (rc == 0) return MPCB_OK;
    set_error("%s failed: %s", what, g_api.get_error_string ? g_api.get_error_string(rc) : "nccl error");
    return MPCB_NCCL_ERROR;
}

}  // namespace

mpcb_status nccl_unique_id(char id[128]) {
    mpcb_status st = load();
    if (st != MPCB_OK) return st;
    UniqueId u;
    st = check(g_api.get_unique_id(&u), "ncclGetUniqueId");
    if (st != MPCB_OK) return st;
    memcpy(id, u.internal, 128);
    return MPCB_OK;
}

mpcb_status nccl_init_rank(void** comm, const char id[128], int rank, int world) {
    mpcb_status st = load();
    if (st != MPCB_OK) return st;
    UniqueId u;
    memcpy(u.internal, id, 128);
    return check(g_api.comm_init_rank(comm, world, u, rank), "ncclCommInitRank");
}

mpcb_status nccl_all_gather(void* comm, const double* send, double* recv, size_t count, cudaStream_t stream) {
    if (!comm || !g_api.all_gather) {
        set_error("no NCCL communicator attached");
        return MPCB_NCCL_ERROR;
    }
    return check(g_api.all_gather(send, recv, count, /*ncclFloat64*/ 8, comm, stream), "ncclAllGather");
}

void nccl_destroy(void* comm) {
    if (comm && g_api.comm_destroy) g_api.comm_destroy(comm);
}

}  // namespace mpcb

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
    const char* why = nullptr;  // set when the library is unusable
};

// Resolved once, by the first caller, behind the C++11 function-local-static guard: handles are Send and two
// threads may reach NCCL at the same time (like nvrtc_api() / library_api()).
Api load_api() {
    Api a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) {
        a.why = "dlopen(libnccl.so.2) failed";
        return a;
    }
    a.get_unique_id = (GetUniqueIdFn)dlsym(a.lib, "ncclGetUniqueId");
    a.comm_init_rank = (CommInitRankFn)dlsym(a.lib, "ncclCommInitRank");
    a.all_gather = (AllGatherFn)dlsym(a.lib, "ncclAllGather");
    a.comm_destroy = (CommDestroyFn)dlsym(a.lib, "ncclCommDestroy");
    a.get_error_string = (GetErrorStringFn)dlsym(a.lib, "ncclGetErrorString");
    if (!a.get_unique_id || !a.comm_init_rank || !a.all_gather || !a.comm_destroy) {
        a.why = "NCCL symbols missing";
        a.lib = nullptr;
        a.all_gather = nullptr;
        a.comm_destroy = nullptr;
    }
    return a;
}

const Api& api() {
    static const Api a = load_api();
    return a;
}

mpcb_status load() {
    const Api& a = api();
    if (a.lib) return MPCB_OK;
    set_error("NCCL library not available: %s", a.why ? a.why : "unknown");
    return MPCB_NCCL_ERROR;
}

mpcb_status check(int rc, const char* what) {
    if (rc == 0) return MPCB_OK;
    set_error("%s failed: %s", what, api().get_error_string ? api().get_error_string(rc) : "nccl error");
    return MPCB_NCCL_ERROR;
}

}  // namespace

mpcb_status nccl_unique_id(char id[128]) {
    mpcb_status st = load();
    if (st != MPCB_OK) return st;
    UniqueId u;
    st = check(api().get_unique_id(&u), "ncclGetUniqueId");
    if (st != MPCB_OK) return st;
    memcpy(id, u.internal, 128);
    return MPCB_OK;
}

mpcb_status nccl_init_rank(void** comm, const char id[128], int rank, int world) {
    mpcb_status st = load();
    if (st != MPCB_OK) return st;
    UniqueId u;
    memcpy(u.internal, id, 128);
    return check(api().comm_init_rank(comm, world, u, rank), "ncclCommInitRank");
}

mpcb_status nccl_all_gather(void* comm, const double* send, double* recv, size_t count, cudaStream_t stream) {
    if (!comm || !api().all_gather) {
        set_error("no NCCL communicator attached");
        return MPCB_NCCL_ERROR;
    }
    return check(api().all_gather(send, recv, count, /*ncclFloat64*/ 8, comm, stream), "ncclAllGather");
}

void nccl_destroy(void* comm) {
    if (comm && api().comm_destroy) api().comm_destroy(comm);
}

}  // namespace mpcb

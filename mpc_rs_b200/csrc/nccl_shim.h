// nccl_shim.h — NCCL reached through dlopen, so libmpc_b200.so has no link-time NCCL dependency and the
// single-GPU path never touches it.  Inside a process that already loaded an NCCL (e.g. PyTorch's bundled
// libnccl.so.2) the same library instance is reused.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/mpc_b200.h"

namespace mpcb {
mpcb_status nccl_unique_id(char id[128]);
mpcb_status nccl_init_rank(void** comm, const char id[128], int rank, int world);
// all-gather of `count` doubles per rank on `stream`
mpcb_status nccl_all_gather(void* comm, const double* send, double* recv, size_t count, cudaStream_t stream);
void nccl_destroy(void* comm);
}  // namespace mpcb

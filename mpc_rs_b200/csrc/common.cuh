// common.cuh — shared host/device helpers of libmpc_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mpc_b200.h"

namespace mpcb {

// ---- thread-local error detail (mpcb_last_error_string) ----
void set_error(const char* fmt, ...);
const char* last_error();

#define MPCB_CUDA_TRY(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::mpcb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return MPCB_CUDA_ERROR;                                                           \
        }                                                                                     \
    } while (0)

#define MPCB_REQUIRE(cond, msg)                                    \
    do {                                                           \
        if (!(cond)) {                                             \
            ::mpcb::set_error("bad argument: %s (%s)", msg, #cond); \
            return MPCB_BAD_ARG;                                   \
        }                                                          \
    } while (0)

// Number of generic model constants handed to a kernel (filled on the host in f64, see models_host.cpp).
constexpr int kModelConsts = 32;
struct ModelConsts {
    double k[kModelConsts];
    float kf[kModelConsts];  // the same constants rounded to FP32 on the host: the packed kernels read them straight
                             // from the constant bank, so that they stay in UNIFORM registers (see f32x2.cuh)
};
// Fills the constants for `model_id` from the physical parameter block, in the reference's association order.
mpcb_status build_model_consts(int model_id, const mpcb_model_params& p, double dt, ModelConsts* out);

#ifndef __CUDACC_RTC__
// Launches a step kernel with programmatic stream serialization allowed (see pdl_entry below); MPCB_PDL=0 launches plainly.
inline cudaError_t launch_pdl(const void* fn, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t stream) {
    static const bool on = [] {
        const char* e = getenv("MPCB_PDL");
        return !(e && atoi(e) == 0);
    }();
    if (!on) return cudaLaunchKernel(fn, grid, block, args, smem, stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, fn, args);
}
#endif

constexpr int kMaxHorizon = 512;
constexpr int kInlineHorizon = 256;  // u_in up to this length travels inside the kernel parameters

#ifdef __CUDACC__
// Programmatic dependent launch (the step kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization,
// api_common: launch_pdl): the NEXT launch of the stream may be scheduled onto SMs this grid has left as soon as every
// block of this grid has passed this point, and this grid touches global memory only after the PREVIOUS grid has
// completed and flushed.  What overlaps is the launch itself (block scheduling, parameter upload: 2-3 us per launch of a
// back-to-back device loop), never the work.  Both instructions are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_entry() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// ---- warp / block primitives ----
__device__ __forceinline__ double shfl_down_f64(double v, int off) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_down_sync(0xffffffffu, lo, off);
    hi = __shfl_down_sync(0xffffffffu, hi, off);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ long long shfl_down_i64(long long v, int off) {
    int lo = (int)(v & 0xffffffffll), hi = (int)(v >> 32);
    lo = __shfl_down_sync(0xffffffffu, lo, off);
    hi = __shfl_down_sync(0xffffffffu, hi, off);
    return ((long long)hi << 32) | (unsigned int)lo;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += shfl_down_f64(v, off);
    return v;  // valid in lane 0
}
__device__ __forceinline__ bool finite_f64(double v) { return fabs(v) <= 1.79769313486231570815e308; }
#endif

}  // namespace mpcb

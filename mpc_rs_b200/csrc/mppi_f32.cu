// mppi_f32.cu — kernel table of the FP32 fast path (instantiations live in mppi_f32_{L,NL,NL6}.cu).
#include "mppi_kernel.cuh"

namespace mpcb {
MppiKernelFn mppi_kernel_f32_L(int block, int noise, int vt);
MppiKernelFn mppi_kernel_f32_NL(int block, int noise, int vt);
MppiKernelFn mppi_kernel_f32_NL6(int block, int noise, int vt);

MppiKernelFn mppi_kernel_f32x2_L(int block, int noise, int vt);
MppiKernelFn mppi_kernel_f32x2_NL(int block, int noise, int vt);
MppiKernelFn mppi_kernel_f32x2_NL6(int block, int noise, int vt);

// two samples per thread, packed f32x2 arithmetic: `block` threads cover 2*block samples per batch
MppiKernelFn mppi_kernel_f32x2(int model_id, int block, int noise, int vt) {
    switch (model_id) {
        case MPCB_MODEL_L: return mppi_kernel_f32x2_L(block, noise, vt);
        case MPCB_MODEL_NL: return mppi_kernel_f32x2_NL(block, noise, vt);
        case MPCB_MODEL_NL6: return mppi_kernel_f32x2_NL6(block, noise, vt);
        default: return nullptr;
    }
}

MppiKernelFn mppi_kernel_f32(int model_id, int block, int noise, int vt) {
    switch (model_id) {
        case MPCB_MODEL_L: return mppi_kernel_f32_L(block, noise, vt);
        case MPCB_MODEL_NL: return mppi_kernel_f32_NL(block, noise, vt);
        case MPCB_MODEL_NL6: return mppi_kernel_f32_NL6(block, noise, vt);
        default: return nullptr;
    }
}
}  // namespace mpcb

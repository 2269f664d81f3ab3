// mppi_f64.cu — kernel tables of the FP64 paths: reference order (instantiations in mppi_f64_{L,NL,NL6}.cu, no FMA) and
// the fast forms (mppi_f64fast_{L,NL,NL6}.cu).
#include "mppi_kernel.cuh"

namespace mpcb {
MppiKernelFn mppi_kernel_f64_L(int block, int noise, int vt);
MppiKernelFn mppi_kernel_f64_NL(int block, int noise, int vt);
MppiKernelFn mppi_kernel_f64_NL6(int block, int noise, int vt);

MppiKernelFn mppi_kernel_f64(int model_id, int block, int noise, int vt) {
    switch (model_id) {
        case MPCB_MODEL_L: return mppi_kernel_f64_L(block, noise, vt);
        case MPCB_MODEL_NL: return mppi_kernel_f64_NL(block, noise, vt);
        case MPCB_MODEL_NL6: return mppi_kernel_f64_NL6(block, noise, vt);
        default: return nullptr;
    }
}

MppiKernelFn mppi_kernel_f64fast_L(int block, int noise, int vt);
MppiKernelFn mppi_kernel_f64fast_NL(int block, int noise, int vt);
MppiKernelFn mppi_kernel_f64fast_NL6(int block, int noise, int vt);

MppiKernelFn mppi_kernel_f64fast(int model_id, int block, int noise, int vt) {
    switch (model_id) {
        case MPCB_MODEL_L: return mppi_kernel_f64fast_L(block, noise, vt);
        case MPCB_MODEL_NL: return mppi_kernel_f64fast_NL(block, noise, vt);
        case MPCB_MODEL_NL6: return mppi_kernel_f64fast_NL6(block, noise, vt);
        default: return nullptr;
    }
}
}  // namespace mpcb

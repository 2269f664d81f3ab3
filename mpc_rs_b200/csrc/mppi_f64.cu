// mppi_f64.cu — kernel table of the FP64 reference-order path (instantiations live in mppi_f64_{L,NL,NL6}.cu).
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
}  // namespace mpcb

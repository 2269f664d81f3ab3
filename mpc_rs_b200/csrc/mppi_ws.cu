// mppi_ws.cu — kernel table of the warp-specialised FP32 path (instantiations live in mppi_ws_{L,NL,NL6}.cu).
#include "mppi_ws_kernel.cuh"

namespace mpcb {
MppiKernelFn mppi_kernel_ws_L(int variant, int noise);
MppiKernelFn mppi_kernel_ws_NL(int variant, int noise);
MppiKernelFn mppi_kernel_ws_NL6(int variant, int noise);

MppiKernelFn mppi_kernel_ws(int model_id, int variant, int noise) {
    switch (model_id) {
        case MPCB_MODEL_L: return mppi_kernel_ws_L(variant, noise);
        case MPCB_MODEL_NL: return mppi_kernel_ws_NL(variant, noise);
        case MPCB_MODEL_NL6: return mppi_kernel_ws_NL6(variant, noise);
        default: return nullptr;
    }
}
}  // namespace mpcb

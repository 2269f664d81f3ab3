// ukf_n4.cu — 4-state UKF kernels: examples/ukf-pen.rs (n=4,o=2) and mpc::ukf (n=4,o=3).  -fmad=false.
#include "ukf_kernel.cuh"

#ifndef MPCB_UKF_FAST
#define MPCB_UKF_FAST false
#define MPCB_UKF_ENTRY ukf_kernel_n4
#endif

namespace mpcb {

template <int O, int MODEL, int SQRT, int ORDER>
static UkfKernelFn pick_mode4(int mode) {
    switch (mode) {
        case UKF_PREDICT: return ukf_kernel<4, O, MODEL, SQRT, ORDER, UKF_PREDICT, MPCB_UKF_FAST>;
        case UKF_UPDATE: return ukf_kernel<4, O, MODEL, SQRT, ORDER, UKF_UPDATE, MPCB_UKF_FAST>;
        case UKF_FUSED: return ukf_kernel<4, O, MODEL, SQRT, ORDER, UKF_FUSED, MPCB_UKF_FAST>;
        default: return nullptr;
    }
}

template <int O, int MODEL>
static UkfKernelFn pick4(int sqrt_mode, int order, int mode) {
    if (sqrt_mode == MPCB_SQRT_CHOLESKY) {
        return order == MPCB_ORDER_INTERLEAVED ? pick_mode4<O, MODEL, MPCB_SQRT_CHOLESKY, MPCB_ORDER_INTERLEAVED>(mode)
                                               : pick_mode4<O, MODEL, MPCB_SQRT_CHOLESKY, MPCB_ORDER_LIBRARY>(mode);
    }
    return order == MPCB_ORDER_INTERLEAVED ? pick_mode4<O, MODEL, MPCB_SQRT_EIG, MPCB_ORDER_INTERLEAVED>(mode)
                                           : pick_mode4<O, MODEL, MPCB_SQRT_EIG, MPCB_ORDER_LIBRARY>(mode);
}

UkfKernelFn MPCB_UKF_ENTRY(int model_id, int sqrt_mode, int order, int mode) {
    switch (model_id) {
        case MPCB_MODEL_PEN_LIN: return pick4<2, MPCB_MODEL_PEN_LIN>(sqrt_mode, order, mode);
        case MPCB_MODEL_PEN_NL: return pick4<3, MPCB_MODEL_PEN_NL>(sqrt_mode, order, mode);
        default: return nullptr;
    }
}

}  // namespace mpcb

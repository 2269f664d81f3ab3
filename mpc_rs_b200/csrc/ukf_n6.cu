// ukf_n6.cu — 6-state UKF kernels: mpc::ukf2 (n=6,o=5) with the PEN6 and NL6_UKF models.  -fmad=false.
// Only the library sigma-point order (src/ukf2.rs:129-135) exists for n = 6 in the reference.
#include "ukf_kernel.cuh"

#ifndef MPCB_UKF_FAST
#define MPCB_UKF_FAST false
#define MPCB_UKF_ENTRY ukf_kernel_n6
#endif

namespace mpcb {

template <int MODEL, int SQRT>
static UkfKernelFn pick_mode6(int mode) {
    switch (mode) {
        case UKF_PREDICT: return ukf_kernel<6, 5, MODEL, SQRT, MPCB_ORDER_LIBRARY, UKF_PREDICT, MPCB_UKF_FAST>;
        case UKF_UPDATE: return ukf_kernel<6, 5, MODEL, SQRT, MPCB_ORDER_LIBRARY, UKF_UPDATE, MPCB_UKF_FAST>;
        case UKF_FUSED: return ukf_kernel<6, 5, MODEL, SQRT, MPCB_ORDER_LIBRARY, UKF_FUSED, MPCB_UKF_FAST>;
        default: return nullptr;
    }
}

UkfKernelFn MPCB_UKF_ENTRY(int model_id, int sqrt_mode, int order, int mode) {
    if (order != MPCB_ORDER_LIBRARY) return nullptr;
    const bool chol = sqrt_mode == MPCB_SQRT_CHOLESKY;
    switch (model_id) {
        case MPCB_MODEL_PEN6:
            return chol ? pick_mode6<MPCB_MODEL_PEN6, MPCB_SQRT_CHOLESKY>(mode) : pick_mode6<MPCB_MODEL_PEN6, MPCB_SQRT_EIG>(mode);
        case MPCB_MODEL_NL6_UKF:
            return chol ? pick_mode6<MPCB_MODEL_NL6_UKF, MPCB_SQRT_CHOLESKY>(mode)
                        : pick_mode6<MPCB_MODEL_NL6_UKF, MPCB_SQRT_EIG>(mode);
        default: return nullptr;
    }
}

}  // namespace mpcb

// ukf_n6_stream.cu — the fused fast kernels of the six-state filters (PEN6, NL6_UKF) in the streaming form of
// ukf_stream_kernel.cuh.  FMA contraction on (fast arithmetic); library sigma-point order only, like ukf_n6.cu.
#include "ukf_stream_kernel.cuh"

namespace mpcb {

UkfKernelFn ukf_stream_kernel_n6(int model_id, int sqrt_mode, int order, size_t* smem_bytes) {
    if (order != MPCB_ORDER_LIBRARY) return nullptr;
    *smem_bytes = ukf_stream_smem_bytes<6>();
    const bool chol = sqrt_mode == MPCB_SQRT_CHOLESKY;
    switch (model_id) {
        case MPCB_MODEL_PEN6:
            return chol ? ukf_stream_kernel<6, 5, MPCB_MODEL_PEN6, MPCB_SQRT_CHOLESKY> : ukf_stream_kernel<6, 5, MPCB_MODEL_PEN6, MPCB_SQRT_EIG>;
        case MPCB_MODEL_NL6_UKF:
            return chol ? ukf_stream_kernel<6, 5, MPCB_MODEL_NL6_UKF, MPCB_SQRT_CHOLESKY>
                        : ukf_stream_kernel<6, 5, MPCB_MODEL_NL6_UKF, MPCB_SQRT_EIG>;
        default: return nullptr;
    }
}

}  // namespace mpcb

// mppi_f32.cu — instantiates the fused MPPI kernel for real = float (all models, block sizes, noise modes).
// FP32 fast path: FMA contraction on, constants pre-folded on the host (see models.cuh).
#include "mppi_kernel.cuh"

namespace mpcb {

template <template <typename> class ModelT, int BLOCK>
static MppiKernelFn pick_noise_f32(int noise) {
    switch (noise) {
        case NOISE_GENERATE: return mppi_rollout_kernel<ModelT, float, BLOCK, NOISE_GENERATE>;
        case NOISE_GENERATE_DUMP: return mppi_rollout_kernel<ModelT, float, BLOCK, NOISE_GENERATE_DUMP>;
        case NOISE_REPLAY: return mppi_rollout_kernel<ModelT, float, BLOCK, NOISE_REPLAY>;
        default: return nullptr;
    }
}

template <template <typename> class ModelT>
static MppiKernelFn pick_f32(int block, int noise) {
    switch (block) {
        case 128: return pick_noise_f32<ModelT, 128>(noise);
        case 64: return pick_noise_f32<ModelT, 64>(noise);
        case 32: return pick_noise_f32<ModelT, 32>(noise);
        default: return nullptr;
    }
}

MppiKernelFn mppi_kernel_f32(int model_id, int block, int noise) {
    switch (model_id) {
        case MPCB_MODEL_L: return pick_f32<ModelL>(block, noise);
        case MPCB_MODEL_NL: return pick_f32<ModelNL>(block, noise);
        case MPCB_MODEL_NL6: return pick_f32<ModelNL6>(block, noise);
        default: return nullptr;
    }
}

}  // namespace mpcb

// mppi_inst.cuh — one translation unit per (precision, model) instantiates the fused MPPI kernel for every block
// size and noise mode, so the kernels compile in parallel.  Included by mppi_f32_*.cu / mppi_f64_*.cu after
// defining MPCB_INST_MODEL, MPCB_INST_REAL and MPCB_INST_FN.
#include "mppi_kernel.cuh"

namespace mpcb {

#ifndef MPCB_INST_SPT
#define MPCB_INST_SPT 1
#endif

template <int BLOCK>
static MppiKernelFn inst_pick_noise(int noise) {
    switch (noise) {
        case NOISE_GENERATE: return mppi_rollout_kernel<MPCB_INST_MODEL, MPCB_INST_REAL, BLOCK, NOISE_GENERATE, MPCB_INST_SPT>;
        case NOISE_GENERATE_DUMP: return mppi_rollout_kernel<MPCB_INST_MODEL, MPCB_INST_REAL, BLOCK, NOISE_GENERATE_DUMP, MPCB_INST_SPT>;
        case NOISE_REPLAY: return mppi_rollout_kernel<MPCB_INST_MODEL, MPCB_INST_REAL, BLOCK, NOISE_REPLAY, MPCB_INST_SPT>;
        default: return nullptr;
    }
}

// `block` = threads per block; a block covers block * MPCB_INST_SPT samples per batch
MppiKernelFn MPCB_INST_FN(int block, int noise) {
    switch (block) {
#ifdef MPCB_INST_BLOCK512
        case 512: return inst_pick_noise<512>(noise);
#endif
        case 256: return inst_pick_noise<256>(noise);
        case 128: return inst_pick_noise<128>(noise);
        case 64: return inst_pick_noise<64>(noise);
        case 32: return inst_pick_noise<32>(noise);
        default: return nullptr;
    }
}

}  // namespace mpcb

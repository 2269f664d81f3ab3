// mppi_inst.cuh — one translation unit per (precision, model) instantiates the fused MPPI kernel for every block
// size and noise mode, so the kernels compile in parallel.  Included by mppi_f32_*.cu / mppi_f64_*.cu after
// defining MPCB_INST_MODEL, MPCB_INST_REAL and MPCB_INST_FN.
#include "mppi_kernel.cuh"

namespace mpcb {

#ifndef MPCB_INST_SPT
#define MPCB_INST_SPT 1
#endif

template <int BLOCK, bool VT>
static MppiKernelFn inst_pick_noise(int noise) {
    switch (noise) {
        case NOISE_GENERATE: return mppi_rollout_kernel<MPCB_INST_MODEL, MPCB_INST_REAL, BLOCK, NOISE_GENERATE, MPCB_INST_SPT, VT>;
        case NOISE_GENERATE_DUMP: return mppi_rollout_kernel<MPCB_INST_MODEL, MPCB_INST_REAL, BLOCK, NOISE_GENERATE_DUMP, MPCB_INST_SPT, VT>;
        case NOISE_REPLAY: return mppi_rollout_kernel<MPCB_INST_MODEL, MPCB_INST_REAL, BLOCK, NOISE_REPLAY, MPCB_INST_SPT, VT>;
        default: return nullptr;
    }
}

// `block` = threads per block (a block covers block * MPCB_INST_SPT samples per batch); vt = keep the v tile.
// With the tile: 128/256/512 samples per block.  Without: 128 (multi-batch long horizons) or 64 samples per block
// (a horizon so long that not even a 128-sample tile fits).
MppiKernelFn MPCB_INST_FN(int block, int noise, int vt) {
    constexpr int kS = MPCB_INST_SPT;
    if (vt) {
        switch (block * kS) {
#ifdef MPCB_INST_SAMPLES512
            case 512: return inst_pick_noise<512 / kS, true>(noise);
#endif
            case 256: return inst_pick_noise<256 / kS, true>(noise);
            case 128: return inst_pick_noise<128 / kS, true>(noise);
            default: return nullptr;
        }
    }
    switch (block * kS) {
        case 128: return inst_pick_noise<128 / kS, false>(noise);
        case 64: return inst_pick_noise<64 / kS, false>(noise);
        default: return nullptr;
    }
}

}  // namespace mpcb

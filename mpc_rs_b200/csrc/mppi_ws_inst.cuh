// mppi_ws_inst.cuh — one translation unit per model instantiates the warp-specialised FP32 kernels
// (mppi_ws_kernel.cuh) for every variant of kWsVariants and every noise mode.  Included by mppi_ws_{L,NL,NL6}.cu
// after defining MPCB_INST_MODEL and MPCB_INST_FN.
#include "mppi_ws_kernel.cuh"

namespace mpcb {

template <int NCW, int NPW, int SPT>
static MppiKernelFn ws_pick_noise(int noise) {
    switch (noise) {
        case NOISE_GENERATE: return mppi_ws_kernel<MPCB_INST_MODEL, NCW, NPW, NOISE_GENERATE, SPT>;
        case NOISE_GENERATE_DUMP: return mppi_ws_kernel<MPCB_INST_MODEL, NCW, NPW, NOISE_GENERATE_DUMP, SPT>;
        case NOISE_REPLAY: return mppi_ws_kernel<MPCB_INST_MODEL, NCW, NPW, NOISE_REPLAY, SPT>;
        default: return nullptr;
    }
}

MppiKernelFn MPCB_INST_FN(int variant, int noise) {
    switch (variant) {
        case 0: return ws_pick_noise<kWsVariants[0].ncw, kWsVariants[0].npw, kWsVariants[0].spt>(noise);
        case 1: return ws_pick_noise<kWsVariants[1].ncw, kWsVariants[1].npw, kWsVariants[1].spt>(noise);
        case 2: return ws_pick_noise<kWsVariants[2].ncw, kWsVariants[2].npw, kWsVariants[2].spt>(noise);
        case 3: return ws_pick_noise<kWsVariants[3].ncw, kWsVariants[3].npw, kWsVariants[3].spt>(noise);
        case 4: return ws_pick_noise<kWsVariants[4].ncw, kWsVariants[4].npw, kWsVariants[4].spt>(noise);
        case 5: return ws_pick_noise<kWsVariants[5].ncw, kWsVariants[5].npw, kWsVariants[5].spt>(noise);
        case 6: return ws_pick_noise<kWsVariants[6].ncw, kWsVariants[6].npw, kWsVariants[6].spt>(noise);
        case 7: return ws_pick_noise<kWsVariants[7].ncw, kWsVariants[7].npw, kWsVariants[7].spt>(noise);
        default: return nullptr;
    }
}

}  // namespace mpcb

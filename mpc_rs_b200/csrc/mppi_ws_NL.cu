// mppi_ws_NL.cu — warp-specialised FP32 kernels (mppi_ws_kernel.cuh), model NL.
#define MPCB_INST_MODEL ModelNL
#define MPCB_INST_FN mppi_kernel_ws_NL
#include "mppi_ws_inst.cuh"

// mppi_ws_NL6.cu — warp-specialised FP32 kernels (mppi_ws_kernel.cuh), model NL6.
#define MPCB_INST_MODEL ModelNL6
#define MPCB_INST_FN mppi_kernel_ws_NL6
#include "mppi_ws_inst.cuh"
